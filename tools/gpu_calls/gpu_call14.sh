#!/bin/bash
# GPU call 14 of round 2 (1 GPU): ncu of the streaming kernels (kick, drift, refresh) and of the typed SPC/E kernel
set -u
mkdir -p gpurun_out
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis"
for kn in k_boost k_displace k_refresh_positions; do
timeout 200 ncu $SEC --clock-control none -k regex:$kn -s 20 -c 1 -o gpurun_out/r2e_$kn python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
ncu -i gpurun_out/r2e_$kn.ncu-rep --page details 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|L2 Cache Throughput|Achieved Occupancy|Registers Per|Mem Busy|Max Bandwidth|L1/TEX Hit|L2 Hit|Mem Pipes|Issued Warp|No Eligible|Executed Ipc|Theoretical Occ" | head -24
done
cat > /tmp/spce_one.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import common as cm
from emdee_b200 import api
lib = api.load()
s, c = cm.spce_sample_system(lib, lambda l: l.EmDee_coul_damped_square_smoothed(0.2, 1.0), replicas=8, threads=1)
for _ in range(3):
    s.upload("coordinates", c["R"]); s.compute_forces()
s.finalize()
PY
timeout 300 ncu --set full --clock-control none -k regex:k_pair_forces_typed -s 2 -c 1 -o gpurun_out/r2e_typed python /tmp/spce_one.py > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2e_typed.ncu-rep > gpurun_out/r2e_typed.txt 2>&1; cat gpurun_out/r2e_typed.txt
ncu -i gpurun_out/r2e_typed.ncu-rep --page details 2>/dev/null | grep -E "Achieved Occupancy|Theoretical Occ|Executed Ipc|No Eligible|Issued Warp|FP64|Pipe" | head -20
rm -f gpurun_out/r2e_typed.ncu-rep
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/tests14.txt 2>&1; tail -3 gpurun_out/tests14.txt
timeout 300 python bench.py --steps 200 --warmup 30 > gpurun_out/bench14_1gpu.json 2> gpurun_out/bench14_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench14_1gpu.json").read().strip().splitlines()[-1])
print("1gpu bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
print("kernel ms/step", d["timing"]["kernel_ms_per_step"]); print("spce", d["spce"]["value"], d["spce"]["ms_per_step"], d["spce"]["timing"], d["spce"]["e2e"]["value"], d["spce"]["roofline"]["frac"], d["spce"].get("cpu_baseline", {}).get("value"))
print("roofline", d["roofline"]["frac"], d["roofline_fp64"]["frac"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
du -sh gpurun_out
