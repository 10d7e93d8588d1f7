#!/bin/bash
# GPU call 18 of round 2 (1 GPU): typed kernel with constant-bank exp / erfc -- GPU suite, SPC/E timing, 1-GPU bench line
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests18.txt 2>&1; tail -3 gpurun_out/tests18.txt
timeout 300 python tools/spce_lab.py --variants 0 > gpurun_out/spce_lab18.txt 2>&1; cat gpurun_out/spce_lab18.txt
timeout 300 python bench.py --steps 200 --warmup 30 > gpurun_out/bench18_1gpu.json 2> gpurun_out/bench18_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench18_1gpu.json").read().strip().splitlines()[-1])
print("1gpu bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
print("kernel ms/step", d["timing"]["kernel_ms_per_step"]); print("spce", d["spce"]["value"], d["spce"]["ms_per_step"], d["spce"]["timing"], d["spce"]["e2e"]["value"], d["spce"]["roofline"]["frac"], d["spce"]["state"], d["spce"].get("cpu_baseline", {}).get("sample"))
print("roofline", d["roofline"]["frac"], d["roofline_fp64"]["frac"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
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
timeout 300 ncu --set full --clock-control none -k regex:k_pair_forces_typed -s 2 -c 1 -o gpurun_out/r2g_typed python /tmp/spce_one.py > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2g_typed.ncu-rep > gpurun_out/r2g_typed.txt 2>&1; cat gpurun_out/r2g_typed.txt
rm -f gpurun_out/r2g_typed.ncu-rep
