#!/bin/bash
# GPU call 24 of round 2 (1 GPU): compact-record pair kernel (force_variant 50) against the shipped kernel; two-wide list build (40)
set -u
mkdir -p gpurun_out
timeout 300 python tools/force_lab.py --variants 0,50,0,50,40,0 --steps 60 > gpurun_out/lab24.txt 2>&1; cat gpurun_out/lab24.txt
EMDEE_FORCE_VARIANT=50 timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline --no-spce --no-e2e > gpurun_out/bench24_rec16.json 2> gpurun_out/bench24_rec16.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench24_rec16.json").read().strip().splitlines()[-1])
print("rec16 bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["gpu_launches"]))
print("kernel ms/step", d["timing"]["kernel_ms_per_step"]); print("parity", d.get("parity"))
PY
EMDEE_FORCE_VARIANT=50 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair_forces_rec16 -s 30 -c 1 -o gpurun_out/r2h_rec16 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2h_rec16.ncu-rep > gpurun_out/r2h_rec16.txt 2>&1; cat gpurun_out/r2h_rec16.txt
