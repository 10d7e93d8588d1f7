#!/bin/bash
# GPU call 3 of round 2 (1 GPU): full -m gpu suite with the flattened list-build kernel, force-kernel variant battery
# (branchless / SoA / Newton-3 half list with red.add), bench, ncu of build + force kernels, launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests3.txt 2>&1
tail -5 gpurun_out/tests3.txt
timeout 600 python tools/force_lab.py --variants 0,1,3,4,6,7,10,11,12,13,14,15,16,19,20 --carveouts -1 --steps 40 > gpurun_out/lab3.txt 2>&1
cat gpurun_out/lab3.txt
timeout 300 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench3.json").read().strip().splitlines()[-1])
print("bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_build_list -s 3 -c 1 -o gpurun_out/r2c_build python bench.py --steps 40 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2c_build.ncu-rep > gpurun_out/r2c_build.txt 2>&1
cat gpurun_out/r2c_build.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/ | head -30
