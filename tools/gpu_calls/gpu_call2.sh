#!/bin/bash
# GPU call 2 of round 2: force-kernel variant battery + parity of the refactored step path + one ncu capture
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_zz_edge_cases.py tests/test_zzz_rigid_bodies.py -m gpu -x -q > gpurun_out/tests2.txt 2>&1
tail -5 gpurun_out/tests2.txt
timeout 600 python tools/force_lab.py --variants 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16 --carveouts -1 --steps 40 > gpurun_out/lab2.txt 2>&1
timeout 300 python tools/force_lab.py --variants 0,1 --carveouts 0,25,50,100 --steps 40 >> gpurun_out/lab2.txt 2>&1
cat gpurun_out/lab2.txt
timeout 300 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench2.json").read().strip().splitlines()[-1])
print("bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
PY
EMDEE_PROFILE=1 timeout 200 python tools/phase_profile.py > gpurun_out/phase2.txt 2>&1; tail -3 gpurun_out/phase2.txt
timeout 400 ncu --set full --clock-control none -k regex:k_pair_forces -s 30 -c 1 -o gpurun_out/r2b_force python bench.py --steps 40 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2b_force.ncu-rep > gpurun_out/r2b_force.txt 2>&1
cat gpurun_out/r2b_force.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/ | head -30
du -sh gpurun_out
