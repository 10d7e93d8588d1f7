#!/bin/bash
# GPU call 27 of round 2 (1 GPU): final validation -- GPU suite, smoke, bench line, launch list
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/tests27.txt 2>&1; tail -3 gpurun_out/tests27.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 200 --warmup 30 > gpurun_out/bench27_1gpu.json 2> gpurun_out/bench27_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench27_1gpu.json").read().strip().splitlines()[-1])
print("1gpu bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.4f ms) launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
print("kernel ms/step", d["timing"]["kernel_ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline_fp64"]["frac"], "clocks", d["clocks"])
print("spce", d["spce"]["value"], d["spce"]["ms_per_step"], d["spce"]["timing"], "e2e", d["spce"]["e2e"]["value"], d["spce"]["roofline"]["frac"])
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
ls -la gpurun_out/r2h_launches.csv
