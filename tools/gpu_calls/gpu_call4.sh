#!/bin/bash
# GPU call 4 of round 2 (1 GPU): force-kernel variants on a healthy trajectory (probes last), list-build forms, full bench line
set -u
mkdir -p gpurun_out
timeout 600 python tools/force_lab.py --variants 0,1,6,7,10,11,12,13,14,15,16,20,19,0 --carveouts -1 --steps 40 --build-variants 0,1,0 > gpurun_out/lab4.txt 2>&1
cat gpurun_out/lab4.txt
timeout 600 python bench.py --steps 200 --warmup 30 > gpurun_out/bench4.json 2> gpurun_out/bench4.err
tail -3 gpurun_out/bench4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench4.json").read().strip().splitlines()[-1])
print("bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
print("kernel ms/step", d["timing"]["kernel_ms_per_step"])
print("parity", d.get("parity"))
print("spce", {k: d["spce"][k] for k in ("value", "ms_per_step", "timing")} if "spce" in d else None, d.get("spce", {}).get("e2e"))
print("cpu", d.get("cpu_baseline"))
PY
