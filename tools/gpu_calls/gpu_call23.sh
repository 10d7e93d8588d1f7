#!/bin/bash
# GPU call 23 of round 2 (1 GPU): GPU suite; two-wide list-build lab; bench line (speculative launch after uploads in the e2e arm)
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/tests23.txt 2>&1; tail -3 gpurun_out/tests23.txt
for v in 0 40 0 40; do
EMDEE_FORCE_VARIANT=$v timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-spce --no-e2e --no-parity > gpurun_out/bench23_v$v.json 2> gpurun_out/bench23_v$v.err
python - $v <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench23_v{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("build variant", sys.argv[1], "ms/step %.4f force_ms %.4f build_ms %.4f U %r" % (d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["state"]["U"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
timeout 300 python bench.py --steps 200 --warmup 30 > gpurun_out/bench23_1gpu.json 2> gpurun_out/bench23_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench23_1gpu.json").read().strip().splitlines()[-1])
print("1gpu bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.4f ms) launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
print("spce", d["spce"]["value"], d["spce"]["ms_per_step"], d["spce"]["timing"], "e2e", d["spce"]["e2e"]["value"], d["spce"]["roofline"]["frac"])
PY
