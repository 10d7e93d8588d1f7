#!/bin/bash
# GPU call 20 of round 2 (1 GPU): launch shapes of the single-type LJ + coul_sf kernel; bench line with the SPC/E host-buffer arm
# uploading pre-made pinned frames
set -u
mkdir -p gpurun_out
for v in 0 31 32 33 34 35 0; do
EMDEE_FORCE_VARIANT=$v timeout 200 python bench.py --steps 100 --warmup 20 --workload lj_coul_sf --atoms-per-gpu 1000000 --no-cpu-baseline --no-spce --no-e2e --no-parity > gpurun_out/bench20_coul$v.json 2> gpurun_out/bench20_coul$v.err
python - $v <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench20_coul{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("coul variant", sys.argv[1], "ms/step %.4f force_ms %.4f U %r" % (d["ms_per_step"], d["timing"]["force_kernel_ms"], d["state"]["U"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
timeout 300 python bench.py --steps 200 --warmup 30 > gpurun_out/bench20_1gpu.json 2> gpurun_out/bench20_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench20_1gpu.json").read().strip().splitlines()[-1])
print("1gpu bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
print("spce", d["spce"]["value"], d["spce"]["ms_per_step"], d["spce"]["timing"], "e2e", d["spce"]["e2e"], d["spce"]["roofline"]["frac"], d["spce"]["state"])
PY
