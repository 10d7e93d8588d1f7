#!/bin/bash
# GPU call 16 of round 2 (1 GPU): A/B of the software-pipelined list build and of 2 vs 4 atoms per thread in the streaming kernels
set -u
mkdir -p gpurun_out
show() {
python - $1 <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench16_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "value %.4e ms/step %.4f force_ms %.4f build_ms %.4f launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["gpu_launches"]))
    print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()})
except Exception as e:
    print(tag, "FAILED", e)
PY
}
for rep in a b; do
timeout 200 python bench.py --steps 400 --warmup 30 --no-cpu-baseline --no-spce --no-parity --no-e2e > gpurun_out/bench16_apt4$rep.json 2> gpurun_out/bench16_apt4$rep.err; show apt4$rep
done
cp emdee_b200/lib/libemdee.so /tmp/libemdee_apt4.so
cp emdee_b200/lib_apt2/libemdee.so emdee_b200/lib/libemdee.so
for rep in a b; do
timeout 200 python bench.py --steps 400 --warmup 30 --no-cpu-baseline --no-spce --no-parity --no-e2e > gpurun_out/bench16_apt2$rep.json 2> gpurun_out/bench16_apt2$rep.err; show apt2$rep
done
cp /tmp/libemdee_apt4.so emdee_b200/lib/libemdee.so
