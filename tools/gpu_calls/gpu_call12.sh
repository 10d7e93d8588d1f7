#!/bin/bash
# GPU call 12 of round 2 (4 GPUs): where does the time of the 4-GPU weak run go? per-step host trace + host phase timers
set -u
mkdir -p gpurun_out
for rep in 1 2; do
EMDEE_BENCH_TRACE=1 EMDEE_PROFILE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 200 --warmup 30 --no-e2e --no-parity > gpurun_out/bench12_$rep.json 2> gpurun_out/bench12_$rep.err
grep "bench trace\|emdee profile" gpurun_out/bench12_$rep.err | cut -c1-700
python - $rep <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench12_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("bench: N %d ms/step %.4f builds %d" % (d["config"]["atoms_total"], d["ms_per_step"], d["timing"]["list_builds_in_timed_region"]), {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()})
PY
done
