#!/bin/bash
# GPU call 22 of round 2 (2 GPUs): coalesced peer halo push -- multi-GPU parity, 2-GPU lines with small and large halos
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/tests22.txt 2>&1; tail -3 gpurun_out/tests22.txt
show() {
  python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench22_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {"value": 0, "ms_per_step": 0}
    print(sys.argv[1], "bench: N %d value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.3f ms) launches %d %s" % (d["config"]["atoms_total"], d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], e["value"], e["ms_per_step"], d["gpu_launches"], d["timing"]["comm_mode"]))
    print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()}, "parity", d.get("parity", {}).get("U_rel"), d.get("parity", {}).get("pairs_equal"), "U", d["state"]["U"], "builds", d["timing"]["list_builds_in_timed_region"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() {
  tag=$1; np=$2; shift; shift
  EMDEE_BENCH_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np "$@" > gpurun_out/bench22_$tag.json 2> gpurun_out/bench22_$tag.err
  grep "bench trace r0" gpurun_out/bench22_$tag.err | cut -c1-300; tail -1 gpurun_out/bench22_$tag.err | cut -c1-300; show $tag
}
run 2gpu_weak 2 --steps 200 --warmup 30 --no-e2e
run 2gpu_coul8M 2 --steps 30 --warmup 8 --workload lj_coul_sf --atoms-per-gpu 8000000 --no-e2e --no-parity
