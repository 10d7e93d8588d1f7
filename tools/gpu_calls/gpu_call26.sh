#!/bin/bash
# GPU call 26 of round 2 (8 GPUs): final multi-GPU lines -- 8-GPU weak / strong LJ, 64M LJ + coul_sf (configs[4])
set -u
mkdir -p gpurun_out
show() {
  python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench26_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {"value": 0, "ms_per_step": 0}
    print(sys.argv[1], "bench: N %d value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.3f ms) launches %d %s mem %.1f GB" % (d["config"]["atoms_total"], d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], e["value"], e["ms_per_step"], d["gpu_launches"], d["timing"]["comm_mode"], d["timing"]["device_memory_used_bytes"] / 1e9))
    print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()}, "parity", d.get("parity", {}).get("U_rel"), d.get("parity", {}).get("pairs_equal"), "U", d["state"]["U"], "builds", d["timing"]["list_builds_in_timed_region"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() {
  tag=$1; np=$2; shift; shift
  EMDEE_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np "$@" > gpurun_out/bench26_$tag.json 2> gpurun_out/bench26_$tag.err
  grep "bench trace r0" gpurun_out/bench26_$tag.err | cut -c1-400; tail -1 gpurun_out/bench26_$tag.err | cut -c1-300; show $tag
}
run 8gpu_weak 8 --steps 200 --warmup 30
run 8gpu_64M_coul 8 --steps 40 --warmup 10 --workload lj_coul_sf --atoms-per-gpu 8000000 --no-e2e
