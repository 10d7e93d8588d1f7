#!/bin/bash
# GPU call 11 of round 2 (4 GPUs): 2- and 4-rank parity test on the compact rebuild + peer path, 4-GPU weak bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/tests11.txt 2>&1
tail -4 gpurun_out/tests11.txt; grep "mgpu\]" gpurun_out/tests11.txt | tail -12
show() {
  python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench11_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {"value": 0, "ms_per_step": 0}
    print(sys.argv[1], "bench: N %d value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.3f ms) launches %d %s mem %.1f GB" % (d["config"]["atoms_total"], d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], e["value"], e["ms_per_step"], d["gpu_launches"], d["timing"]["comm_mode"], d["timing"]["device_memory_used_bytes"] / 1e9))
    print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()}, "parity", d.get("parity", {}).get("U_rel"), d.get("parity", {}).get("pairs_equal"), "U", d["state"]["U"], "builds", d["timing"]["list_builds_in_timed_region"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() {
  tag=$1; np=$2; shift; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np "$@" > gpurun_out/bench11_$tag.json 2> gpurun_out/bench11_$tag.err
  tail -2 gpurun_out/bench11_$tag.err | cut -c1-400; show $tag
}
run 4gpu_weak 4 --steps 200 --warmup 30
run 4gpu_strong 4 --steps 200 --warmup 30 --scaling strong
