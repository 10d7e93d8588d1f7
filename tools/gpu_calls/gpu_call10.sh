#!/bin/bash
# GPU call 10 of round 2 (8 GPUs): 4-rank parity test, 8-GPU weak / strong LJ, 4-GPU weak, 64M LJ + coul_sf (configs[4])
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
free -g | head -2
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/tests10.txt 2>&1
tail -4 gpurun_out/tests10.txt
show() {
  python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench10_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {"value": 0, "ms_per_step": 0}
    print(sys.argv[1], "bench: N %d value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.3f ms) launches %d %s mem %.1f GB" % (d["config"]["atoms_total"], d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], e["value"], e["ms_per_step"], d["gpu_launches"], d["timing"]["comm_mode"], d["timing"]["device_memory_used_bytes"] / 1e9))
    print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()}, "parity", d.get("parity", {}).get("U_rel"), d.get("parity", {}).get("pairs_equal"), "U", d["state"]["U"], "builds", d["timing"]["list_builds_in_timed_region"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() {  # tag nproc args...
  tag=$1; np=$2; shift; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np "$@" > gpurun_out/bench10_$tag.json 2> gpurun_out/bench10_$tag.err
  tail -2 gpurun_out/bench10_$tag.err | cut -c1-400; show $tag
}
run 8gpu_weak 8 --steps 200 --warmup 30
run 4gpu_weak 4 --steps 200 --warmup 30
run 8gpu_strong1M 8 --steps 200 --warmup 30 --scaling strong
run 8gpu_64M_coul 8 --steps 40 --warmup 10 --workload lj_coul_sf --atoms-per-gpu 8000000 --no-e2e
