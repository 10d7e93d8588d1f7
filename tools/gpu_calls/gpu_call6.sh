#!/bin/bash
# GPU call 6 of round 2 (2 GPUs): NVLink peer path -- multi-GPU parity (mgpu_check incl. local I/O), 2-GPU bench weak (peer vs NCCL) + strong
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/tests6.txt 2>&1
tail -12 gpurun_out/tests6.txt
run() {  # tag, extra env, args
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 30 $ARGS > gpurun_out/bench6_$tag.json 2> gpurun_out/bench6_$tag.err
  tail -2 gpurun_out/bench6_$tag.err | cut -c1-300
  python - $tag <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench6_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.3f ms) launches %d %s" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["timing"]["comm_mode"]))
    print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()}, "parity", d["parity"]["U_rel"], d["parity"]["pairs_equal"], "U", d["state"]["U"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
ARGS="--scaling weak" run weak_peer EMDEE_X=1
ARGS="--scaling weak" run weak_nccl EMDEE_NO_PEER=1
ARGS="--scaling strong" run strong_peer EMDEE_X=1
