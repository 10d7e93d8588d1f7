#!/bin/bash
# GPU call 7 of round 2 (2 GPUs): full -m gpu suite (numpy-golden rows, planned kick, multi-GPU test) + benches 1 and 2 GPUs
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests7.txt 2>&1
tail -6 gpurun_out/tests7.txt
show() {
  python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench7_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.3f ms) launches %d %s" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["timing"]["comm_mode"]))
    print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()}, "parity", d["parity"]["U_rel"], d["parity"]["pairs_equal"], "U", d["state"]["U"], "K", d["state"]["K"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
timeout 300 python bench.py --steps 200 --warmup 30 --no-cpu-baseline --no-spce > gpurun_out/bench7_1gpu.json 2> gpurun_out/bench7_1gpu.err; tail -2 gpurun_out/bench7_1gpu.err | cut -c1-300; show 1gpu
for mode in weak strong; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 30 --scaling $mode > gpurun_out/bench7_2gpu_$mode.json 2> gpurun_out/bench7_2gpu_$mode.err
tail -2 gpurun_out/bench7_2gpu_$mode.err | cut -c1-300; show 2gpu_$mode
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 --workload lj_coul_sf > gpurun_out/bench7_2gpu_coul.json 2> gpurun_out/bench7_2gpu_coul.err
tail -2 gpurun_out/bench7_2gpu_coul.err | cut -c1-300; show 2gpu_coul
