#!/bin/bash
# GPU call 5 of round 2 (2 GPUs): multi-GPU parity (mgpu_check incl. local I/O) and the 2-GPU bench line, weak + strong
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/tests5.txt 2>&1
tail -25 gpurun_out/tests5.txt
for mode in weak strong; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 30 --scaling $mode > gpurun_out/bench5_2gpu_$mode.json 2> gpurun_out/bench5_$mode.err
tail -3 gpurun_out/bench5_$mode.err
python - $mode <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench5_2gpu_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
print("kernel ms/step", d["timing"]["kernel_ms_per_step"])
print("parity", d.get("parity"))
print("e2e", d.get("e2e"))
PY
done
timeout 300 python bench.py --steps 200 --warmup 30 --no-cpu-baseline --no-spce > gpurun_out/bench5_1gpu.json 2> gpurun_out/bench5_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench5_1gpu.json").read().strip().splitlines()[-1])
print("1gpu bench: value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["e2e"]["value"], d["gpu_launches"]))
PY
