#!/bin/bash
# GPU call 19 of round 2 (2 GPUs): single-type LJ + coul_sf through the typed kernel (A/B against the generic kernel, 1 GPU each);
# GPU suite; 2-GPU weak line with 4 entries per thread in k_boost_owned; 1-GPU bench line (SPC/E timing from the resident loop)
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/tests19.txt 2>&1; tail -3 gpurun_out/tests19.txt
show() {
  python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench19_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {"value": 0, "ms_per_step": 0}
    print(sys.argv[1], "bench: N %d value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.3f ms) launches %d %s" % (d["config"]["atoms_total"], d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], e["value"], e["ms_per_step"], d["gpu_launches"], d["timing"]["comm_mode"]))
    print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()}, "parity", d.get("parity", {}).get("U_rel"), d.get("parity", {}).get("pairs_equal"), "U", d["state"]["U"], "builds", d["timing"]["list_builds_in_timed_region"])
    if "spce" in d: print("  spce", d["spce"]["value"], d["spce"]["ms_per_step"], d["spce"]["timing"], d["spce"]["e2e"]["value"], d["spce"]["roofline"]["frac"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 100 --warmup 20 --workload lj_coul_sf --atoms-per-gpu 1000000 --no-cpu-baseline --no-spce --no-e2e > gpurun_out/bench19_coul_typed.json 2> gpurun_out/bench19_coul_typed.err; show coul_typed
CUDA_VISIBLE_DEVICES=0 EMDEE_FORCE_VARIANT=30 timeout 300 python bench.py --steps 100 --warmup 20 --workload lj_coul_sf --atoms-per-gpu 1000000 --no-cpu-baseline --no-spce --no-e2e > gpurun_out/bench19_coul_generic.json 2> gpurun_out/bench19_coul_generic.err; show coul_generic
run() {
  tag=$1; np=$2; shift; shift
  EMDEE_BENCH_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np "$@" > gpurun_out/bench19_$tag.json 2> gpurun_out/bench19_$tag.err
  grep "bench trace r0" gpurun_out/bench19_$tag.err | cut -c1-400; tail -1 gpurun_out/bench19_$tag.err | cut -c1-300; show $tag
}
run 2gpu_weak 2 --steps 200 --warmup 30
run 2gpu_coul 2 --steps 60 --warmup 10 --workload lj_coul_sf --atoms-per-gpu 2000000 --no-e2e
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 200 --warmup 30 > gpurun_out/bench19_1gpu.json 2> gpurun_out/bench19_1gpu.err; show 1gpu
