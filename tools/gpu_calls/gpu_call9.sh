#!/bin/bash
# GPU call 9 of round 2 (2 GPUs): race fix + compact rebuild path: multi-GPU parity x3, 2-GPU benches; SPC/E typed variants;
# ncu captures of the shipped force and build kernels + launch list (1 GPU)
set -u
mkdir -p gpurun_out
for rep in 1 2 3; do
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/tests9_$rep.txt 2>&1
tail -2 gpurun_out/tests9_$rep.txt
done
show() {
  python - $1 <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench9_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {"value": 0, "ms_per_step": 0}
    print(sys.argv[1], "bench: N %d value %.4e ms/step %.4f force_ms %.4f build_ms %.4f e2e %.4e (%.3f ms) launches %d %s" % (d["config"]["atoms_total"], d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], e["value"], e["ms_per_step"], d["gpu_launches"], d["timing"]["comm_mode"]))
    print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()}, "parity", d.get("parity", {}).get("U_rel"), d.get("parity", {}).get("pairs_equal"), "U", d["state"]["U"], "builds", d["timing"]["list_builds_in_timed_region"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run() {
  tag=$1; np=$2; shift; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np "$@" > gpurun_out/bench9_$tag.json 2> gpurun_out/bench9_$tag.err
  tail -2 gpurun_out/bench9_$tag.err | cut -c1-400; show $tag
}
run 2gpu_weak 2 --steps 200 --warmup 30
EMDEE_NO_PEER=1 run 2gpu_weak_nccl 2 --steps 200 --warmup 30 --no-e2e
timeout 300 python tools/spce_lab.py > gpurun_out/spce_lab9.txt 2>&1; cat gpurun_out/spce_lab9.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pair_forces -s 30 -c 1 -o gpurun_out/r2d_force python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2d_force.ncu-rep > gpurun_out/r2d_force.txt 2>&1; cat gpurun_out/r2d_force.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_build_list -s 3 -c 1 -o gpurun_out/r2d_build python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2d_build.ncu-rep > gpurun_out/r2d_build.txt 2>&1; head -12 gpurun_out/r2d_build.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
ls -la gpurun_out | tail -8
for kn in k_boost k_displace k_refresh_positions; do
timeout 300 ncu --set full --clock-control none -k regex:$kn -s 20 -c 1 -o gpurun_out/r2d_$kn python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2d_$kn.ncu-rep > gpurun_out/r2d_$kn.txt 2>&1; head -22 gpurun_out/r2d_$kn.txt
done
