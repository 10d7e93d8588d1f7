#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root; ~30 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/next_gpu_call.sh'
# 1. gather-cost microbenchmark (decides between the layouts discussed in DESIGN.md section 5)
# 2. parity of the unconfirmed opt-in paths (EMDEE_ROWS, EMDEE_CLUSTER2), of the box-rescale scenario and of the
#    kernels written after the last GPU session (rigid bodies, verlet_step, bonded, Ewald, memory_address, sharing)
# 3. LJ-1M bench + one ncu capture each: default path vs EMDEE_ROWS=4/8/16/32 vs EMDEE_CLUSTER2=1 vs EMDEE_TEX=1/2 (texture-pipe gathers) vs EMDEE_REC16=1 (16-byte records; also with EMDEE_ROWS=4/8) vs EMDEE_TILESCHED=1; SPC/E line with and without EMDEE_TYPED=1
set -u
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/lsu_probe tools/lsu_probe.cu && timeout 120 /tmp/lsu_probe > gpurun_out/lsu_probe.txt 2>&1
EMDEE_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/experimental_tests.txt 2>&1
tail -30 gpurun_out/experimental_tests.txt
timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for g in 4 8 16 32; do
  EMDEE_ROWS=$g timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_rows$g.json 2> gpurun_out/bench_rows$g.err
done
EMDEE_CLUSTER2=1 timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_cluster2.json 2> gpurun_out/bench_cluster2.err
EMDEE_REC16=1 timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_rec16.json 2> gpurun_out/bench_rec16.err
for g in 4 8; do
  EMDEE_REC16=1 EMDEE_ROWS=$g timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_rows16_$g.json 2> gpurun_out/bench_rows16_$g.err
done
EMDEE_TILESCHED=1 timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_tilesched.json 2> gpurun_out/bench_tilesched.err
for m in 1 2; do
  EMDEE_TEX=$m timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline > gpurun_out/bench_tex$m.json 2> gpurun_out/bench_tex$m.err
done
# one full ncu capture of the force kernel per mapping worth comparing (wavefronts, L1 hit rate, pipe utilisation)
NCU="ncu --set full --clock-control none --import-source on -s 30 -c 1"
timeout 300 $NCU -k regex:k_pair_forces -o gpurun_out/r2a_force_default python bench.py --steps 30 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
EMDEE_ROWS=8 timeout 300 $NCU -k regex:k_pair_forces_rows -o gpurun_out/r2a_force_rows8 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
EMDEE_TEX=2 timeout 300 $NCU -k regex:k_pair_forces_tex -o gpurun_out/r2a_force_tex2 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
EMDEE_REC16=1 EMDEE_ROWS=8 timeout 300 $NCU -k regex:k_pair_forces_rows16 -o gpurun_out/r2a_force_rows16_8 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
EMDEE_REC16=1 timeout 300 $NCU -k regex:k_pair_forces_rec16 -o gpurun_out/r2a_force_rec16 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
EMDEE_TILESCHED=1 timeout 300 $NCU -k regex:k_pair_forces_sched -o gpurun_out/r2a_force_sched python bench.py --steps 30 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
for f in gpurun_out/r2a_force_*.ncu-rep; do python tools/ncu_summary.py "$f" > "${f%.ncu-rep}.txt" 2>&1; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "force_ms %.4f" % d["timing"]["force_kernel_ms"],
              "build_ms %.4f" % d["timing"]["build_kernel_ms"])
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
timeout 600 python bench.py --workload spce --steps 20 > gpurun_out/bench_spce.json 2> gpurun_out/bench_spce.err
tail -c 1500 gpurun_out/bench_spce.json
EMDEE_TYPED=1 timeout 600 python bench.py --workload spce --steps 20 > gpurun_out/bench_spce_typed.json 2> gpurun_out/bench_spce_typed.err
tail -c 1500 gpurun_out/bench_spce_typed.json
cat gpurun_out/lsu_probe.txt
# Multi-GPU follow-up (separate call, `gpurun --gpus 2|4|8`): python -m pytest tests/test_multi_gpu.py -m gpu -q ;
#   torchrun bench.py --gpus N (LJ, contract line) ; torchrun bench.py --workload spce --gpus N (rigid water, informational)
