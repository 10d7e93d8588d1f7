#!/bin/bash
# GPU call 15 of round 2 (1 GPU): GPU suite; force lab with the duo variants; bench (deferred kick, straight-line list build)
# with the kick A/B; SPC/E launch shapes; ncu of the build kernel and of the duo kernel
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests15.txt 2>&1; tail -3 gpurun_out/tests15.txt
timeout 300 python tools/force_lab.py --variants 0,20,21,22,23,24,25,0 > gpurun_out/lab15.txt 2>&1; cat gpurun_out/lab15.txt
timeout 300 python bench.py --steps 200 --warmup 30 > gpurun_out/bench15_1gpu.json 2> gpurun_out/bench15_1gpu.err
EMDEE_NO_DEFER_KICK=1 timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline --no-spce --no-parity --no-e2e > gpurun_out/bench15_nodefer.json 2> gpurun_out/bench15_nodefer.err
EMDEE_FORCE_VARIANT=20 timeout 200 python bench.py --steps 200 --warmup 30 --no-cpu-baseline --no-spce --no-e2e > gpurun_out/bench15_duo20.json 2> gpurun_out/bench15_duo20.err
python - <<'PY'
import json
for tag in ("1gpu", "nodefer", "duo20"):
    try:
        d = json.loads(open(f"gpurun_out/bench15_{tag}.json").read().strip().splitlines()[-1])
        print(tag, "value %.4e ms/step %.4f force_ms %.4f build_ms %.4f launches %d" % (d["value"], d["ms_per_step"], d["timing"]["force_kernel_ms"], d["timing"]["build_kernel_ms"], d["gpu_launches"]))
        print("  kernel ms/step", {k: round(v, 4) for k, v in d["timing"]["kernel_ms_per_step"].items()}, "parity", d.get("parity"))
        if "e2e" in d and d["e2e"]: print("  e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline_fp64"]["frac"])
        if "spce" in d: print("  spce", d["spce"]["value"], d["spce"]["ms_per_step"], d["spce"]["timing"])
    except Exception as e:
        print(tag, "FAILED", e)
PY
timeout 300 python tools/spce_lab.py --variants 0,1 > gpurun_out/spce_lab15.txt 2>&1; cat gpurun_out/spce_lab15.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_build_list -s 3 -c 1 -o gpurun_out/r2f_build python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2f_build.ncu-rep > gpurun_out/r2f_build.txt 2>&1; head -30 gpurun_out/r2f_build.txt
EMDEE_FORCE_VARIANT=20 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair_forces_duo -s 30 -c 1 -o gpurun_out/r2f_duo python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2f_duo.ncu-rep > gpurun_out/r2f_duo.txt 2>&1; cat gpurun_out/r2f_duo.txt
EMDEE_FORCE_VARIANT=20 timeout 300 ncu --set full --clock-control none -k regex:k_merge_duos -s 2 -c 1 -o gpurun_out/r2f_merge python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-spce --no-parity --no-e2e > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2f_merge.ncu-rep > gpurun_out/r2f_merge.txt 2>&1; head -12 gpurun_out/r2f_merge.txt
du -sh gpurun_out
