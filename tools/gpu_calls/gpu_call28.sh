#!/bin/bash
# GPU call 28 of round 2 (1 GPU): device-math probe test + GPU suite with the final typed kernel; SPC/E kernel time
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/tests28.txt 2>&1; tail -3 gpurun_out/tests28.txt
timeout 200 python tools/spce_lab.py --variants 0 --evals 4 > gpurun_out/spce_lab28.txt 2>&1; cat gpurun_out/spce_lab28.txt
