#!/bin/bash
# GPU call 29 of round 2 (1 GPU): GPU suite at the round's final commit
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests29.txt 2>&1; tail -3 gpurun_out/tests29.txt
