#!/bin/bash
# GPU call 25 of round 2 (1 GPU): launch shapes / index prefetch of the compact-record pair kernel
set -u
mkdir -p gpurun_out
timeout 400 python tools/force_lab.py --variants 0,54,51,53,56,57,58,59,54,0 --steps 40 > gpurun_out/lab25.txt 2>&1; cat gpurun_out/lab25.txt
