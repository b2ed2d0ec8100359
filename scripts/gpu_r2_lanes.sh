#!/bin/bash
# e2e leg: how many lanes with double-buffered calls?
T=${1:-r2lanes}
mkdir -p gpurun_out
for l in 8 4 6 12 8; do (timeout 100 python scripts/gpu_e2e_timeline.py 64 $l 12 1 2>&1 | grep "reg/s") >> gpurun_out/${T}.log 2>&1; done
cat gpurun_out/${T}.log
