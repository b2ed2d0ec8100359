#!/bin/bash
# ncu --set full of the k_search launches of ONE resident 64-pair run (iterations 0..5), plus the timing without ncu.
T=${1:-r2b}
mkdir -p gpurun_out
(timeout 200 python scripts/gpu_search_profile.py 64 3 2>&1 | tail -5) > gpurun_out/${T}_search.log 2>&1
N=$(grep -o "over [0-9]* launches" gpurun_out/${T}_search.log | head -1 | grep -o "[0-9]*")
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_search -s ${N:-11} -c 6 -f -o gpurun_out/${T}_search \
    python scripts/gpu_search_profile.py 64 2 > gpurun_out/${T}_ncu_search.log 2>&1)
cat gpurun_out/${T}_search.log; tail -3 gpurun_out/${T}_ncu_search.log
