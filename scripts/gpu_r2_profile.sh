#!/bin/bash
# Profiles of the current build: ncu launch list of a bench step, ncu --set full of k_search (iterations 0, 2, 3, 5)
T=${1:-r2p}
mkdir -p gpurun_out
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --host-pack 1 > gpurun_out/${T}_ncu_bench.log 2>&1)
(timeout 200 python scripts/gpu_search_profile.py 64 2 2>&1 | tail -3) > gpurun_out/${T}_search.log 2>&1
N=$(grep -o "over [0-9]* launches" gpurun_out/${T}_search.log | head -1 | grep -o "[0-9]*")
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_search -s ${N:-10} -c 6 -f -o gpurun_out/${T}_search \
    python scripts/gpu_search_profile.py 64 2 > gpurun_out/${T}_ncu_search.log 2>&1)
cat gpurun_out/${T}_search.log; tail -2 gpurun_out/${T}_ncu_search.log; tail -2 gpurun_out/${T}_ncu_bench.log
