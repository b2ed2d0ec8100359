#!/bin/bash
# GPU tests, compute-sanitizer (memcheck + racecheck) on the small end-to-end run, glibc heap checking on a small bench,
# ncu launch list of two resident 64-pair runs of one context
T=${1:-r2q}
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/${T}_tests.log 2>&1
(echo "== memcheck"; timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_small.py 2>&1 | grep -v "^=========$" | tail -12;
 echo "== racecheck"; timeout 400 compute-sanitizer --tool racecheck python scripts/sanitize_small.py 2>&1 | grep -v "^=========$" | tail -12) > gpurun_out/${T}_sanitizer.log 2>&1
(MALLOC_CHECK_=3 timeout 300 python bench.py --pairs 16 --steps 3 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-400) > gpurun_out/${T}_malloc_check.log 2>&1
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python scripts/gpu_search_profile.py 64 2 > gpurun_out/${T}_ncu_launches.log 2>&1)
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_sanitizer.log gpurun_out/${T}_malloc_check.log; tail -2 gpurun_out/${T}_ncu_launches.log
