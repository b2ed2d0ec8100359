#!/bin/bash
# Round-2 state capture on the GPU box (one gpurun call): GPU tests, smoke, bench, ncu launch list of the bench step,
# ncu --set full of three k_search launches of a 64-pair resident run. Everything lands in gpurun_out/.
T=${1:-r2a}
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/${T}_tests.log 2>&1
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/${T}_smoke.log 2>&1
(timeout 300 python bench.py 2> gpurun_out/${T}_bench.err | tail -1) > gpurun_out/${T}_bench.json
(timeout 200 python scripts/gpu_search_profile.py 64 3 2>&1 | tail -5) > gpurun_out/${T}_search.log 2>&1
(timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --host-pack 1 > gpurun_out/${T}_ncu_bench.log 2>&1)
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_search -s 20 -c 4 -f -o gpurun_out/${T}_search \
    python scripts/gpu_search_profile.py 64 2 > gpurun_out/${T}_ncu_search.log 2>&1)
tail -4 gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log gpurun_out/${T}_search.log
cat gpurun_out/${T}_bench.json
