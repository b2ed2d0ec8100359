#!/bin/bash
# Round-2 final sequence on the GPU box (one gpurun call): GPU tests, smoke, both bench arms, ncu launch list of two resident
# 64-pair runs (one context, host launch loop), ncu --set full of the search kernels, sanitizer. Everything lands in gpurun_out/.
T=${1:-r2z}
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > gpurun_out/${T}_tests.log 2>&1
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/${T}_smoke.log 2>&1
(timeout 400 python bench.py 2> gpurun_out/${T}_bench.err | tail -1) > gpurun_out/${T}_bench.json
(timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/${T}_ref.err | tail -1) > gpurun_out/${T}_bench_reference.json
(timeout 200 python scripts/gpu_search_profile.py 64 3 2>&1 | tail -4) > gpurun_out/${T}_search.log 2>&1
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python scripts/gpu_search_profile.py 64 2 > gpurun_out/${T}_ncu_launches.log 2>&1)
N=$(grep -o "over [0-9]* launches" gpurun_out/${T}_search.log | head -1 | grep -o "[0-9]*")
(timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_search|k_accumulate" -s $(( 2 * ${N:-10} )) -c 12 -f -o gpurun_out/${T}_search \
    python scripts/gpu_search_profile.py 64 2 > gpurun_out/${T}_ncu_search.log 2>&1)
(echo "== memcheck (iteration graph)"; timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_small.py 2>&1 | grep -v "^=========$" | tail -9;
 echo "== racecheck (host launch loop)"; MULLS_SANITIZE_GRAPH=0 timeout 400 compute-sanitizer --tool racecheck python scripts/sanitize_small.py 2>&1 | grep -v "^=========$" | tail -9) > gpurun_out/${T}_sanitizer.log 2>&1
(timeout 120 python scripts/gpu_latency.py 2>&1 | tail -5) > gpurun_out/${T}_latency.log 2>&1
tail -3 gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log gpurun_out/${T}_sanitizer.log gpurun_out/${T}_search.log
cat gpurun_out/${T}_bench.json gpurun_out/${T}_bench_reference.json
