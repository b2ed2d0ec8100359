#!/bin/bash
# Round-end measurement sequence on the GPU box (one gpurun call): GPU tests, smoke, bench (both arms), ncu launch
# lists of the bench step and of the per-frame front end, sanitizer, throughput sweep. Everything lands in gpurun_out/.
T=${1:-r1z}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > gpurun_out/${T}_tests.log 2>&1
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3) > gpurun_out/${T}_smoke.log 2>&1
(timeout 240 python bench.py 2> gpurun_out/${T}_bench.err | tail -1) > gpurun_out/${T}_bench.json
(timeout 120 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/${T}_ref.err | tail -1) > gpurun_out/${T}_bench_reference.json
(timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${T}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --host-pack 1 > gpurun_out/${T}_ncu_bench.log 2>&1)
(timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}_extract.csv \
    python scripts/gpu_extract_bench.py --profile > gpurun_out/${T}_ncu_extract.log 2>&1)
(echo "== memcheck"; timeout 240 compute-sanitizer --tool memcheck python scripts/sanitize_small.py 2>&1 | grep -v "^=========$" | tail -8) > gpurun_out/${T}_sanitizer.log 2>&1
(timeout 120 python scripts/gpu_extract_bench.py 2>&1 | tail -8) > gpurun_out/${T}_extract.log 2>&1
(SWEEP_STEPS=10 timeout 200 python scripts/gpu_pack_sweep.py 2>&1 | tail -12) > gpurun_out/${T}_sweep.log 2>&1
tail -3 gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log gpurun_out/${T}_sanitizer.log gpurun_out/${T}_sweep.log
cat gpurun_out/${T}_bench.json gpurun_out/${T}_bench_reference.json
