#!/bin/bash
# GPU test-suite, e2e timeline with and without double buffering, bench
T=${1:-r2c}
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/${T}_tests.log 2>&1
(timeout 200 python scripts/gpu_e2e_timeline.py 64 8 12 1 2>&1 | tail -22) > gpurun_out/${T}_e2e_db1.log 2>&1
(timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/${T}_bench.err | tail -1) > gpurun_out/${T}_bench.json
tail -3 gpurun_out/${T}_tests.log; head -12 gpurun_out/${T}_e2e_db1.log; cat gpurun_out/${T}_bench.json
