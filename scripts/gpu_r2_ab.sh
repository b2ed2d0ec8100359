#!/bin/bash
# GPU test-suite, per-iteration search times of one resident 64-pair batch, bench
T=${1:-r2c}
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/${T}_tests.log 2>&1
(timeout 400 python scripts/gpu_search_ab.py 64 c2 "defer_from_iter=3" 2>&1 | tail -4) > gpurun_out/${T}_ab.log 2>&1
(timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/${T}_bench.err | tail -1) > gpurun_out/${T}_bench.json
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_ab.log; cat gpurun_out/${T}_bench.json
