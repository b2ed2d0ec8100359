#!/bin/bash
# A/B of k_search tunables on one box + bench
T=${1:-r2c}
mkdir -p gpurun_out
(timeout 400 python scripts/gpu_search_ab.py 64 c2 "search_blocks=16" "search_blocks=12" "search_blocks=10" 2>&1 | tail -8) > gpurun_out/${T}_ab.log 2>&1
(timeout 400 python bench.py 2> gpurun_out/${T}_bench.err | tail -1) > gpurun_out/${T}_bench.json
(timeout 200 python scripts/gpu_latency.py 2>&1 | tail -6) > gpurun_out/${T}_latency.log 2>&1
cat gpurun_out/${T}_ab.log; cat gpurun_out/${T}_bench.json; cat gpurun_out/${T}_latency.log
