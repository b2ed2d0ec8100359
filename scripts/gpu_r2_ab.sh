#!/bin/bash
# A/B of k_search tunables on one box + the GPU test-suite
T=${1:-r2c}
mkdir -p gpurun_out
(timeout 400 python scripts/gpu_search_ab.py 64 c2 "search_blocks=10" "search_blocks=12" "search_blocks=16" \
   "search_blocks=12,leaf_count=48" "search_blocks=12,leaf_count=64" "search_blocks=12,leaf_count=24" \
   "search_blocks=12,leaf_count=32,defer_from_iter=2" "search_blocks=12,defer_from_iter=99" "search_blocks=12,defer_from_iter=3,reseed_cells_x4=8" \
   "search_blocks=12,reseed_cells_x4=32" 2>&1 | tail -12) > gpurun_out/${T}_ab.log 2>&1
(timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/${T}_tests.log 2>&1
cat gpurun_out/${T}_ab.log; tail -3 gpurun_out/${T}_tests.log
