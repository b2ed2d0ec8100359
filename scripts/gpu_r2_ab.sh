#!/bin/bash
# GPU test-suite + single-call latency (small registrations run the whole loop in one cooperative kernel)
T=${1:-r2c}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/${T}_tests.log 2>&1
(timeout 120 python scripts/gpu_latency.py 2>&1 | tail -5) > gpurun_out/${T}_latency.log 2>&1
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2) > gpurun_out/${T}_smoke.log 2>&1
tail -4 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_latency.log gpurun_out/${T}_smoke.log
