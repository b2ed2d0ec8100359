#!/bin/bash
# single-call latency: loop kernel thresholds
T=${1:-r2c}
mkdir -p gpurun_out
(timeout 200 python scripts/gpu_latency.py 2>&1 | tail -10) > gpurun_out/${T}_latency.log 2>&1
cat gpurun_out/${T}_latency.log
