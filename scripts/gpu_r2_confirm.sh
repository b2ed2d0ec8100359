#!/bin/bash
# last confirmation of the committed tree: GPU tests, smoke, both bench arms
T=${1:-r2end}
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -4) > gpurun_out/${T}_tests.log 2>&1
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1) > gpurun_out/${T}_smoke.log 2>&1
(timeout 400 python bench.py 2> gpurun_out/${T}_bench.err | tail -1) > gpurun_out/${T}_bench.json
(timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/${T}_ref.err | tail -1) > gpurun_out/${T}_bench_reference.json
cat gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log; cut -c1-200 gpurun_out/${T}_bench.json; cut -c1-200 gpurun_out/${T}_bench_reference.json
