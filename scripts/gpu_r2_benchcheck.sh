#!/bin/bash
# both bench arms of the committed bench.py
T=${1:-r2y}
mkdir -p gpurun_out
(timeout 400 python bench.py 2> gpurun_out/${T}_bench.err | tail -1) > gpurun_out/${T}_bench.json
(timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/${T}_ref.err | tail -1) > gpurun_out/${T}_bench_reference.json
cat gpurun_out/${T}_bench.json gpurun_out/${T}_bench_reference.json; tail -3 gpurun_out/${T}_bench.err gpurun_out/${T}_ref.err
