#!/bin/bash
# does the clock sampler cost the e2e leg something? the same bench with the three samplers, one box
T=${1:-r2s}
mkdir -p gpurun_out
for k in smi nvml off nvml smi; do
  (timeout 300 python bench.py --no-cpu-baseline --steps 12 --clock-sampler $k 2>> gpurun_out/${T}_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$k', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'rows48', round(d['e2e']['variants']['rows48']), 'link', round(d['e2e']['pinned_h2d_gbs'],1), d['clocks'])") >> gpurun_out/${T}_sampler_ab.log 2>&1
done
cat gpurun_out/${T}_sampler_ab.log; tail -3 gpurun_out/${T}_err.log
