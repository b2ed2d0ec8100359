#!/bin/bash
# e2e timeline per lane, single-call latency, ncu --set full of the non-search kernels of one resident 64-pair run
T=${1:-r2d}
mkdir -p gpurun_out
(timeout 200 python scripts/gpu_e2e_timeline.py 64 8 12 1 2>&1 | tail -24) > gpurun_out/${T}_e2e_pack1.log 2>&1
(timeout 200 python scripts/gpu_e2e_timeline.py 64 8 12 0 2>&1 | tail -14) > gpurun_out/${T}_e2e_pack0.log 2>&1
(timeout 200 python scripts/gpu_e2e_timeline.py 64 8 12 1 0 2>&1 | tail -14) > gpurun_out/${T}_e2e_pack1_nonuma.log 2>&1
(timeout 200 python scripts/gpu_latency.py 2>&1 | tail -6) > gpurun_out/${T}_latency.log 2>&1
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_accumulate|k_ingest_transform|k_hash_build|k_resolve|k_gather|k_make_keys" -s 20 -c 14 -f -o gpurun_out/${T}_others \
    python scripts/gpu_search_profile.py 64 2 > gpurun_out/${T}_ncu_others.log 2>&1)
cat gpurun_out/${T}_e2e_pack1.log gpurun_out/${T}_e2e_pack0.log gpurun_out/${T}_e2e_pack1_nonuma.log gpurun_out/${T}_latency.log; tail -3 gpurun_out/${T}_ncu_others.log
