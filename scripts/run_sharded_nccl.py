"""torchrun script: source-sharded registration over NCCL (BASELINE config 5 shape), checked against the
oracle and the unsharded single-GPU run on rank 0, with device timing.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/run_sharded_nccl.py [c5|c2|small]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from mulls_b200 import synth
from mulls_b200.dist import nccl_init_from_torch, shard_sources
from mulls_b200.registration import Context

cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pair = synth.make_pair(1000, cfg)
shards, base, glob = shard_sources(pair["src"], rank, world)
ctx = Context(local, 1, max(1, sum(len(s) for s in shards)), sum(len(t) for t in pair["tgt"]))
nccl_init_from_torch(ctx)  # the library's own communicator: the exchanges are ncclAllReduce calls inside the C++ loop
for _ in range(2):
    res, tr = ctx.run_sharded_nccl(dict(pair, src=shards), base, glob, want_trace=True)
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5
ms = 0.0
for _ in range(n):
    res, tr = ctx.run_sharded_nccl(dict(pair, src=shards), base, glob, want_trace=True)
    ms += ctx.stats()["ms_total"]
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n
t = torch.tensor([ms / n], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    from oracle import oracle

    o, ot = oracle.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"], threads=8)
    dt, dr = synth.pose_error(res["T"], o["T"])
    ok = (res["code"] == o["code"] and res["iters"] == o["iters"] and np.array_equal(tr["n_corr"], ot["n_corr"])
          and np.array_equal(tr["n_src"], ot["n_src"]) and dt <= 1e-4 and dr <= 1e-4)
    single = Context(local, 1, sum(len(s) for s in pair["src"]), sum(len(t) for t in pair["tgt"]))
    single.upload([pair])
    for _ in range(3):
        single.run_resident()
    s_ms = single.stats()["ms_total"]
    print(f"SHARDED {cfg} world={world}: parity_vs_oracle={ok} pose_diff=({dt:.2e} m, {dr:.2e} rad) iters={res['iters']} "
          f"device_ms_per_registration(max over ranks)={t.item():.3f} wall_ms={wall*1e3:.3f} single_gpu_resident_ms={s_ms:.3f}")
dist.barrier()
dist.destroy_process_group()
