"""Developer script: throughput of a 16-pair batch with 1/2/4 concurrent contexts (lanes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mulls_b200.registration import PipelinedContext

P = 16
pairs = bench.make_pairs(bench.rank_seeds(0, P), "c2")
keep = bench.pin_pairs(pairs)
ms = max(sum(len(s) for s in p["src"]) for p in pairs); mt = max(sum(len(t) for t in p["tgt"]) for p in pairs)
for lanes in (1, 2, 4, 8):
    pc = PipelinedContext(0, lanes, (P + lanes - 1) // lanes, ms, mt)
    pc.upload(pairs)
    for _ in range(3): pc.run_resident()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): res = pc.run_resident()
    torch.cuda.synchronize(); t_res = (time.perf_counter() - t0) / 5
    for _ in range(2): pc.run_batch(pairs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): res2 = pc.run_batch(pairs)
    torch.cuda.synchronize(); t_e2e = (time.perf_counter() - t0) / 5
    ok = all(r["code"] == 1 for r in res) and all(r["code"] == 1 for r in res2)
    print(f"lanes={lanes}: resident {t_res*1e3:.2f} ms/step = {P/t_res:.0f} reg/s (wall); e2e {t_e2e*1e3:.2f} ms/step = {P/t_e2e:.0f} reg/s; ok={ok}", flush=True)
    pc.close()
# raw pinned H2D bandwidth for reference
import torch
x = torch.empty(184 * 1024 * 1024, dtype=torch.uint8).pin_memory(); y = torch.empty_like(x, device="cuda")
for _ in range(2): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"pinned H2D 184 MiB: {dt*1e3:.2f} ms = {x.numel()/dt/1e9:.1f} GB/s")
