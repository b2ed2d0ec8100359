"""Developer script: resident throughput vs (pairs per step, lanes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mulls_b200.registration import PipelinedContext

allpairs = bench.make_pairs(bench.rank_seeds(0, 32), "c2")
keep = bench.pin_pairs(allpairs)
ms = max(sum(len(s) for s in p["src"]) for p in allpairs); mt = max(sum(len(t) for t in p["tgt"]) for p in allpairs)
for P, lanes in ((32, 8),):
    pairs = allpairs[:P]
    pc = PipelinedContext(0, lanes, (P + lanes - 1) // lanes, ms, mt)
    pc.upload(pairs)
    for _ in range(3): pc.run_resident()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(6): res = pc.run_resident()
    torch.cuda.synchronize(); t_res = (time.perf_counter() - t0) / 6
    for _ in range(2): pc.run_batch(pairs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): res2 = pc.run_batch(pairs)
    torch.cuda.synchronize(); t_e2e = (time.perf_counter() - t0) / 4
    print(f"pairs={P} lanes={lanes}: resident {t_res*1e3:.2f} ms/step = {P/t_res:.0f} reg/s; e2e {t_e2e*1e3:.2f} ms = {P/t_e2e:.0f} reg/s", flush=True)
    pc.close()

# native lanes (mulls_create_pipelined): one C-ABI call per step
from mulls_b200.registration import Context
for P, lanes in ((32, 8), (32, 4)):
    pairs = allpairs[:P]
    c = Context(0, P, ms, mt, lanes=lanes)
    c.upload(pairs)
    for _ in range(3): c.run_resident()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(6): c.run_resident()
    torch.cuda.synchronize(); t_res = (time.perf_counter() - t0) / 6
    for _ in range(2): c.run_batch(pairs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(6): c.run_batch(pairs)
    torch.cuda.synchronize(); t_e2e = (time.perf_counter() - t0) / 6
    print(f"NATIVE pairs={P} lanes={lanes}: resident {t_res*1e3:.2f} ms/step = {P/t_res:.0f} reg/s; e2e {t_e2e*1e3:.2f} ms = {P/t_e2e:.0f} reg/s", flush=True)
    c.close()
