"""Developer script: end-to-end rate of batched C2 registrations vs how the clouds cross PCIe — the caller's 48-byte
rows (host_pack 0) or the 28 B/point wire format packed on the host cores (host_pack 1) with 4..12 pack workers, with
plain or write-combined pinned staging. Run with MULLS_PACK_THREADS=4 so that the pool starts small (it only grows)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from mulls_b200.registration import PipelinedContext

P, STEPS = 32, int(os.environ.get("SWEEP_STEPS", "20"))
pairs = bench.make_pairs(bench.rank_seeds(0, P), "c2")
keep = bench.pin_pairs(pairs)
ms = max(sum(len(s) for s in p["src"]) for p in pairs); mt = max(sum(len(t) for t in p["tgt"]) for p in pairs)
print(f"host cores {os.cpu_count()}, {STEPS} steps per point", flush=True)
ref = None
lanes = 8
pc = PipelinedContext(0, lanes, (P + lanes - 1) // lanes, ms, mt)
for hp, th, wc in ((0, 0, 0), (1, 4, 0), (1, 4, 1), (1, 6, 0), (1, 6, 1), (1, 8, 0), (1, 8, 1), (1, 12, 0), (1, 12, 1), (0, 0, 0)):
    pc.set_tunable("host_pack", hp)
    pc.set_tunable("stage_wc", wc)
    if th:
        pc.set_tunable("pack_threads", th)
    pc.run_batch_steps(pairs, 3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = pc.run_batch_steps(pairs, STEPS)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / STEPS
    if ref is None:
        ref = res
    same = all(np.array_equal(a["T"], b["T"]) for a, b in zip(ref, res))
    print(f"lanes={lanes} host_pack={hp} pack_threads={th} wc={wc}: e2e {dt*1e3:.2f} ms/step = {P/dt:.0f} reg/s, identical={same}", flush=True)
pc.close()
