"""Developer script: throughput of batched C2 registrations vs (pairs per step, lanes), resident and end to end
(host_pack 0 / 1); plus the single-thread repacking rate of this host."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from mulls_b200 import abi
from mulls_b200.registration import PipelinedContext

STEPS = int(os.environ.get("SWEEP_STEPS", "12"))
lib = abi.load_library()
n = 2_000_000
rows = np.random.rand(n, 12).astype(np.float32)
outb = torch.empty(7 * n + 8, dtype=torch.float32).pin_memory().numpy()
off = (-outb.ctypes.data % 16) // 4
fp = C.POINTER(C.c_float)
for _ in range(3):
    t0 = time.perf_counter(); lib.mulls_pack_rows(rows.ctypes.data_as(fp), n, 1, outb[off:].ctypes.data_as(fp)); dt = time.perf_counter() - t0
print(f"host cores {os.cpu_count()}; mulls_pack_rows on one thread: {dt/n*1e9:.2f} ns/point ({48*n/dt/1e9:.1f} GB/s read)", flush=True)
allpairs = bench.make_pairs(bench.rank_seeds(0, int(os.environ.get("SWEEP_MAX_PAIRS", "64"))), "c2")
keep = bench.pin_pairs(allpairs)
ms = max(sum(len(s) for s in p["src"]) for p in allpairs); mt = max(sum(len(t) for t in p["tgt"]) for p in allpairs)
CONFIGS = [tuple(int(v) for v in c.split("x")) for c in os.environ.get("SWEEP_CONFIGS", "32x8,48x12,64x16,64x8,32x8").split(",")]
for P, lanes in CONFIGS:
    pairs = allpairs[:P]
    pc = PipelinedContext(0, lanes, (P + lanes - 1) // lanes, ms, mt)
    pc.upload(pairs)
    for _ in range(3):
        pc.run_resident()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pc.run_resident_steps(STEPS)
    torch.cuda.synchronize(); t_res = (time.perf_counter() - t0) / STEPS
    out = [f"pairs={P} lanes={lanes}: resident {P/t_res:.0f} reg/s"]
    for hp in (0, 1):
        pc.set_tunable("host_pack", hp)
        pc.run_batch_steps(pairs, 3)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pc.run_batch_steps(pairs, STEPS)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / STEPS
        out.append(f"e2e host_pack={hp} {P/dt:.0f} reg/s")
    print("; ".join(out), flush=True)
    pc.close()
