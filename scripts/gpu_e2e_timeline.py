"""Where the wall time of the end-to-end leg goes: every lane's calls of mulls_icp_run_batch (pinned host clouds in,
results out) with the library's own host / device timings (mulls_run_stats.ms_host_*, ms_h2d, ms_total).
    python scripts/gpu_e2e_timeline.py [pairs=64] [lanes=8] [steps=12] [host_pack=1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from mulls_b200.registration import PipelinedContext

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
hp = int(sys.argv[4]) if len(sys.argv) > 4 else 1
numa = int(sys.argv[5]) if len(sys.argv) > 5 else 1
if numa:
    cpus = bench.gpu_numa_cpus(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
        print(f"affinity: {len(cpus)} cores of the GPU's NUMA node")
pairs = bench.make_pairs([1000 + i for i in range(n_pairs)], "c2")
keep = bench.pin_pairs(pairs)
ms = max(sum(len(s) for s in p["src"]) for p in pairs); mt = max(sum(len(t) for t in p["tgt"]) for p in pairs)
pipe = PipelinedContext(0, lanes, (n_pairs + lanes - 1) // lanes, ms, mt)
pipe.set_tunable("host_pack", hp)
parts = pipe._split(pairs)
pipe.run_batch_steps(pairs, 2)
T0 = time.perf_counter()

def work(c, p):
    rows = []
    for _ in range(steps):
        t0 = time.perf_counter() - T0
        c.run_batch(p)
        st = c.stats()
        rows.append((t0 * 1e3, (time.perf_counter() - T0) * 1e3, st["ms_host_call"], st["ms_host_upload"], st["ms_host_pack"], st["ms_h2d"], st["ms_total"]))
    return rows

futs = [pipe.pool.submit(work, c, p) for c, p in zip(pipe.lanes, parts)]
res = [f.result() for f in futs]
wall = time.perf_counter() - T0
print(f"pairs {n_pairs} lanes {lanes} steps {steps} host_pack {hp}: {n_pairs * steps / wall:.0f} reg/s, {wall * 1e3 / steps:.2f} ms per step")
print("lane: mean per call [python wall | lib call | upload host | pack wait | h2d device | compute device] ms")
for i, rows in enumerate(res):
    a = np.array(rows)
    print(f"  {i}: {np.mean(a[:, 1] - a[:, 0]):7.2f} | {a[:, 2].mean():7.2f} | {a[:, 3].mean():7.2f} | {a[:, 4].mean():7.2f} | {a[:, 5].mean():7.2f} | {a[:, 6].mean():7.2f}")
a = np.array(res[0])
print("lane 0 calls: start, end, lib, upload, pack, h2d, compute")
for r in a[:8]:
    print("   " + " ".join(f"{v:8.2f}" for v in r))
pipe.close()
