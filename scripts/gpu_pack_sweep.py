"""Developer script: throughput of batched C2 registrations (32 pairs per step) vs the number of lanes and the pause
between two event polls of the launch loop (poll_pause), resident and end to end (host_pack 0 / 1)."""
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
for lanes in (8, 12, 16):
    pc = PipelinedContext(0, lanes, (P + lanes - 1) // lanes, ms, mt)
    for pause in (0, 64, 512):
        pc.set_tunable("poll_pause", pause)
        pc.set_tunable("host_pack", 2)
        pc.upload(pairs)
        for _ in range(3):
            pc.run_resident()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pc.run_resident_steps(STEPS)
        torch.cuda.synchronize(); t_res = (time.perf_counter() - t0) / STEPS
        out = [f"lanes={lanes} poll_pause={pause}: resident {P/t_res:.0f} reg/s"]
        for hp in (0, 1):
            pc.set_tunable("host_pack", hp)
            pc.run_batch_steps(pairs, 3)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = pc.run_batch_steps(pairs, STEPS)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / STEPS
            if ref is None:
                ref = res
            same = all(np.array_equal(a["T"], b["T"]) for a, b in zip(ref, res))
            out.append(f"e2e host_pack={hp} {P/dt:.0f} reg/s{'' if same else ' (RESULTS DIFFER)'}")
        print("; ".join(out), flush=True)
    pc.close()
