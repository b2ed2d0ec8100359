"""Developer script: 16-pair resident batch, sweep of search tunables; prints device timings per setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from mulls_b200.registration import Context

P = int(os.environ.get("PAIRS", "16"))
pairs = bench.make_pairs(bench.rank_seeds(0, P), "c2")
ms = max(sum(len(s) for s in p["src"]) for p in pairs); mt = max(sum(len(t) for t in p["tgt"]) for p in pairs)
ctx = Context(0, P, ms, mt)
ctx.upload(pairs)
ref = None
settings = [dict(packet_max_ext_mm=e, leaf_count=l) for l in (32, 16) for e in (0, 500, 1000, 2000, 4000)]

for st in settings:
    for k, v in st.items(): ctx.set_tunable(k, v)
    best = None
    for _ in range(4):
        res, _ = ctx.run_resident()
        s = ctx.stats()
        if best is None or s["ms_total"] < best["ms_total"]: best = s
    sig = tuple((r["code"], r["iters"], tuple(r["n_corr"]), r["T"].tobytes()) for r in res)
    if ref is None: ref = sig
    it = " ".join(f"{v:.3f}" for v in best["ms_search_iter"][:8])
    print(f"{st}: total {best['ms_total']:.3f} ingest {best['ms_ingest']:.3f} search {best['ms_search']:.3f} same={sig==ref} | {it}", flush=True)
