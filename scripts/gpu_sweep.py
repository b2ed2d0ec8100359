"""Developer script: sweep the search tunables on one pair, print device timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mulls_b200 import synth
from mulls_b200.registration import Context

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
pair = synth.make_pair(1000, cfg)
ctx = Context(0, 2, 700000, 700000)
ref = None
for h0 in (125, 250):
    for sl in (2, 4, 5, 6):
        for leaf in (16, 32, 64, 128):
            ctx.set_tunable("h0_min_mm", h0); ctx.set_tunable("start_level", sl); ctx.set_tunable("leaf_count", leaf)
            ctx.upload([pair])
            best = None
            for _ in range(3):
                res, _ = ctx.run_resident()
                st = ctx.stats()
                if best is None or st["ms_total"] < best["ms_total"]: best = st
            r = res[0]
            sig = (r["code"], r["iters"], tuple(r["n_corr"]), tuple(np.round(r["T"].ravel(), 12)))
            if ref is None: ref = sig
            it = " ".join(f"{v:.3f}" for v in best["ms_search_iter"][:r["iters"]])
            print(f"h0={h0} start={sl} leaf={leaf}: total {best['ms_total']:.3f} ingest {best['ms_ingest']:.3f} search {best['ms_search']:.3f} iterate {best['ms_iterate']:.3f} same={sig==ref} | {it}", flush=True)
