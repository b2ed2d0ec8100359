"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mulls_b200 import synth, abi
from mulls_b200.registration import Context
pair = synth.make_pair(1000, "small", n_points=6000)
ctx = Context(0, 2, 20000, 20000)
res, _ = ctx.run_batch([pair, pair])
print("icp", res[0]["code"], res[0]["iters"], res[0]["n_corr"])
cloud = np.concatenate([pair["tgt"][c] for c in (0, 2)], axis=0)[:4000]
out = ctx.pca_features(cloud, 0.6, 25, 2)
print("pca", int(out["pt_num"].sum()))
ctx.close()
