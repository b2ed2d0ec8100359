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
# the wire format packed on the host cores, and the per-frame front end (voxel filter -> ground filter -> classification)
ctx = Context(0, 2, 20000, 20000)
ctx.set_tunable("host_pack", 1)
res2, _ = ctx.run_batch([pair, pair])
print("icp packed", res2[0]["code"], res2[0]["iters"], bool(np.array_equal(res[0]["T"], res2[0]["T"])))
raw = np.concatenate(pair["tgt"], axis=0).copy()
raw[:, 3:8] = 0
gp = abi.default_ground_params()
gp.grid_resolution, gp.min_grid_pt_num = 2.0, 6
cp = abi.default_classify_params()
cp.neighbor_k, cp.pca_down_rate = 20, 1
ex = ctx.extract_semantic_pts(raw, 0.1, gp, cp)
print("extract", {k: int(v.shape[0]) for k, v in ex.items()})
gp.estimate_ground_normal_method, gp.fixed_num_downsampling, gp.down_ground_fixed_num = 0, 1, 50
g = ctx.fast_ground_filter(raw, gp)
print("ground", {k: int(v.shape[0]) for k, v in g.items()})
ctx.close()
