"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mulls_b200 import synth, abi
from mulls_b200.registration import Context as _Context
USE_GRAPH = int(os.environ.get("MULLS_SANITIZE_GRAPH", "1"))  # 0: the host launch loop instead of the iteration graph
def Context(*a, **k):
    c = _Context(*a, **k)
    c.set_tunable("use_graph", USE_GRAPH)
    return c
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
# the device-resident local map (update with dynamic removal and direction re-estimation, scan-to-map registration):
# the single-block ordered kernels of kernels_map.cuh
from mulls_b200.map_manager import LocalMap
seq = synth.make_sequence(77, 4, "small", n_points=6000)
ctx = Context(0, 1, 50000, 100000)
lm = LocalMap(ctx, 1 << 15)
motion = np.eye(4)
for k in range(4):
    sc = seq["scans"][k]
    if k > 0:
        r, _ = lm.icp_run(sc, seq["params"], motion)
        pose = np.asarray(r["T"], dtype=np.float64).reshape(4, 4)
    else:
        pose = np.eye(4)
    mp = abi.default_map_params()
    mp.local_map_radius, mp.max_num_pts, mp.kept_vertex_num = 60.0, 8000, 500
    mp.map_based_dynamic_removal_on = 1 if k >= 2 else 0
    mp.dynamic_removal_center_radius, mp.dynamic_dist_thre_min, mp.dynamic_dist_thre_max, mp.near_dist_thre = 15.0, 0.45, 1.5, 0.03
    mp.recalculate_feature_on = 1 if k == 3 else 0
    mp.random_seed = k
    info = lm.update(sc, pose, mp)
print("map", {k: v for k, v in lm.info().items() if k in ("feature_point_num",)} or "ok")
lm.close()
ctx.close()
