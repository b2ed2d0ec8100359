"""Odometry loop on the device-resident local map vs the CPU restatement.

Per frame, as test/mulls_slam.cpp:432-482 does: update_local_map(previous scan) -> mm_lls_icp(new scan -> map).
KITTI-urban-like settings (script/config/lo_gflag_list_kitti_urban.txt): radius 90 m, 20000 map points, dynamic removal on,
linear features recalculated every 10th frame. Prints per-stage medians; the GPU numbers are wall clock around the
C-ABI calls (host buffers in, results out), the CPU numbers are the oracle on this box's cores.
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mulls_b200 import abi, synth  # noqa: E402
from mulls_b200.map_manager import LocalMap  # noqa: E402
from mulls_b200.registration import Context  # noqa: E402
from oracle import oracle  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n_points = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
seq = synth.make_sequence(77, n_frames, "c2", n_points=n_points)
ctx = Context(0, 1, 200000, 400000)
lm = LocalMap(ctx, 1 << 17)
EMPTY = [np.zeros((0, 12), np.float32) for _ in range(6)]


def params(frame):
    p = abi.default_map_params()
    p.local_map_radius = 90.0
    p.max_num_pts = 20000
    p.kept_vertex_num = 2000
    p.map_based_dynamic_removal_on = 1 if frame >= 2 else 0
    p.dynamic_removal_center_radius = 15.0
    p.dynamic_dist_thre_min = 0.45
    p.dynamic_dist_thre_max = 1.5
    p.near_dist_thre = 0.03
    p.recalculate_feature_on = 1 if frame % 10 == 0 else 0
    p.random_seed = frame
    return p


icp = seq["params"]
t_up_g, t_up_dev, t_reg_g, t_reg_host, t_up_o, t_reg_o = [], [], [], [], [], []
omap, opose, oinfo = EMPTY, np.eye(4), None
pose_prev, motion, trees = np.eye(4), np.eye(4), None
for k in range(n_frames):
    sc = seq["scans"][k]
    if k == 0:
        pose = np.eye(4)
    else:
        t0 = time.perf_counter()
        r_g, _ = lm.icp_run(sc, icp, motion)
        t_reg_g.append(time.perf_counter() - t0)
        # the same registration with the map as a HOST target (what mulls_icp_run does without the resident map)
        host_map = lm.download()
        hp = abi.IcpParams.from_buffer_copy(icp)
        hp.target_bound[:] = list(lm.info()["local_bound"])
        pr = dict(tgt=host_map, src=sc, params=hp, init_guess=motion)
        t0 = time.perf_counter()
        (r_h,), _ = ctx.run_batch([pr])
        t_reg_host.append(time.perf_counter() - t0)
        assert np.array_equal(r_h["T"], r_g["T"])
        lm.icp_run(sc, icp, motion)  # restore block1's "trees" for the dynamic removal of the next update
        op = abi.IcpParams.from_buffer_copy(icp)
        op.target_bound[:] = list(oinfo["local_bound"])
        t0 = time.perf_counter()
        r_o, trees = oracle.icp_run_trees(omap, sc, op, motion)
        t_reg_o.append(time.perf_counter() - t0)
        assert r_o["code"] == r_g["code"] == 1 and np.allclose(r_o["T"], r_g["T"], atol=1e-9)
        motion = r_g["T"]
        pose = pose_prev @ r_g["T"]
    p = params(k)
    t0 = time.perf_counter()
    gi = lm.update(sc, pose, p)
    t_up_g.append(time.perf_counter() - t0)
    t_up_dev.append(gi["ms_update"])
    t0 = time.perf_counter()
    omap, oinfo = oracle.map_update(omap, opose, sc, pose, p, trees=trees if p.map_based_dynamic_removal_on else None)
    t_up_o.append(time.perf_counter() - t0)
    opose = oinfo["pose_lo"]
    assert np.array_equal(gi["n"], oinfo["n"]), (k, gi["n"], oinfo["n"])
    if p.recalculate_feature_on:  # PCA is tolerance-level: continue from identical maps
        omap = lm.download()
    pose_prev = pose
    print(f"frame {k}: map {gi['n'].tolist()} appended {gi['n_appended'].tolist()} recalc={p.recalculate_feature_on} "
          f"dyn={p.map_based_dynamic_removal_on} update gpu {t_up_g[-1]*1e3:.2f} ms (device {gi['ms_update']:.3f}) "
          f"cpu {t_up_o[-1]*1e3:.2f} ms", flush=True)

dt, dr = synth.pose_error(pose_prev, seq["poses"][-1])
med = lambda v: float(np.median(v)) * 1e3  # noqa: E731
print(f"scan features {sum(s.shape[0] for s in seq['scans'][0])} pts/frame, {n_frames} frames, final drift {dt:.4f} m {dr:.5f} rad")
print(f"update_local_map : GPU {med(t_up_g[1:]):.3f} ms wall ({float(np.median(t_up_dev[1:])):.3f} ms device) | "
      f"CPU restatement {med(t_up_o[1:]):.3f} ms")
print(f"scan-to-map ICP  : GPU resident map {med(t_reg_g):.3f} ms | GPU host map {med(t_reg_host):.3f} ms | "
      f"CPU restatement ({oracle.num_threads()} threads, reference-shaped) {med(t_reg_o):.3f} ms")
print(f"frame (update+ICP): GPU {med(t_up_g[1:]) + med(t_reg_g):.3f} ms | CPU {med(t_up_o[1:]) + med(t_reg_o):.3f} ms")
