"""Generates tests/golden/demo_chain.npz — BASELINE config 1 (script/run_mulls_reg.sh on demo_data) as a fixture.
Run from the repo root IN THE BUILD CONTAINER (needs /root/reference):

    python tests/golden/make_golden_chain.py

Real data: the sixteen scans /root/reference/demo_data/pcd/000000.pcd .. 000015.pcd, every 2nd return, coordinates
rounded to the millimetre (stored as delta-coded int32 so that the file stays a few MB; the stored normals and labels
are wiped: the front end computes its own), intensity as stored. Every scan goes through the ORACLE's chain of
CFilter::extract_semantic_pts with the arguments test/mulls_reg.cpp:134-145 passes under script/run_mulls_reg.sh
(no voxel filter, gf grid 2.0 / 0.25 / 1.2, down rates 10 / 3, quadratic distance-inverse sampling, PCA r = 1.0 k = 50,
thresholds 0.65 / 0.65 / 0.10), then the fifteen consecutive pairs k -> k+1 and the script's own pair 000000 <-> 000015
are registered the way test/mulls_reg.cpp:164-195 does: determine_source_target_cloud (the block with more down-sampled
feature points is the target, cregistration.hpp:857-870), identity initial guess, mm_lls_icp(reg_con, 10, 3.0, 0.001,
0.01, 0.75, 1.1, "111110", "1101", 1.0, 0.1, 0.1, 0.1) on target = pc_* / source = pc_*_down (:1180-1181).
Stored per scan: row count and SHA-256 of the thirteen output clouds; per pair: which block became the target, the
oracle's Trans1_2, code, iterations and per-iteration correspondence / source counts.
The reference holds no expected outputs for this path: the fixture pins the CUDA path (and future oracle edits) to the
oracle as committed; parity with the reference binary stays UNPINNED.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mulls_b200 import abi, io  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N_SCANS = 16
STEP = 2
PAIRS = [(k, k + 1) for k in range(N_SCANS - 1)] + [(0, N_SCANS - 1)]
CLOUDS = ("down", "ground", "ground_down") + tuple(abi.OUT_NAMES)
# class order of the registration: ground, pillar, facade, beam, roof, vertex
TGT_KEYS = ("ground", "pillar", "facade", "beam", "roof", "vertex")
SRC_KEYS = ("ground_down", "pillar_down", "facade_down", "beam_down", "roof_down", "vertex")


def chain_params():
    gp = abi.default_ground_params()
    gp.min_grid_pt_num, gp.grid_resolution, gp.max_height_difference, gp.neighbor_height_diff = 8, 2.0, 0.25, 1.2
    gp.max_ground_height = float("inf")  # (float)DBL_MAX, test/mulls_reg.cpp:33,84
    gp.ground_random_down_rate, gp.ground_random_down_down_rate, gp.nonground_random_down_rate = 10, 2, 3
    gp.reliable_neighbor_grid_num_thre, gp.estimate_ground_normal_method, gp.normal_estimation_radius = 0, 3, 2.0
    gp.distance_weight_downsampling_method, gp.standard_distance = 2, 15.0
    gp.fixed_num_downsampling, gp.down_ground_fixed_num = 0, 500
    gp.random_seed = 7
    cp = abi.default_classify_params()
    cp.neighbor_searching_radius, cp.neighbor_k, cp.neigh_k_min, cp.pca_down_rate = 1.0, 50, 8, 1
    cp.edge_thre, cp.planar_thre, cp.edge_thre_down, cp.planar_thre_down = 0.65, 0.65, 0.75, 0.75
    cp.extract_vertex_points_method, cp.curvature_thre, cp.vertex_curvature_non_max_radius = 2, 0.10, 1.5
    cp.fixed_num_downsampling = 0
    cp.pillar_down_fixed_num, cp.facade_down_fixed_num, cp.beam_down_fixed_num, cp.roof_down_fixed_num = 200, 800, 200, 200
    cp.unground_down_fixed_num = 20000
    cp.roof_height_min = 0.0
    cp.random_seed = 7
    return gp, cp


def icp_params(bound):
    p = abi.default_params()
    p.max_iter_num, p.dis_thre_unit, p.converge_translation, p.converge_rotation_d = 10, 3.0, 0.001, 0.01
    p.dis_thre_min, p.dis_thre_update_rate = 0.75, 1.1
    p.used_feature_type, p.weight_strategy = b"111110", b"1101"
    p.z_xy_balanced_ratio, p.pt2pt_residual_window, p.pt2pl_residual_window, p.pt2li_residual_window = 1.0, 0.1, 0.1, 0.1
    p.target_bound[:] = bound
    return p


def encode_scan(xyz_mm):
    """int32 millimetres -> first row + row deltas (small numbers compress well)"""
    d = np.diff(xyz_mm, axis=0, prepend=np.zeros((1, 3), np.int32))
    return d.astype(np.int32)


def decode_scan(delta, intensity):
    mm = np.cumsum(delta.astype(np.int64), axis=0)
    raw = np.zeros((mm.shape[0], 12), np.float32)
    raw[:, 0:3] = (mm.astype(np.float64) / 1000.0).astype(np.float32)
    raw[:, 8] = intensity
    return raw


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype=np.float32).tobytes()).digest(), dtype=np.uint8)


def oracle_features(raw, gp, cp):
    from oracle import oracle

    down = oracle.voxel_downsample(raw, 0.0)
    g = oracle.fast_ground_filter(down, gp)
    c = oracle.classify_nground(g["unground"], cp)
    out = {"down": down, "ground": g["ground"], "ground_down": g["ground_down"]}
    out.update({k: c[k] for k in abi.OUT_NAMES})
    return out


def down_feature_point_num(f):  # cfilter.hpp:2401-2402
    return sum(len(f[k]) for k in ("ground_down", "pillar_down", "beam_down", "facade_down", "roof_down", "vertex"))


def make_pair(feats, raws, a, b):
    """(target block index, source block index, pair dict) as determine_source_target_cloud + mm_lls_icp see them"""
    t, s = (a, b) if down_feature_point_num(feats[a]) > down_feature_point_num(feats[b]) else (b, a)
    x = raws[t][:, :3].astype(np.float64)
    bound = [x[:, 0].min(), x[:, 1].min(), x[:, 2].min(), x[:, 0].max(), x[:, 1].max(), x[:, 2].max()]  # dataio.hpp:1736
    pair = {"tgt": [abi.as_aos48(feats[t][k]) for k in TGT_KEYS], "src": [abi.as_aos48(feats[s][k]) for k in SRC_KEYS],
            "params": icp_params(bound), "init_guess": np.eye(4)}
    return t, s, pair


def main():
    from oracle import oracle

    gp, cp = chain_params()
    d, raws, feats = {}, [], []
    for k in range(N_SCANS):
        scan = io.read_pcd(f"/root/reference/demo_data/pcd/{k:06d}.pcd")[::STEP]
        mm = np.round(scan[:, 0:3].astype(np.float64) * 1000.0).astype(np.int32)
        d[f"scan{k}_dmm"] = encode_scan(mm)
        d[f"scan{k}_i"] = scan[:, 8].astype(np.float32)
        raw = decode_scan(d[f"scan{k}_dmm"], d[f"scan{k}_i"])
        raws.append(raw)
        f = oracle_features(raw, gp, cp)
        feats.append(f)
        d[f"scan{k}_n"] = np.array([len(f[c]) for c in CLOUDS], np.int64)
        d[f"scan{k}_sha"] = np.stack([digest(f[c]) for c in CLOUDS])
        print(k, "raw", len(raw), {c: len(f[c]) for c in CLOUDS})
    for i, (a, b) in enumerate(PAIRS):
        t, s, pair = make_pair(feats, raws, a, b)
        res, tr = oracle.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"], threads=0)
        d[f"pair{i}_ts"] = np.array([t, s], np.int32)
        d[f"pair{i}_T"] = res["T"]
        d[f"pair{i}_code_iters"] = np.array([res["code"], res["iters"]], np.int32)
        d[f"pair{i}_sigma"] = np.float32(res["sigma"])
        d[f"pair{i}_trace_n_corr"] = tr["n_corr"][: tr["n_iter"]]
        d[f"pair{i}_trace_n_src"] = tr["n_src"][: tr["n_iter"]]
        print("pair", a, b, "target", t, "code", res["code"], "iters", res["iters"], "t =", np.round(res["T"][:3, 3], 3),
              "n_corr", res["n_corr"])
    path = os.path.join(HERE, "demo_chain.npz")
    np.savez_compressed(path, **d)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
