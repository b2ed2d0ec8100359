"""Generates tests/golden/frontend_demo.npz. Run from the repo root IN THE BUILD CONTAINER (needs /root/reference):

    python tests/golden/make_golden_frontend.py

Real data: every 3rd return of /root/reference/demo_data/pcd/000005.pcd (x y z intensity; the stored normals are wiped:
the front end computes its own), pushed through the ORACLE's chain of CFilter::extract_semantic_pts —
voxel_downsample (cfilter.hpp:83-165), fast_ground_filter (:1658-2036, RANSAC plane per cell) and classify_nground_pts
(:2058-2290) — with the parameters test/mulls_slam.cpp passes by default. For every output cloud the fixture stores the
row count and the SHA-256 of its bytes (the clouds themselves would be several MB), plus the ground cloud in full.
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
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
VOXEL = 0.05


def frontend_params():
    gp = abi.default_ground_params()
    cp = abi.default_classify_params()
    cp.neighbor_searching_radius, cp.neighbor_k, cp.neigh_k_min, cp.pca_down_rate = 1.0, 30, 8, 1
    cp.fixed_num_downsampling, cp.unground_down_fixed_num, cp.random_seed = 1, 10000, 11
    gp.random_seed = 11
    return gp, cp


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype=np.float32).tobytes()).digest(), dtype=np.uint8)


def oracle_chain(raw, gp, cp):
    down = oracle.voxel_downsample(raw, VOXEL)
    g = oracle.fast_ground_filter(down, gp)
    c = oracle.classify_nground(g["unground"], cp)
    out = {"down": down, "ground": g["ground"], "ground_down": g["ground_down"]}
    out.update({k: c[k] for k in abi.OUT_NAMES})
    return out


def main():
    scan = io.read_pcd("/root/reference/demo_data/pcd/000005.pcd")[::3]
    raw = np.zeros((scan.shape[0], 12), np.float32)
    raw[:, 0:3] = scan[:, 0:3]
    raw[:, 8] = scan[:, 8]
    gp, cp = frontend_params()
    out = oracle_chain(raw, gp, cp)
    d = {"xyzi": raw[:, [0, 1, 2, 8]].copy(), "exp_ground_rows": out["ground"]}
    for k, v in out.items():
        d["n_" + k] = np.int64(v.shape[0])
        d["sha_" + k] = digest(v)
    np.savez_compressed(os.path.join(HERE, "frontend_demo.npz"), **d)
    print({k: int(v.shape[0]) for k, v in out.items()})


if __name__ == "__main__":
    main()
