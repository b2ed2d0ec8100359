"""Non-ground feature classification: lo::CFilter::classify_nground_pts,
include/common/cfilter.hpp:2058-2290 (PCA -> thresholds -> neighbourhood promotion -> keypoint encoding -> NMS ->
fixed-number down-sampling).

CPU part: the restatement in oracle/ against what the reference code states. GPU part: mulls_classify_nground against
the restatement, every output cloud compared as bit patterns (the PCA of this path accumulates pcl::PCA's float mean /
covariance in radiusSearch order on both sides, so even the floating-point stage is bit-reproducible).
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from mulls_b200 import abi, synth
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def unground_cloud(seed=5, config="c2", n_keep=None):
    """A raw-looking non-ground cloud: every non-ground return of a synthetic sweep, shuffled, normals wiped."""
    pr = synth.make_pair(seed, config)
    ung = np.concatenate([pr["tgt"][c] for c in (abi.PILLAR, abi.FACADE, abi.BEAM, abi.ROOF)], axis=0).copy()
    np.random.default_rng(seed).shuffle(ung)
    ung[:, 4:8] = 0
    ung[:, 9] = np.linspace(0, 1, ung.shape[0], dtype=np.float32)  # "timestamp" in curvature
    return np.ascontiguousarray(ung[:n_keep] if n_keep else ung)


def kitti_params(**kw):
    """script/config/lo_gflag_list_kitti_urban.txt values"""
    p = abi.default_classify_params()
    p.neighbor_searching_radius = 0.7
    p.neighbor_k = 25
    p.neigh_k_min = 7
    p.pca_down_rate = 2
    p.edge_thre = 0.62
    p.planar_thre = 0.62
    p.curvature_thre = 0.08
    p.fixed_num_downsampling = 1
    p.pillar_down_fixed_num = 400
    p.facade_down_fixed_num = 1200
    p.beam_down_fixed_num = 200
    p.roof_down_fixed_num = 0
    p.unground_down_fixed_num = 20000
    p.random_seed = 3
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_classify_structs_match_header():
    code = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "mulls_b200/abi.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu\n", sizeof(mulls_classify_params), sizeof(mulls_classify_out),
             offsetof(mulls_classify_params, beam_height_max), offsetof(mulls_classify_params, random_seed),
             offsetof(mulls_classify_out, n));
      return 0; }"""
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(code)
        exe = os.path.join(td, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = [int(v) for v in subprocess.check_output([exe]).decode().split()]
    assert out == [C.sizeof(abi.ClassifyParams), C.sizeof(abi.ClassifyOut), abi.ClassifyParams.beam_height_max.offset,
                   abi.ClassifyParams.random_seed.offset, abi.ClassifyOut.n.offset]
    lib = abi.load_library()
    p = abi.ClassifyParams()
    lib.mulls_classify_default_params(C.byref(p))
    q = abi.default_classify_params()
    for name, _ in abi.ClassifyParams._fields_:
        assert getattr(p, name) == getattr(q, name), name


def test_oracle_classification_follows_the_scene():
    ung = unground_cloud()
    p = kitti_params()
    o = oracle.classify_nground(ung, p)
    assert o["unground"].shape[0] == 20000  # :2086-2087
    # directions / normals re-estimated from the geometry land in the right class (cfilter.hpp:2103-2166)
    assert o["pillar"].shape[0] > 50 and (np.abs(o["pillar"][:, 6]) > 0.94).all()
    assert o["beam"].shape[0] > 50 and (np.abs(o["beam"][:, 6]) < 0.17).all()
    assert o["facade"].shape[0] > 500 and (np.abs(o["facade"][:, 6]) < 0.34).all()
    assert (np.abs(o["roof"][:, 6]) > 0.98).all()
    # class clouds are sorted by the NMS score, descending (:1252)
    for k in ("pillar", "beam", "facade"):
        if o[k].shape[0] >= 10:
            assert (np.diff(o[k][:, 7]) <= 0).all(), k
    # fixed numbers (:2257-2267): pillar <= 400, facade <= 4 sectors x 300, beam <= 4 x 50, roof cleared
    assert o["pillar_down"].shape[0] <= 400 and o["facade_down"].shape[0] <= 1200
    assert o["beam_down"].shape[0] <= 200 and o["roof_down"].shape[0] == 0
    # keypoints carry the neighbourhood descriptor in curvature / normal_x / normal_y (:1150-1158)
    v = o["vertex"]
    assert v.shape[0] > 0 and (v[:, 9] == np.floor(v[:, 9])).all() and (v[:, 9] >= 0).all() and (v[:, 9] <= 100000000).all()


def test_oracle_nms_is_the_greedy_selection():
    ung = unground_cloud(n_keep=6000)
    p = kitti_params(fixed_num_downsampling=0, pca_down_rate=1)
    o = oracle.classify_nground(ung, p)
    r2 = np.float32(np.float64(np.float32(0.25 * np.float32(0.7))) ** 2)
    checked = 0
    for k in ("facade", "beam", "pillar"):
        pts, down = o[k], o[k + "_down"]
        if pts.shape[0] < 10:
            continue
        kept = []
        alive = np.ones(pts.shape[0], bool)
        for i in range(pts.shape[0]):
            if not alive[i]:
                continue
            kept.append(i)
            d = pts[:, :3] - pts[i, :3]
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            alive &= ~(d2 < r2)
        assert np.array_equal(pts[kept], down), k
        checked += 1
    assert checked >= 1


def _assert_same(g, o, tag):
    for k in abi.OUT_NAMES:
        assert g[k].shape == o[k].shape, f"{tag}: {k} {g[k].shape} vs {o[k].shape}"
        assert np.array_equal(g[k].view(np.uint32), o[k].view(np.uint32)), f"{tag}: {k} differs"


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["kitti", "no_nms", "no_vertex", "dense_stride1", "small"])
def test_gpu_classify_matches_oracle(variant):
    from mulls_b200.registration import Context

    ung = unground_cloud()
    if variant == "kitti":
        p = kitti_params()
    elif variant == "no_nms":
        p = kitti_params(sharpen_with_nms=0)
    elif variant == "no_vertex":
        p = kitti_params(curvature_thre=0.0, fixed_num_downsampling=0)
    elif variant == "dense_stride1":
        p = kitti_params(pca_down_rate=1, neighbor_searching_radius=1.0, neighbor_k=50, neigh_k_min=8,
                         unground_down_fixed_num=12000)
    else:
        ung = unground_cloud(seed=9, config="small")
        p = kitti_params(fixed_num_downsampling=0, pca_down_rate=1)
    ctx = Context(0, 1, 16, 200000)
    g = ctx.classify_nground(ung, p)
    o = oracle.classify_nground(ung, p)
    _assert_same(g, o, variant)
    assert g["pillar"].shape[0] + g["beam"].shape[0] + g["facade"].shape[0] > 100
    if variant != "no_vertex":
        assert g["vertex"].shape[0] > 0
    ctx.close()


@pytest.mark.gpu
def test_gpu_classify_errors_and_empty():
    from mulls_b200.registration import Context

    ctx = Context(0, 1, 16, 1000)
    g = ctx.classify_nground(np.zeros((0, 12), np.float32), kitti_params())
    assert all(g[k].shape[0] == 0 for k in abi.OUT_NAMES)
    with pytest.raises(RuntimeError, match="-102"):
        ctx.classify_nground(unground_cloud(n_keep=2000), kitti_params())
    with pytest.raises(RuntimeError, match="-103"):
        ctx.classify_nground(unground_cloud(n_keep=500), kitti_params(use_distance_adaptive_pca=1))
    with pytest.raises(RuntimeError, match="-101"):
        ctx.classify_nground(unground_cloud(n_keep=500), kitti_params(neighbor_k=100))
    tiny = ctx.classify_nground(unground_cloud(n_keep=8), kitti_params())  # below every threshold
    assert tiny["unground"].shape[0] == 8 and tiny["pillar"].shape[0] == 0
    ctx.close()
