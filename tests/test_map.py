"""Local-map maintenance: lo::MapManager::update_local_map, src/map_manager.cpp:17-145.

CPU part: the restatement in oracle/ against the properties the reference code states (radius crop, point budget,
append order, vertex quirk). GPU part: the device-resident map of libmulls_b200.so against that restatement, bit for
bit, over a short odometry run (update -> scan-to-map registration -> update ...), through the C-ABI.
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
EMPTY = [np.zeros((0, 12), np.float32) for _ in range(6)]


def map_params(**kw):
    p = abi.default_map_params()
    p.local_map_radius = 40.0
    p.max_num_pts = 6000
    p.kept_vertex_num = 150
    p.random_seed = 7
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def test_map_structs_match_header():
    code = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "mulls_b200/abi.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu\n", sizeof(mulls_map_params), sizeof(mulls_map_info),
             offsetof(mulls_map_params, used_feature_type), offsetof(mulls_map_params, random_seed),
             offsetof(mulls_map_info, n_appended));
      return 0; }"""
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(code)
        exe = os.path.join(td, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = [int(v) for v in subprocess.check_output([exe]).decode().split()]
    assert out == [C.sizeof(abi.MapParams), C.sizeof(abi.MapInfo), abi.MapParams.used_feature_type.offset,
                   abi.MapParams.random_seed.offset, abi.MapInfo.n_appended.offset]


def test_default_map_params_match_reference_defaults():
    lib = abi.load_library()
    p = abi.MapParams()
    lib.mulls_map_default_params(C.byref(p))
    q = abi.default_map_params()
    for name, _ in abi.MapParams._fields_:
        assert getattr(p, name) == getattr(q, name), name
    # include/pgo/map_manager.h:22-32
    assert p.local_map_radius == 80 and p.max_num_pts == 20000 and p.kept_vertex_num == 800
    assert p.used_feature_type == b"111110" and abs(p.near_dist_thre - 0.03) < 1e-8


def test_oracle_first_update_is_append_crop_and_budget():
    seq = synth.make_sequence(3, 2)
    sc = seq["scans"][0]
    p = map_params()
    out, info = oracle.map_update(EMPTY, np.eye(4), sc, np.eye(4), p)
    # identity poses: the transforms are exact, so the crop can be checked on the inputs (cfilter.hpp:838-873)
    for c in range(6):
        used = c == 5 or p.used_feature_type[c:c + 1] == b"1"
        a = sc[c]
        inside = a[(a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1]).astype(np.float64) < 40.0 ** 2] if used else a[:0]
        assert info["n_appended"][c] == (a.shape[0] if used else 0)
        assert out[c].shape[0] <= inside.shape[0]
        # the kept points are a subsequence of the cropped cloud, in order
        if out[c].shape[0]:
            keys = {tuple(r) for r in inside[:, [0, 1, 2, 4, 5, 6, 8]].tolist()}
            assert all(tuple(r) in keys for r in out[c][:, [0, 1, 2, 4, 5, 6, 8]].tolist())
    # budget of map_manager.cpp:69-85
    n_crop = []
    for c in range(5):
        a = sc[c]
        n_crop.append(int(((a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1]).astype(np.float64) < 1600.0).sum())
                      if p.used_feature_type[c:c + 1] == b"1" else 0)
    total = sum(n_crop)
    for c in range(5):
        kept = int(1.0 * p.max_num_pts / total * n_crop[c] + 1)
        assert info["n"][c] == min(n_crop[c], kept)
    assert info["n"][5] == min(150, out[5].shape[0]) and info["n"][5] <= 150
    assert info["feature_point_num"] == int(info["n"][:5].sum())
    allp = np.concatenate([o[:, :3] for o in out if len(o)]).astype(np.float64)
    assert np.array_equal(info["local_bound"], np.concatenate([allp.min(0), allp.max(0)]))


def test_oracle_update_is_deterministic_and_moves_the_map_into_the_scan_frame():
    seq = synth.make_sequence(4, 3)
    p = map_params()
    m0, i0 = oracle.map_update(EMPTY, np.eye(4), seq["scans"][0], seq["poses"][0], p)
    m1, i1 = oracle.map_update(m0, i0["pose_lo"], seq["scans"][1], seq["poses"][1], p)
    m1b, i1b = oracle.map_update(m0, i0["pose_lo"], seq["scans"][1], seq["poses"][1], p)
    for a, b in zip(m1, m1b):
        assert np.array_equal(a, b)
    assert np.array_equal(i1["pose_lo"], seq["poses"][1])
    # in the world frame the two scans' ground points are the same plane z ~ 0
    g = m1[0].astype(np.float64)
    w = g[:, :3] @ seq["poses"][1][:3, :3].T + seq["poses"][1][:3, 3]
    assert abs(np.median(w[:, 2]) + synth.SENSOR_HEIGHT) < 0.1 or abs(np.median(w[:, 2])) < 0.1
    # vertex quirk (map_manager.cpp:32 + utility.hpp:469): the scan's pc_vertex is appended in the SCAN frame and then
    # moved by tran_target_map, i.e. it ends up one motion step off
    only_v = [np.zeros((0, 12), np.float32)] * 5 + [seq["scans"][1][5]]
    pv = map_params(kept_vertex_num=10 ** 6, local_map_radius=1000.0)
    mv, _ = oracle.map_update(EMPTY, seq["poses"][0], only_v, seq["poses"][1], pv)
    T = np.linalg.inv(seq["poses"][1]) @ seq["poses"][0]
    v = seq["scans"][1][5][:, :3].astype(np.float64)
    assert np.allclose(mv[5][:, :3], v @ T[:3, :3].T + T[:3, 3], atol=1e-4)


def test_oracle_dynamic_removal_drops_points_close_to_the_map():
    seq = synth.make_sequence(5, 2)
    p = map_params(max_num_pts=20000)
    m0, i0 = oracle.map_update(EMPTY, np.eye(4), seq["scans"][0], np.eye(4), p)
    pd = map_params(max_num_pts=20000, map_based_dynamic_removal_on=1, near_dist_thre=0.2)
    # trees = the map itself; the same scan again: every pillar/beam/facade point inside the centre radius has a
    # neighbour at distance 0 and must go ((0, near] is filtered)
    m1, i1 = oracle.map_update(m0, i0["pose_lo"], seq["scans"][0], np.eye(4), pd, trees=m0)
    m2, i2 = oracle.map_update(m0, i0["pose_lo"], seq["scans"][0], np.eye(4), p)
    for c in (abi.PILLAR, abi.BEAM, abi.FACADE):
        a = seq["scans"][0][c]
        if a.shape[0] <= 10:
            continue
        far = int((a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1] > np.float32(30.0) ** 2).sum())
        in_map = {tuple(r) for r in m0[c][:, :3].tolist()}
        near_unmatched = sum(1 for r in a[(a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1] <= np.float32(900.0))][:, :3].tolist()
                             if tuple(r) not in in_map)
        assert far <= i1["n_appended"][c] <= far + near_unmatched
        assert i2["n_appended"][c] == a.shape[0]
    assert i1["n_appended"][abi.GROUND] == seq["scans"][0][abi.GROUND].shape[0]


def test_oracle_recalculated_vectors_follow_the_geometry():
    """update_cloud_vectors (map_manager.cpp:260-295): pillars keep steep directions, beams flat ones; the direction
    comes from the map's own geometry, the linearity lands in `curvature`."""
    seq = synth.make_sequence(6, 1, "c2")
    sc = seq["scans"][0]
    p = map_params(max_num_pts=40000, local_map_radius=60.0)
    plain, i0 = oracle.map_update(EMPTY, np.eye(4), sc, np.eye(4), p)
    pr = map_params(max_num_pts=40000, local_map_radius=60.0, recalculate_feature_on=1)
    rec, i1 = oracle.map_update(EMPTY, np.eye(4), sc, np.eye(4), pr)
    for c in (abi.GROUND, abi.FACADE, abi.ROOF, abi.VERTEX):
        assert np.array_equal(plain[c], rec[c])
    assert np.array_equal(i0["local_bound"], i1["local_bound"])  # boxes are computed before the recalculation
    pil, beam = rec[abi.PILLAR], rec[abi.BEAM]
    assert 0 < pil.shape[0] <= plain[abi.PILLAR].shape[0] and 0 < beam.shape[0] <= plain[abi.BEAM].shape[0]
    assert (np.abs(pil[:, 6]) > 0.80).all() and (np.abs(beam[:, 6]) < 0.25).all()
    assert (pil[:, 9] > 0.65).all() and (beam[:, 9] > 0.65).all() and (pil[:, 9] <= 1.0).all()
    assert np.allclose(np.linalg.norm(pil[:, 4:7], axis=1), 1.0, atol=1e-5)
    # survivors are a subsequence of the un-recalculated cloud
    keys = {tuple(r) for r in plain[abi.PILLAR][:, :3].tolist()}
    assert all(tuple(r) in keys for r in pil[:, :3].tolist())
    assert i1["feature_point_num"] == int(i1["n"][:5].sum()) == sum(r.shape[0] for r in rec[:5])


# ------------------------------------------------------------------------------------------------
# GPU: the device-resident map against the restatement
# ------------------------------------------------------------------------------------------------
def _assert_maps_equal(gpu, orc, tag):
    for c in range(6):
        assert gpu[c].shape == orc[c].shape, f"{tag}: class {c} size {gpu[c].shape} vs {orc[c].shape}"
        assert np.array_equal(gpu[c].view(np.uint32), orc[c].view(np.uint32)), f"{tag}: class {c} differs"


def _assert_info_equal(gi, oi, tag):
    for k in ("n", "n_appended"):
        assert np.array_equal(gi[k], oi[k]), f"{tag}: {k} {gi[k]} vs {oi[k]}"
    assert gi["feature_point_num"] == oi["feature_point_num"]
    assert np.array_equal(gi["pose_lo"], oi["pose_lo"])
    assert np.array_equal(gi["local_bound"], oi["local_bound"]), f"{tag}: {gi['local_bound']} vs {oi['local_bound']}"
    assert np.array_equal(gi["bound"], oi["bound"]), f"{tag}: {gi['bound']} vs {oi['bound']}"


@pytest.mark.gpu
@pytest.mark.parametrize("dynamic", [0, 1])
def test_gpu_local_map_odometry_matches_oracle(dynamic):
    from mulls_b200.map_manager import LocalMap
    from mulls_b200.registration import Context

    n_frames = 5
    seq = synth.make_sequence(11, n_frames)
    ctx = Context(0, 1, 60000, 120000)
    lm = LocalMap(ctx, 1 << 16)
    p = map_params(map_based_dynamic_removal_on=dynamic, max_num_pts=6000)
    icp = seq["params"]
    omap, opose = EMPTY, np.eye(4)
    pose_prev = np.eye(4)
    motion = np.eye(4)
    trees = None
    n_removed = 0
    for k in range(n_frames):
        sc = seq["scans"][k]
        if k == 0:
            pose = np.eye(4)
        else:
            # scan-to-map registration: block1 = the map (in frame k-1), block2 = the new scan
            o_icp = abi.IcpParams.from_buffer_copy(icp)
            o_icp.target_bound[:] = list(oinfo["local_bound"])
            r_o, trees = oracle.icp_run_trees(omap, sc, o_icp, motion)
            r_g, _ = lm.icp_run(sc, icp, motion)
            assert r_g["code"] == r_o["code"] == 1
            assert np.array_equal(r_g["n_corr"], r_o["n_corr"])
            assert np.allclose(r_g["T"], r_o["T"], atol=1e-9, rtol=0)
            dt, dr = synth.pose_error(pose_prev @ r_g["T"], seq["poses"][k])
            assert dt < 0.05 * k + 0.05 and dr < 0.01
            motion = r_g["T"]
            pose = pose_prev @ r_g["T"]
        ginfo = lm.update(sc, pose, p)
        omap, oinfo = oracle.map_update(omap, opose, sc, pose, p, trees=trees if dynamic else None)
        opose = oinfo["pose_lo"]
        _assert_info_equal(ginfo, oinfo, f"frame {k}")
        _assert_maps_equal(lm.download(), omap, f"frame {k}")
        n_removed += int(sum(sc[c].shape[0] for c in (1, 2, 3)) - oinfo["n_appended"][[1, 2, 3]].sum())
        pose_prev = pose
    assert (n_removed > 0) == bool(dynamic)
    assert oinfo["feature_point_num"] <= p.max_num_pts + 5
    lm.close()
    ctx.close()


@pytest.mark.gpu
def test_gpu_map_set_download_and_run_to_map_equals_host_target():
    """mulls_icp_run_to_map(map) == mulls_icp_run(host copy of the map): the resident target is the same data."""
    from mulls_b200.map_manager import LocalMap
    from mulls_b200.registration import Context

    pr = synth.make_pair(21, "small")
    ctx = Context(0, 1, 60000, 60000)
    lm = LocalMap(ctx, 1 << 16)
    lm.set(pr["tgt"], np.eye(4))
    back = lm.download()
    for c in range(6):
        assert np.array_equal(back[c], pr["tgt"][c])
    info = lm.info()
    assert np.allclose(info["local_bound"], synth.cloud_bound(pr["tgt"]))
    r_map, tr_map = lm.icp_run(pr["src"], pr["params"], pr["init_guess"], want_trace=True)
    (r_host,), (tr_host,) = ctx.run_batch([pr], want_trace=True)
    assert r_map["code"] == r_host["code"] == 1
    assert np.array_equal(r_map["T"], r_host["T"]) and np.array_equal(tr_map["atpa"], tr_host["atpa"])
    lm.close()
    ctx.close()


@pytest.mark.gpu
def test_gpu_map_errors():
    from mulls_b200.map_manager import LocalMap
    from mulls_b200.registration import Context

    seq = synth.make_sequence(2, 1)
    ctx = Context(0, 1, 60000, 60000)
    lm = LocalMap(ctx, 4096)
    with pytest.raises(RuntimeError, match="-102"):  # capacity
        lm.update(seq["scans"][0], np.eye(4), map_params())
    lm2 = LocalMap(ctx, 1 << 16)
    lm2.update(seq["scans"][0], np.eye(4), map_params(max_num_pts=2000))
    with pytest.raises(RuntimeError, match="-101"):  # dynamic removal without the preceding scan-to-map registration
        lm2.update(seq["scans"][0], np.eye(4), map_params(max_num_pts=2000, map_based_dynamic_removal_on=1))
    lm.close()
    lm2.close()
    ctx.close()


@pytest.mark.gpu
def test_gpu_map_recalculate_feature_matches_oracle():
    """recalculate_feature_on: the re-estimated directions and linearities are bit-identical to the restatement's (float
    mean / covariance in radiusSearch order, fp64 Jacobi; the reference solves with float Eigen)."""
    from mulls_b200.map_manager import LocalMap
    from mulls_b200.registration import Context

    seq = synth.make_sequence(6, 2, "c2")
    ctx = Context(0, 1, 200000, 400000)
    lm = LocalMap(ctx, 1 << 18)
    p0 = map_params(max_num_pts=40000, local_map_radius=60.0)
    p1 = map_params(max_num_pts=40000, local_map_radius=60.0, recalculate_feature_on=1)
    lm.update(seq["scans"][0], seq["poses"][0], p0)
    omap, oinfo = oracle.map_update(EMPTY, np.eye(4), seq["scans"][0], seq["poses"][0], p0)
    ginfo = lm.update(seq["scans"][1], seq["poses"][1], p1)
    omap, oinfo = oracle.map_update(omap, oinfo["pose_lo"], seq["scans"][1], seq["poses"][1], p1)
    _assert_info_equal(ginfo, oinfo, "recalc")
    g = lm.download()
    assert g[abi.PILLAR].shape[0] > 50 and g[abi.BEAM].shape[0] > 50
    _assert_maps_equal(g, omap, "recalc")  # the PCA accumulates in radiusSearch order on both sides: identical bits
    lm.close()
    ctx.close()
