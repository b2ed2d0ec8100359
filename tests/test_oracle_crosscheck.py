"""Independent numpy/scipy restatement of the first ICP iteration, cross-checking the C++ oracle
(SURVEY.md §8c item iii): NN with scipy's cKDTree, determine_corres and the point-to-plane /
point-to-line accumulation written again from the reference's formulas with numpy float32/float64
arithmetic. Agreement here means two separately written restatements of cregistration.hpp compute
the same normal equations; it does not pin the reference binary (PARITY UNPINNED)."""
import numpy as np
from scipy.spatial import cKDTree

from mulls_b200 import abi

F32 = np.float32


def flann_d2(p, q):
    d = (p - q).astype(F32)
    return ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(F32) + d[:, 2] * d[:, 2]).astype(F32)


def test_nn_against_ckdtree(oracle_mod, small_pair):
    for c in (abi.GROUND, abi.FACADE, abi.PILLAR):
        tgt, src = small_pair["tgt"][c], small_pair["src"][c]
        idx, d2 = oracle_mod.nn(tgt, src, 3.5)
        dist, j = cKDTree(tgt[:, :3].astype(np.float64)).query(src[:, :3].astype(np.float64), k=1)
        d2_ref = flann_d2(src[:, :3], tgt[j, :3])
        within = d2_ref.astype(np.float64) <= 3.5 ** 2
        m = idx >= 0
        # same matched set, and the same float distance (indices may differ only on exact float ties)
        assert (m == within).mean() > 0.9999
        both = m & within
        assert (d2[both] <= d2_ref[both]).all()  # the oracle minimises the FLOAT distance
        assert (idx[both] == j[both]).mean() > 0.999
        np.testing.assert_allclose(d2[both], d2_ref[both], rtol=2e-6)


def _weights(q, pi, qi, it):
    dist = np.sqrt((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2]).astype(F32)).astype(F32)
    b = min(F32(0.7) + F32(0.05) * F32(it), F32(1.3))
    wd = (np.float64(b) + (1.0 - np.float64(b)) * dist.astype(np.float64) / 30.0).astype(F32)
    wd = np.where(wd.astype(np.float64) > 0.01, wd, F32(0.01))
    i1 = (pi.astype(np.float64) + 0.0001).astype(F32)
    i2 = (qi.astype(np.float64) + 0.0001).astype(F32)
    ratio = (np.abs(i1 - i2) / F32(255.0)).astype(F32)
    wi = np.exp(-1.0 * ratio.astype(np.float64)).astype(F32)
    return wd, wi


def test_first_iteration_normal_equations(oracle_mod, small_pair):
    P = small_pair["params"]
    res, tr = oracle_mod.icp_run(small_pair["tgt"], small_pair["src"], P, small_pair["init_guess"])
    thre = F32(P.dis_thre_unit)
    rmax = float(F32(2.5) * thre)
    cos_thre = np.cos(P.normal_bearing / 180.0 * np.pi)
    # intersection filter (initial guess is identity for this pair)
    sb = [small_pair["src"][c][:, :3].astype(np.float64) for c in (abi.GROUND, abi.PILLAR, abi.FACADE)]
    smin = np.min([s.min(0) for s in sb if len(s)], axis=0)
    smax = np.max([s.max(0) for s in sb if len(s)], axis=0)
    tb = np.array(P.target_bound[:])
    lo = np.maximum(tb[:3], smin) - 1.0
    hi = np.minimum(tb[3:], smax) + 1.0

    def crop(a):
        x = a[:, :3].astype(np.float64)
        return a[((x > lo) & (x < hi)).all(1)]

    corr = {}
    for c in range(5):
        tgt, src = crop(small_pair["tgt"][c]), crop(small_pair["src"][c])
        idx, d2 = oracle_mod.nn(tgt, src, rmax)  # NN itself is cross-checked above
        s_i = np.flatnonzero(idx >= 0)
        t_i = idx[s_i]
        if len(src) >= 500:  # duplicate check: first source (in order) claiming a target wins
            _, first = np.unique(t_i, return_index=True)
            keep = np.zeros(len(s_i), bool)
            keep[first] = True
            s_i, t_i = s_i[keep], t_i[keep]
        dd = d2[s_i]
        ok = dd < thre * thre
        s_i, t_i = s_i[ok], t_i[ok]
        dot = (src[s_i, 4:7].astype(np.float64) * tgt[t_i, 4:7].astype(np.float64)).sum(1)
        ok = ~(np.abs(dot).astype(F32).astype(np.float64) < cos_thre)
        corr[c] = (src[s_i[ok]], tgt[t_i[ok]])
    counts = np.array([len(corr[c][0]) for c in range(5)] + [0])
    np.testing.assert_array_equal(counts, tr["n_corr"][0])

    m1 = counts[abi.GROUND] + counts[abi.ROOF]
    m2, m3, m4 = counts[abi.FACADE], counts[abi.PILLAR], counts[abi.BEAM]
    w_ground = F32(max(0.01, float(F32(P.z_xy_balanced_ratio) * F32(m2 + 2 * m3 - m4)) / (0.0001 + 2.0 * m1)))
    A = np.zeros((6, 6))
    bvec = np.zeros(6)
    for c, wc in ((abi.GROUND, w_ground), (abi.FACADE, F32(1.0)), (abi.ROOF, w_ground)):
        s, t = corr[c]
        p, q, n = s[:, 0:3], t[:, 0:3], t[:, 4:7]
        wd, wi = _weights(q, s[:, 8], t[:, 8], 0)
        w = ((wc * wd).astype(F32) * wi).astype(F32)
        a = (n[:, 2] * p[:, 1] - n[:, 1] * p[:, 2]).astype(F32)
        b = (n[:, 0] * p[:, 2] - n[:, 2] * p[:, 0]).astype(F32)
        cc = (n[:, 1] * p[:, 0] - n[:, 0] * p[:, 1]).astype(F32)
        d = (((((n[:, 0] * q[:, 0] + n[:, 1] * q[:, 1]).astype(F32) + n[:, 2] * q[:, 2]).astype(F32)
               - n[:, 0] * p[:, 0]).astype(F32) - n[:, 1] * p[:, 1]).astype(F32) - n[:, 2] * p[:, 2]).astype(F32)
        J = np.stack([n[:, 0], n[:, 1], n[:, 2], a, b, cc], axis=1).astype(np.float64)
        A += (J * w[:, None].astype(np.float64)).T @ J
        bvec += (J * (w * d).astype(np.float64)[:, None]).sum(0)
    for c in (abi.PILLAR, abi.BEAM):
        s, t = corr[c]
        p, q, v = s[:, 0:3].astype(np.float64), t[:, 0:3].astype(np.float64), t[:, 4:7].astype(np.float64)
        wd, wi = _weights(t[:, 0:3], s[:, 8], t[:, 8], 0)
        w = (wd * wi).astype(F32).astype(np.float64)
        dlt = p - q
        for k in range(len(p)):
            vx, vy, vz = v[k]
            px, py, pz = p[k]
            Am = np.array([[0, -vz, vy, vy * py + vz * pz, -vy * px, -vz * px],
                           [vz, 0, -vx, -vx * py, vz * pz + vx * px, -vz * py],
                           [-vy, vx, 0, -vx * pz, -vy * pz, vx * px + vy * py]])
            bb = np.array([-vy * dlt[k, 2] + vz * dlt[k, 1], -vz * dlt[k, 0] + vx * dlt[k, 2],
                           -vx * dlt[k, 1] + vy * dlt[k, 0]])
            A[np.arange(6), np.arange(6)] += w[k] * (Am * Am).sum(0)  # only the diagonal survives (Q1)
            bvec += w[k] * (Am.T @ bb)
    # float32 intermediates are not reproduced product-by-product here: agreement to 1e-5 relative
    scale = np.abs(tr["atpa"][0]).max()
    np.testing.assert_allclose(A, tr["atpa"][0], rtol=0, atol=2e-5 * scale)
    np.testing.assert_allclose(bvec, tr["atpb"][0], rtol=0, atol=2e-5 * np.abs(tr["atpb"][0]).max())
    x = np.linalg.solve(tr["atpa"][0], tr["atpb"][0])
    np.testing.assert_allclose(x, tr["x"][0], rtol=1e-8, atol=1e-12)


def test_increment_matrix_is_the_roll_pitch_yaw_rotation_of_scipy(small_pair):
    """construct_trans_a (cregistration.hpp:2740-2764) builds Rz(gamma) Ry(beta) Rx(alpha) from the solved 6-vector
    (tx ty tz alpha beta gamma): with max_iter_num = 1 the registration result IS that matrix (:1357-1405), so it must
    equal scipy's extrinsic 'xyz' Euler rotation of the solution the trace reports."""
    from scipy.spatial.transform import Rotation

    from mulls_b200 import abi
    from oracle import oracle

    p = abi.IcpParams.from_buffer_copy(small_pair["params"])
    p.max_iter_num = 1
    res, tr = oracle.icp_run(small_pair["tgt"], small_pair["src"], p, np.eye(4))
    assert res["code"] == 1 and res["iters"] == 1
    x = np.asarray(tr["x"][0], dtype=np.float64)
    T = np.asarray(res["T"], dtype=np.float64).reshape(4, 4)
    assert np.linalg.norm(x[3:]) > 1e-4  # a real rotation, not the identity
    R = Rotation.from_euler("xyz", x[3:6]).as_matrix()
    np.testing.assert_allclose(T[:3, :3], R, rtol=0, atol=1e-13)
    np.testing.assert_allclose(T[:3, 3], x[:3], rtol=0, atol=1e-15)
    assert np.array_equal(T[3], [0.0, 0.0, 0.0, 1.0])
