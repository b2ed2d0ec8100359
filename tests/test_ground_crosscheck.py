"""Independent cross-checks of the ground-filter restatement in oracle/ (the reference holds no expected outputs for it,
SURVEY §8c): a second, differently structured numpy restatement of CFilter::fast_ground_filter with
estimate_ground_normal_method = 0 (cfilter.hpp:1658-2036, vectorised over cells instead of walking the cloud), numpy's
own MT19937 against the sample stream the plane fit tabulates, and a pure-Python PCL-style RANSAC driven by that
stream against the oracle's plane fit. CPU only."""
import numpy as np
import pytest

from mulls_b200 import abi
from oracle import oracle
from test_ground import params, raw_scan


def numpy_ground_filter_method0(raw, P):
    """fast_ground_filter with the fixed (0,0,1) normal and no distance weighting, written from the reference's comments
    (:1645-1656) rather than from its loops: per-cell minima by ufunc.at, neighbourhood minima by shifted views,
    position-in-cell by a stable sort."""
    n = raw.shape[0]
    z = raw[:, 2]
    s = np.float32(0.001)
    for j in range(0, n, 100):
        s = np.float32(s + z[j])
    mean_h = np.float32(s / np.float32(len(range(0, n, 100))))
    high_thre = np.float32(mean_h + np.float32(P.max_ground_height))
    x64, y64 = raw[:, 0].astype(np.float64), raw[:, 1].astype(np.float64)
    min_x, min_y, max_x, max_y = x64.min(), y64.min(), x64.max(), y64.max()
    res = np.float64(np.float32(P.grid_resolution))
    row, col = int(np.ceil((max_y - min_y) / res)), int(np.ceil((max_x - min_x) / res))
    cid = np.floor((y64 - min_y) / res).astype(np.int64) * col + np.floor((x64 - min_x) / res).astype(np.int64)
    inside = (cid >= 0) & (cid < row * col)
    high = inside & (z > high_thre)
    counted = inside & ~high & (z > -np.finfo(np.float32).max)
    out_u = []
    rate_u = P.nonground_random_down_rate
    jh = np.flatnonzero(high)
    jh = jh[(jh % rate_u == 0) | (raw[jh, 8] > P.intensity_thre)]
    hu = raw[jh].copy()
    hu[:, 3] = (z[jh].astype(np.float64) - (np.float64(mean_h) - 3.0)).astype(np.float32)
    out_u.append(hu)
    G = row * col
    min_z = np.full(G, np.finfo(np.float32).max, np.float32)
    np.minimum.at(min_z, cid[counted], z[counted])
    cnt = np.bincount(cid[counted], minlength=G)
    grid_min = min_z.reshape(row, col)
    nb = grid_min.copy()
    rel = np.zeros((row, col), np.int64)
    reliable = (cnt.reshape(row, col) > P.min_grid_pt_num - 1)
    if row > 2 and col > 2:
        acc = np.full((row - 2, col - 2), np.finfo(np.float32).max, np.float32)
        racc = np.zeros((row - 2, col - 2), np.int64)
        for dj in (0, 1, 2):
            for dk in (0, 1, 2):
                acc = np.minimum(acc, grid_min[dj:dj + row - 2, dk:dk + col - 2])
                racc += reliable[dj:dj + row - 2, dk:dk + col - 2]
        nb[1:-1, 1:-1] = np.minimum(nb[1:-1, 1:-1], acc)
        rel[1:-1, 1:-1] = racc
    nb, rel = nb.reshape(-1), rel.reshape(-1)
    idx = np.flatnonzero(counted)
    order = idx[np.argsort(cid[idx], kind="stable")]          # cell by cell, index order inside a cell
    c_sorted = cid[order]
    first = np.r_[0, np.flatnonzero(np.diff(c_sorted)) + 1]
    pos = np.arange(len(order)) - np.repeat(first, np.diff(np.r_[first, len(order)]))
    ok_cell = (cnt[c_sorted] >= P.min_grid_pt_num) & (rel[c_sorted] >= P.reliable_neighbor_grid_num_thre)
    ground_cell = (min_z[c_sorted] - nb[c_sorted]) < np.float32(P.neighbor_height_diff)
    zz = z[order]
    low = (zz - min_z[c_sorted]) < np.float32(P.max_height_difference)
    take_u = (pos % rate_u == 0) | (raw[order, 8] > P.intensity_thre)
    is_g = ok_cell & ground_cell & low & (pos % P.ground_random_down_rate == 0)
    is_u = ok_cell & take_u & ((ground_cell & ~low) | ~ground_cell)
    g = raw[order[is_g]].copy()
    g[:, 4:7] = (0.0, 0.0, 1.0)
    u = raw[order[is_u]].copy()
    ref = np.where(ground_cell[is_u], min_z[c_sorted[is_u]], nb[c_sorted[is_u]])
    u[:, 3] = zz[is_u] - ref
    out_u.append(u)
    return {"ground": g, "ground_down": g[::P.ground_random_down_down_rate], "unground": np.concatenate(out_u, axis=0)}


@pytest.mark.parametrize("kw", [dict(), dict(max_ground_height=0.8, nonground_random_down_rate=5),
                                dict(grid_resolution=1.5, min_grid_pt_num=6, reliable_neighbor_grid_num_thre=4,
                                     intensity_thre=180.0)])
def test_numpy_restatement_of_the_fixed_normal_ground_filter_equals_the_oracle(kw):
    raw, _ = raw_scan()
    P = params(estimate_ground_normal_method=0, distance_weight_downsampling_method=0, **kw)
    o = oracle.fast_ground_filter(raw, P)
    m = numpy_ground_filter_method0(raw, P)
    for k in ("ground", "ground_down", "unground"):
        assert o[k].shape == m[k].shape, (k, o[k].shape, m[k].shape)
        assert np.array_equal(o[k].view(np.uint32), m[k].view(np.uint32)), k


def mt19937_12345(count):
    bg = np.random.MT19937()
    bg._legacy_seeding(12345)  # init_genrand(12345) == std::mt19937(12345u) == boost::mt19937 seeded the same way
    return bg.random_raw(count).astype(np.uint64)


def python_pcl_ransac_plane(pts, threshold, max_iterations=20):
    """RandomSampleConsensus + SampleConsensusModelPlane as published for PCL 1.10, float32 arithmetic via numpy scalars,
    sample stream from numpy's MT19937; returns the best model's inlier indices BEFORE coefficient optimisation."""
    f = np.float32
    n = len(pts)
    draws = mt19937_12345(3 * 1000)
    nxt = 0
    shuffled = list(range(n))
    iterations, n_best, k, skipped = 0, -2 ** 31, 1.0, 0
    best, best_c = None, None
    logp = np.log(1.0 - 0.99)
    while iterations < k and skipped < max_iterations * 10:
        for _ in range(1000):
            for i in range(3):
                r = int(draws[nxt]) >> 1
                nxt += 1
                o = i + r % (n - i)
                shuffled[i], shuffled[o] = shuffled[o], shuffled[i]
            s = shuffled[:3]
            p0, p1, p2 = pts[s[0]], pts[s[1]], pts[s[2]]
            with np.errstate(all="ignore"):
                d = (p1 - p0) / (p2 - p0)
            if (d[0] != d[1]) or (d[2] != d[1]):
                break
        a, b = p1 - p0, p2 - p0
        c = np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0], f(0)], np.float32)
        zz = f(f(c[0] * c[0] + c[2] * c[2]) + f(c[1] * c[1]))
        c = (c / np.sqrt(zz)).astype(np.float32)
        c[3] = f(-1) * f(f(c[0] * p0[0] + c[2] * p0[2]) + f(c[1] * p0[1]))
        dist = np.abs((c[0] * pts[:, 0] + c[2] * pts[:, 2]).astype(np.float32) + (c[1] * pts[:, 1] + c[3]).astype(np.float32))
        cnt = int((dist.astype(np.float64) < threshold).sum())
        if cnt > n_best:
            n_best, best, best_c = cnt, np.flatnonzero(dist.astype(np.float64) < threshold), c.copy()
            w = n_best / n
            p_no = min(max(1.0 - w ** 3, np.finfo(np.float64).eps), 1.0 - np.finfo(np.float64).eps)
            k = logp / np.log(p_no)
        iterations += 1
        if iterations > max_iterations:
            break
    return best, best_c, iterations


def test_sample_stream_and_ransac_against_an_independent_implementation():
    # (1) the tabulated mt19937(12345) stream: the oracle's plane fit on three points must pick them in the order numpy's
    #     generator dictates; checked through a full independent RANSAC below. Known first output of mt19937(12345):
    assert int(mt19937_12345(1)[0]) == 3992670690
    import ctypes as C

    lib = oracle.load()
    tab = np.zeros(16384, np.uint32)
    assert lib.orc_sac_draws(tab.ctypes.data_as(C.POINTER(C.c_uint32)), 16384) == 16384
    assert np.array_equal(tab.astype(np.uint64), mt19937_12345(16384))  # numpy's MT19937 == the tabulated stream
    # (2) a noisy plane with outliers: every model the independent RANSAC scores best must lead to the oracle's final
    #     inlier set after the oracle's own refinement step (refined set contains >= 90 % of the RANSAC consensus set)
    rng = np.random.default_rng(3)
    pts = np.zeros((240, 12), np.float32)
    pts[:, 0:2] = rng.uniform(-1.5, 1.5, (240, 2))
    pts[:, 2] = (0.08 * pts[:, 0] + 0.01 * pts[:, 1] - 1.7 + rng.normal(0, 0.012, 240)).astype(np.float32)
    pts[::9, 2] += np.float32(0.25)
    best, best_c, its = python_pcl_ransac_plane(pts[:, :3].copy(), float(np.float32(0.09)))
    ok, inl, coeff = oracle.sac_plane(pts, float(np.float32(0.09)), 20)
    assert ok and best is not None
    # the RANSAC stage itself, bit for bit: same winning sample model, same number of trials
    rc = np.zeros(4, np.float32)
    n_it = C.c_int32(0)
    lib.orc_sac_plane_ransac.restype = C.c_int
    lib.orc_sac_plane_ransac.argtypes = [abi.CloudView, C.c_double, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    assert lib.orc_sac_plane_ransac(abi.cloud_view(pts), float(np.float32(0.09)), 20, rc.ctypes.data_as(C.POINTER(C.c_float)),
                                    C.byref(n_it)) == 1
    assert n_it.value == its
    assert np.array_equal(rc.view(np.uint32), best_c.view(np.uint32)), (rc, best_c)
    common = np.intersect1d(best, inl).size
    assert common >= 0.9 * len(best) and abs(len(inl) - len(best)) <= 0.1 * len(best)
    # the refined plane is the least-squares plane of the consensus set: its normal is within 0.5 degrees of numpy's SVD fit
    q = pts[best, :3].astype(np.float64)
    _, _, vt = np.linalg.svd(q - q.mean(0))
    nrm = vt[2] * np.sign(vt[2][2])
    c = coeff[:3].astype(np.float64) * np.sign(coeff[2])
    assert np.degrees(np.arccos(np.clip(abs(nrm @ c), -1, 1))) < 0.5
