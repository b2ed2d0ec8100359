"""Third-party pin for the oracle's nearest-neighbour stage: OpenCV ships a descendant of FLANN (cv2.flann, kd-tree
single index, exact search), the library PCL's KdTreeFLANN wraps. On the same clouds the oracle's own
kd-tree (oracle/mulls_oracle.cpp, standing in for pcl::search::KdTree -> flann::KDTreeSingleIndex, L2_Simple<float>)
must return the same neighbour and bit-identical float squared distances. CPU only; skipped if cv2 is absent."""
import os

import numpy as np
import pytest

from conftest import load_golden_pair
from mulls_b200 import synth
from oracle import oracle

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def flann_nn(tgt_xyz, src_xyz):
    index = cv2.flann_Index(np.ascontiguousarray(tgt_xyz, dtype=np.float32), dict(algorithm=4, leaf_max_size=15))  # KDTREE_SINGLE
    ind, d2 = index.knnSearch(np.ascontiguousarray(src_xyz, dtype=np.float32), 1, params=dict(checks=-1, eps=0.0, sorted=True))
    return ind.ravel(), d2.ravel()


def _check(tgt, src, tag):
    oi, od = oracle.nn(tgt, src, 1e9)
    fi, fd = flann_nn(tgt[:, :3], src[:, :3])
    assert np.array_equal(od.view(np.uint32), fd.view(np.uint32)), f"{tag}: squared distances differ"
    differ = np.flatnonzero(oi != fi)
    # a different index is only acceptable on an exact tie of the float distance (the oracle then takes the lower index)
    for q in differ:
        a = tgt[oi[q], :3] - src[q, :3]
        b = tgt[fi[q], :3] - src[q, :3]
        da = np.float32(np.float32(a[0] * a[0] + a[1] * a[1]) + a[2] * a[2])
        db = np.float32(np.float32(b[0] * b[0] + b[1] * b[1]) + b[2] * b[2])
        assert da == db and oi[q] < fi[q], (tag, q)
    return len(differ)


def test_oracle_nn_equals_flann_on_a_synthetic_scan_pair():
    pair = synth.make_pair(1001, "c2")
    for c in range(5):
        if len(pair["tgt"][c]) >= 10:
            _check(pair["tgt"][c], pair["src"][c], f"class {c}")


def test_oracle_nn_equals_flann_on_the_real_data_fixture():
    pair, _ = load_golden_pair(os.path.join(ROOT, "tests", "golden", "demo_pair.npz"))
    for c in range(6):
        if len(pair["tgt"][c]) >= 10 and len(pair["src"][c]):
            _check(pair["tgt"][c], pair["src"][c], f"class {c}")


def test_oracle_pca_neighbourhoods_equal_flann_radius_search():
    """PrincipleComponentAnalysis::get_pc_pca_feature (pca.hpp:294-354) takes the K nearest neighbours within R from
    KdTreeFLANN::radiusSearch. The oracle's neighbourhoods (size, eigenvalues, normal) against cv2.flann's radius search
    followed by a float64 numpy eigen-decomposition of the same covariance."""
    from test_classify import unground_cloud

    ung = unground_cloud(n_keep=20000)
    R, K = 0.7, 25
    o = oracle.pca_features(ung, R, K, 1)
    xyz = np.ascontiguousarray(ung[:, :3])
    index = cv2.flann_Index(xyz, dict(algorithm=4, leaf_max_size=15))
    r2 = float(np.float32(np.float64(np.float32(R)) ** 2))  # KdTreeFLANN::radiusSearch: (float)(radius * radius)
    checked = 0
    for q in np.random.default_rng(0).choice(len(ung), 400, replace=False):
        n, ind, _ = index.radiusSearch(xyz[q:q + 1], r2, K, params=dict(checks=-1, eps=0.0, sorted=True))
        m = min(int(n), K)
        assert m == o["pt_num"][q], q
        if m <= 3:
            continue
        nb = xyz[ind[0, :m]].astype(np.float64)
        w, v = np.linalg.eigh(np.cov(nb.T))  # ascending; cov / (n - 1) like pcl::PCA
        ev = o["eigenvalues"][q].astype(np.float64)  # descending
        assert np.allclose(ev, w[::-1], rtol=2e-3, atol=1e-6), (q, ev, w[::-1])
        if w[1] - w[0] > 1e-3 * w[2]:  # the normal is well defined
            nrm = o["normal"][q].astype(np.float64)
            assert abs(nrm @ v[:, 0]) > np.cos(np.radians(1.0)), q
        checked += 1
    assert checked > 300
