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
