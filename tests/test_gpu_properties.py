"""Size-independent properties at BASELINE's full sizes (no oracle needed at these sizes):
self-registration is the identity, a known rigid motion is recovered, the result does not depend on
the order of the input points' storage beyond what the reference's own index-order rules imply."""
import numpy as np
import pytest

from mulls_b200 import abi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from mulls_b200.registration import Context

    c = Context(0, 4, 700000, 700000)
    yield c
    c.close()


def test_self_registration_is_identity(ctx):
    pair = synth.make_pair(1010, "c2")
    same = dict(pair, src=[t.copy() for t in pair["tgt"]])
    res, _ = ctx.run_batch([same])
    r = res[0]
    assert r["code"] == 1
    dt, dr = synth.pose_error(r["T"], np.eye(4))
    assert dt < 1e-5 and dr < 1e-6
    assert r["sigma"] < 1e-4


def test_known_rigid_motion_is_recovered(ctx):
    """Source = target moved by a known transform (no noise): T must invert it."""
    pair = synth.make_pair(1011, "c2")
    M = np.eye(4)
    M[:3, :3] = synth.rpy_matrix(0.004, -0.003, 0.012)
    M[:3, 3] = (0.35, -0.2, 0.05)
    Minv = np.linalg.inv(M)
    src = []
    for t in pair["tgt"]:
        s = t.copy()
        s[:, 0:3] = (t[:, 0:3].astype(np.float64) @ Minv[:3, :3].T + Minv[:3, 3]).astype(np.float32)
        s[:, 4:7] = (t[:, 4:7].astype(np.float64) @ Minv[:3, :3].T).astype(np.float32)
        src.append(s)
    res, _ = ctx.run_batch([dict(pair, src=src)])
    r = res[0]
    assert r["code"] == 1
    dt, dr = synth.pose_error(r["T"], M)
    assert dt < 2e-4 and dr < 2e-5


def test_scan_to_map_full_size_c3(ctx):
    """BASELINE config 3 at full size: 120k source vs 600k map; recovers the ground truth."""
    pair = synth.make_pair(1012, "c3")
    assert sum(len(t) for t in pair["tgt"]) > 590000  # five scans of <= 120k returns each
    res, _ = ctx.run_batch([pair])
    r = res[0]
    assert r["code"] == 1
    dt, dr = synth.pose_error(r["T"], pair["T_gt"])
    assert dt < 0.03 and dr < 2e-3
