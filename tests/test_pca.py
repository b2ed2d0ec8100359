"""PCA neighbourhood features (pca.hpp:294-354): oracle sanity on CPU, CUDA kernel vs oracle on GPU.

Floating-point parity bar (stated here as the task requires): neighbour COUNT exact; eigenvalues within
1e-4 relative to the largest eigenvalue; eigenvector directions within 1e-3 rad up to sign, checked where
the eigenvalue gap makes the direction well conditioned (gap > 5% of lambda1). The reference computes the
covariance in float (pcl::PCA) and solves with Eigen's float SelfAdjointEigenSolver; both restatements use
fp64 Jacobi on the same neighbour set, so they agree far tighter than that bar — for k <= 64 (the reference's
range) the CUDA path accumulates the float covariance in radiusSearch order like the CPU path and the outputs are
bit-identical; for larger / unlimited k it reduces in fp64 across the warp and the stated tolerance applies."""
import numpy as np
import pytest

from mulls_b200 import abi, synth


def _cloud(n_keep=20000):
    pair = synth.make_pair(1000, "small")
    allp = np.concatenate([pair["tgt"][c] for c in (abi.GROUND, abi.FACADE, abi.PILLAR, abi.BEAM)], axis=0)
    return np.ascontiguousarray(allp[:n_keep])


def test_oracle_pca_planes_and_lines(oracle_mod):
    rng = np.random.default_rng(0)
    plane = np.zeros((4000, 12), np.float32)
    plane[:, 0:2] = rng.uniform(-3, 3, (4000, 2))
    plane[:, 2] = rng.normal(0, 0.002, 4000)
    out = oracle_mod.pca_features(plane, 0.6, 30, 1)
    m = out["pt_num"] == 30
    assert m.mean() > 0.95
    assert np.abs(out["normal"][m, 2]).min() > 0.99          # normal of a z=0 plane
    lam = out["eigenvalues"][m]
    assert (lam[:, 0] >= lam[:, 1]).all() and (lam[:, 1] >= lam[:, 2]).all()
    line = np.zeros((2000, 12), np.float32)
    line[:, 0] = np.linspace(-5, 5, 2000)
    line[:, 1:3] = rng.normal(0, 0.002, (2000, 2))
    out = oracle_mod.pca_features(line, 0.5, 25, 2)
    assert (out["pt_num"][1::2] == 0).all()                   # stride: skipped points stay empty
    sel = out["pt_num"] > 3
    assert np.abs(out["principal"][sel, 0]).min() > 0.99      # principal direction along x


@pytest.mark.gpu
@pytest.mark.parametrize("radius,k,stride", [(0.6, 25, 1), (1.0, 50, 3), (0.3, 0, 1)])
def test_gpu_pca_matches_oracle(oracle_mod, radius, k, stride):
    from mulls_b200.registration import Context

    cloud = _cloud()
    ctx = Context(0, 1, 16, 100000)
    g = ctx.pca_features(cloud, radius, k, stride)
    o = oracle_mod.pca_features(cloud, radius, k, stride)
    np.testing.assert_array_equal(g["pt_num"], o["pt_num"])
    sel = o["pt_num"] > 3
    lam_o, lam_g = o["eigenvalues"][sel].astype(np.float64), g["eigenvalues"][sel].astype(np.float64)
    scale = lam_o[:, :1] + 1e-12
    assert np.abs(lam_g - lam_o).max() / 1.0 < 1e3 and (np.abs(lam_g - lam_o) / scale).max() < 1e-4
    gap01 = (lam_o[:, 0] - lam_o[:, 1]) / scale[:, 0] > 0.05
    gap12 = (lam_o[:, 1] - lam_o[:, 2]) / scale[:, 0] > 0.05
    dp = np.abs((g["principal"][sel] * o["principal"][sel]).sum(1))
    dn = np.abs((g["normal"][sel] * o["normal"][sel]).sum(1))
    assert dp[gap01].min() > np.cos(1e-3)
    assert dn[gap01 & gap12].min() > np.cos(1e-3)
    assert (g["pt_num"][~sel] <= 3).all()
    if 1 <= k <= 64:
        # list mode: pcl::PCA's float mean / covariance in radiusSearch order on both sides -> identical bits
        for name in ("eigenvalues", "principal", "normal"):
            assert np.array_equal(g[name].view(np.uint32), o[name].view(np.uint32)), name
    ctx.close()
