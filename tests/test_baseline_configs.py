"""Parity on BASELINE.json's configurations at their stated sizes.

* config 1 — script/run_mulls_reg.sh on demo_data: the sixteen real scans of tests/golden/demo_chain.npz through the
  whole chain raw scan -> CFilter::extract_semantic_pts -> determine_source_target_cloud -> mm_lls_icp
  (test/mulls_reg.cpp:134-195) for the fifteen consecutive pairs and 000000 <-> 000015. CPU: the oracle reproduces
  the committed expectations. GPU: mulls_extract_semantic_pts + mulls_icp_run_batch against them (no oracle call).
* config 3 — 120k-point source against the 600k-point 5-scan map, FULL size, CUDA vs oracle.
* config 5 — 300k-point 128-beam pair, FULL size, unsharded and as two source shards, CUDA vs oracle.
(config 2 at size: tests/test_gpu_parity.py::test_full_size_c2; config 4 = many config-2 pairs: test_batch_*.)"""
import hashlib
import importlib.util
import os
import threading

import numpy as np
import pytest

from mulls_b200 import abi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POSE_TOL_M = 1e-4
POSE_TOL_RAD = 1e-4


def _chain_mod():
    spec = importlib.util.spec_from_file_location("make_golden_chain", os.path.join(ROOT, "tests", "golden", "make_golden_chain.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def chain():
    mod = _chain_mod()
    z = np.load(os.path.join(ROOT, "tests", "golden", "demo_chain.npz"))
    raws = [mod.decode_scan(z[f"scan{k}_dmm"], z[f"scan{k}_i"]) for k in range(mod.N_SCANS)]
    return mod, z, raws


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype=np.float32).tobytes()).digest(), dtype=np.uint8)


def _check_features(mod, z, k, f):
    for i, c in enumerate(mod.CLOUDS):
        assert len(f[c]) == int(z[f"scan{k}_n"][i]), (k, c)
        assert np.array_equal(_sha(f[c]), z[f"scan{k}_sha"][i]), (k, c)


def _check_pair(z, i, res, tr):
    assert [res["code"], res["iters"]] == list(z[f"pair{i}_code_iters"]), i
    n = tr["n_iter"]
    np.testing.assert_array_equal(tr["n_corr"][:n], z[f"pair{i}_trace_n_corr"])
    np.testing.assert_array_equal(tr["n_src"][:n], z[f"pair{i}_trace_n_src"])
    dt, dr = synth.pose_error(res["T"], z[f"pair{i}_T"])
    assert dt <= POSE_TOL_M and dr <= POSE_TOL_RAD, (i, dt, dr)
    np.testing.assert_allclose(res["sigma"], z[f"pair{i}_sigma"], rtol=1e-5)


def test_config1_chain_oracle_reproduces_the_fixture(chain, oracle_mod):
    mod, z, raws = chain
    gp, cp = mod.chain_params()
    feats = [mod.oracle_features(r, gp, cp) for r in raws]
    for k, f in enumerate(feats):
        _check_features(mod, z, k, f)
    for i, (a, b) in enumerate(mod.PAIRS):
        t, s, pair = mod.make_pair(feats, raws, a, b)
        assert [t, s] == list(z[f"pair{i}_ts"])
        res, tr = oracle_mod.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"])
        _check_pair(z, i, res, tr)
    # the consecutive pairs recover the vehicle's forward motion (0.69 .. 0.85 m per scan in this sequence)
    for i in range(15):
        assert 0.6 < abs(float(z[f"pair{i}_T"][0, 3])) < 0.95


@pytest.mark.gpu
def test_config1_chain_on_gpu_matches_the_fixture(chain):
    from mulls_b200.registration import Context

    mod, z, raws = chain
    gp, cp = mod.chain_params()
    ctx = Context(0, len(mod.PAIRS), 70000, 70000)
    feats = []
    for k, r in enumerate(raws):
        f = ctx.extract_semantic_pts(r, 0.0, gp, cp)
        _check_features(mod, z, k, f)
        feats.append(f)
    pairs = []
    for i, (a, b) in enumerate(mod.PAIRS):
        t, s, pair = mod.make_pair(feats, raws, a, b)
        assert [t, s] == list(z[f"pair{i}_ts"])
        pairs.append(pair)
    res, tr = ctx.run_batch(pairs, want_trace=True)
    for i in range(len(pairs)):
        _check_pair(z, i, res[i], tr[i])
    ctx.close()


def _assert_parity(g, gt, o, ot):
    assert g["code"] == o["code"] and g["iters"] == o["iters"]
    assert g["n_corr"] == o["n_corr"] and g["n_src"] == o["n_src"]
    np.testing.assert_array_equal(gt["n_corr"], ot["n_corr"])
    np.testing.assert_array_equal(gt["n_src"], ot["n_src"])
    for i in range(ot["n_iter"]):
        scale = max(np.abs(ot["atpa"][i]).max(), 1e-300)
        np.testing.assert_allclose(gt["atpa"][i], ot["atpa"][i], rtol=0, atol=1e-9 * scale)
        bs = max(np.abs(ot["atpb"][i]).max(), 1e-300)
        np.testing.assert_allclose(gt["atpb"][i], ot["atpb"][i], rtol=0, atol=1e-9 * bs)
    dt, dr = synth.pose_error(g["T"], o["T"])
    assert dt <= POSE_TOL_M and dr <= POSE_TOL_RAD, (dt, dr)
    np.testing.assert_allclose(g["sigma"], o["sigma"], rtol=1e-5)


@pytest.mark.gpu
def test_config3_full_size_scan_to_map(oracle_mod):
    """BASELINE config 3: 120k-point source vs the 600k-point accumulated map (non-identity initial guess)."""
    from mulls_b200.registration import Context

    pair = synth.make_pair(1003, "c3")
    # five 120k-return scans; returns beyond the scene's extent are dropped by the generator (592 017 remain)
    assert sum(len(s) for s in pair["src"]) == 120000 and sum(len(t) for t in pair["tgt"]) > 590000
    ctx = Context(0, 1, 130000, 610000)
    g, gt = ctx.run_batch([pair], want_trace=True)
    o, ot = oracle_mod.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"], threads=16)
    _assert_parity(g[0], gt[0], o, ot)
    dt, dr = synth.pose_error(g[0]["T"], pair["T_gt"])
    assert dt < 0.03 and dr < 3e-3
    ctx.close()


@pytest.fixture(scope="module")
def c5_pair():
    pair = synth.make_pair(1005, "c5")
    # 128 beams x 2344 azimuth steps = 300 032 rays, of which ~263k return (the upper beams leave the scene)
    assert sum(len(s) for s in pair["src"]) > 260000 and sum(len(t) for t in pair["tgt"]) > 260000
    return pair


@pytest.fixture(scope="module")
def c5_oracle(c5_pair, oracle_mod):
    return oracle_mod.icp_run(c5_pair["tgt"], c5_pair["src"], c5_pair["params"], c5_pair["init_guess"], threads=16)


@pytest.mark.gpu
def test_config5_full_size_unsharded(c5_pair, c5_oracle):
    from mulls_b200.registration import Context

    ctx = Context(0, 1, 310000, 310000)
    g, gt = ctx.run_batch([c5_pair], want_trace=True)
    _assert_parity(g[0], gt[0], *c5_oracle)
    ctx.close()


@pytest.mark.gpu
def test_config5_full_size_two_source_shards(c5_pair, c5_oracle):
    """BASELINE config 5's partitioning on ONE device: two contexts hold the full target and one contiguous slice of
    every source class each; the exchange steps (claims: min, counts / per-class sums: sum) are done by a host
    callback over both contexts' buffers, exactly where ncclAllReduce sits in a multi-GPU run."""
    from mulls_b200.dist import run_sharded_local

    res, tr = run_sharded_local(c5_pair, world=2, device=0, max_src=160000, max_tgt=310000, want_trace=True)
    for r in range(2):
        _assert_parity(res[r], tr[r], *c5_oracle)
