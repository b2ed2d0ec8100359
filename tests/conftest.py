import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle

    oracle.load()
    return oracle


@pytest.fixture(scope="session")
def small_pair():
    from mulls_b200 import synth

    return synth.make_pair(1000, "small")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def load_golden_pair(path):
    """npz with tgt_c / src_c (n,7) arrays, params fields, init, and the oracle's outputs."""
    from mulls_b200 import abi

    z = np.load(path, allow_pickle=False)
    p = abi.default_params()
    for name, _ in abi.IcpParams._fields_:
        key = "p_" + name
        if key not in z:
            continue
        v = z[key]
        if name in ("used_feature_type", "weight_strategy"):
            setattr(p, name, bytes(v.tobytes()).rstrip(b"\0"))
        elif name == "target_bound":
            p.target_bound[:] = [float(x) for x in v]
        else:
            setattr(p, name, v.item())
    pair = {
        "tgt": [abi.as_aos48(z[f"tgt_{c}"]) for c in range(6)],
        "src": [abi.as_aos48(z[f"src_{c}"]) for c in range(6)],
        "params": p,
        "init_guess": z["init_guess"].astype(np.float64),
    }
    expect = {k[4:]: z[k] for k in z.files if k.startswith("exp_")}
    return pair, expect
