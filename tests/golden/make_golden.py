"""Generates the golden fixtures in this directory. Run from the repo root IN THE BUILD CONTAINER:

    python tests/golden/make_golden.py

* synth_small.npz  — seeded synthetic pair (mulls_b200.synth, seed 1000, config "small") and the
                     ORACLE's outputs on it (oracle/mulls_oracle.cpp). Inputs are stored too, so the
                     fixture does not depend on numpy's RNG stream staying stable.
* demo_pair_reg.npz — real data, scans 000000 / 000003 with the run_mulls_reg.sh parameter set (see below).
* demo_pair.npz    — real data: /root/reference/demo_data/pcd/000000.pcd (target) and 000001.pcd
                     (source), every 4th point, split into feature classes by the SemanticKITTI label
                     the files carry in `curvature` (ground 40/44/48/49/72, facade 50/51/52,
                     pillar 71/80/81 with a vertical principal direction, roof/beam/vertex empty),
                     normals as stored in the files; plus the ORACLE's outputs.

The reference itself holds no expected outputs for this path: the expected values
are the oracle's, i.e. these fixtures pin the CUDA path (and future oracle edits) to the oracle as
committed — parity with the reference binary stays UNPINNED.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mulls_b200 import abi, synth  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def params_to_npz(p):
    out = {}
    for name, _ in abi.IcpParams._fields_:
        v = getattr(p, name)
        if name in ("used_feature_type", "weight_strategy"):
            out["p_" + name] = np.frombuffer(bytes(v).ljust(8, b"\0"), dtype=np.uint8)
        elif name == "target_bound":
            out["p_" + name] = np.array(list(v), dtype=np.float64)
        else:
            out["p_" + name] = np.array(v)
    return out


def save(path, pair):
    res, tr = oracle.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"], threads=0)
    d = params_to_npz(pair["params"])
    for c in range(6):
        d[f"tgt_{c}"] = pair["tgt"][c][:, [0, 1, 2, 4, 5, 6, 8]]
        d[f"src_{c}"] = pair["src"][c][:, [0, 1, 2, 4, 5, 6, 8]]
    d["init_guess"] = np.asarray(pair["init_guess"], dtype=np.float64)
    d["exp_T"] = res["T"]
    d["exp_info"] = res["info"]
    d["exp_sigma"] = np.float32(res["sigma"])
    d["exp_confidence"] = np.float32(res["confidence"])
    d["exp_code"] = np.int32(res["code"])
    d["exp_iters"] = np.int32(res["iters"])
    d["exp_n_corr"] = np.array(res["n_corr"], dtype=np.uint32)
    d["exp_n_src"] = np.array(res["n_src"], dtype=np.uint32)
    d["exp_trace_n_corr"] = tr["n_corr"]
    d["exp_trace_n_src"] = tr["n_src"]
    d["exp_trace_atpa"] = tr["atpa"]
    d["exp_trace_atpb"] = tr["atpb"]
    d["exp_trace_x"] = tr["x"]
    np.savez_compressed(path, **d)
    print(path, "code", res["code"], "iters", res["iters"], "n_corr", res["n_corr"], os.path.getsize(path) // 1024, "KiB")


def read_pcd(path):
    with open(path, "rb") as f:
        while True:
            line = f.readline()
            if line.startswith(b"DATA"):
                break
        return np.frombuffer(f.read(), dtype=np.float32).reshape(-1, 8).copy()


def demo_cloud(path, step=4):
    a = read_pcd(path)[::step]
    lab = a[:, 7].astype(np.int32)
    xyz, inten, nrm = a[:, 0:3], a[:, 3], a[:, 4:7].copy()
    ok = np.isfinite(nrm).all(1) & (np.abs(np.linalg.norm(nrm, axis=1) - 1.0) < 1e-2)
    groups = {
        abi.GROUND: np.isin(lab, (40, 44, 48, 49, 72)) & ok,
        abi.FACADE: np.isin(lab, (50, 51, 52)) & ok,
        abi.PILLAR: np.isin(lab, (71, 80, 81)),
    }
    out = []
    for c in range(6):
        m = groups.get(c)
        if m is None:
            out.append(np.zeros((0, 7), np.float32))
            continue
        n = nrm[m].copy()
        if c == abi.PILLAR:
            n[:] = (0.0, 0.0, 1.0)
        out.append(np.concatenate([xyz[m], n, inten[m, None]], axis=1).astype(np.float32))
    return out


if __name__ == "__main__":
    save(os.path.join(HERE, "synth_small.npz"), synth.make_pair(1000, "small"))
    ref = "/root/reference/demo_data/pcd"
    if os.path.isdir(ref):
        tgt = demo_cloud(os.path.join(ref, "000000.pcd"))
        src = demo_cloud(os.path.join(ref, "000001.pcd"))
        p = synth.kitti_urban_params(20)
        p.used_feature_type = b"111000"
        p.target_bound[:] = synth.cloud_bound(tgt)
        pair = {"tgt": [abi.as_aos48(t) for t in tgt], "src": [abi.as_aos48(s) for s in src], "params": p,
                "init_guess": np.eye(4)}
        save(os.path.join(HERE, "demo_pair.npz"), pair)
        # the run_mulls_reg.sh parameter set (script/run_mulls_reg.sh:12-46: corr_dis_thre 3.0, 10 iterations, weights
        # "1101", bearing 45) on scans three frames apart, identity initial guess (no TEASER here)
        src3 = demo_cloud(os.path.join(ref, "000003.pcd"))
        q = abi.default_params()
        q.max_iter_num, q.dis_thre_unit, q.dis_thre_min = 10, 3.0, 0.75
        q.converge_translation, q.converge_rotation_d = 0.001, 0.01
        q.used_feature_type = b"111000"
        q.target_bound[:] = synth.cloud_bound(tgt)
        pair = {"tgt": [abi.as_aos48(t) for t in tgt], "src": [abi.as_aos48(s) for s in src3], "params": q,
                "init_guess": np.eye(4)}
        save(os.path.join(HERE, "demo_pair_reg.npz"), pair)
    else:
        print("reference demo_data not present: demo_pair.npz left as committed")
