"""Study (CPU): work per query of the product's search (search_core.cuh: walk_greedy_seed + nn_search_walk, host instantiation) on the synthetic C2
pair, iteration by iteration, with exactness checked against the oracle's kd-tree NN. The query sets of the
iterations are emulated: sources moved by the oracle's per-iteration increments, shrunk by the duplicate check.
    python scripts/studies/search_stats.py [seed] [config]
"""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mulls_b200 import abi, synth  # noqa: E402
from oracle import oracle  # noqa: E402


def load_harness():
    src = os.path.join(ROOT, "tests", "harness", "search_host.cu")
    out = os.path.join(ROOT, "tests", "harness", "_build", "libsearch_host.so")
    deps = [src, os.path.join(ROOT, "mulls_b200", "csrc", "search_core.cuh"), os.path.join(ROOT, "mulls_b200", "csrc", "grid_key.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-diag-suppress", "20014",
                               "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", out, src])
    lib = C.CDLL(out)
    lib.sh_build.restype = C.c_void_p
    lib.sh_build.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
    lib.sh_free.argtypes = [C.c_void_p]
    lib.sh_grid_info.argtypes = [C.c_void_p, C.c_void_p]
    lib.sh_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_int, C.c_float, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def trans_a(x):
    tx, ty, tz, a, b, g = x
    sa, ca, sb, cb, sg, cg = np.sin(a), np.cos(a), np.sin(b), np.cos(b), np.sin(g), np.cos(g)
    T = np.eye(4)
    T[:3, :3] = [[cg * cb, -sg * ca + cg * sb * sa, sg * sa + cg * sb * ca],
                 [sg * cb, cg * ca + sg * sb * sa, -cg * sa + sg * sb * ca],
                 [-sb, cb * sa, cb * ca]]
    T[:3, 3] = [tx, ty, tz]
    return T


DEFER_FROM = 3


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    config = sys.argv[2] if len(sys.argv) > 2 else "c2"
    leaf = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    global DEFER_FROM
    DEFER_FROM = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    lib = load_harness()
    pair = synth.make_pair(seed, config)
    res, tr = oracle.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"], threads=8)
    n_it = tr["n_iter"]
    print("oracle: code", res["code"], "iters", res["iters"], "n_iter", n_it)
    h0 = 0.125
    P = pair["params"]
    thre = P.dis_thre_unit
    names = ["ground", "pillar", "facade", "beam", "roof", "vertex"]
    tot = {}
    for c in range(6):
        tgt = pair["tgt"][c]
        src = pair["src"][c].copy()
        if len(tgt) < 3 or len(src) < 3 or P.used_feature_type[c:c + 1] != b"1":
            continue
        # (the intersection filter is ignored here: a study, not a parity test)
        mn = tgt[:, :3].min(0)
        origin = (mn - 2 * h0).astype(np.float32)
        rmax = 2.5 * P.dis_thre_unit * 1.0001
        L = 2
        while L < 12 and 0.999 * 0.5 * h0 * (1 << (L - 1)) < rmax:
            L += 1
        pts4 = np.ascontiguousarray(np.concatenate([tgt[:, :3], np.zeros((len(tgt), 1), np.float32)], axis=1))
        G = lib.sh_build(pts4.ctypes.data, len(tgt), float(origin[0]), float(origin[1]), float(origin[2]), h0, L, leaf)
        info = np.zeros(3, np.uint64)
        lib.sh_grid_info(G, info.ctypes.data)
        print(f"[{names[c]}] n_t={len(tgt)} n_s={len(src)} levels={L} cells={info[0]} cap={info[1]} "
              f"mean insert chain={info[2] / max(info[0], 1):.3f}")
        # apply the initial guess
        T0 = pair["init_guess"]
        q = (src[:, :3].astype(np.float64) @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32)
        seeds = None
        th = thre
        for it in range(n_it):
            if it > 0:
                Ti = trans_a(tr["x"][it - 1])
                q = (q.astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
            maxd = 2.5 * th
            r2 = np.float32(np.float32(maxd) * np.float32(maxd)) * np.float32(1.0001)
            m = len(q)
            out_idx = np.empty(m, np.int32)
            out_d2 = np.empty(m, np.float32)
            stats = np.zeros(12, np.uint64)
            stats[10] = 1 if it >= DEFER_FROM else 0
            epq = np.zeros(m, np.uint32)
            qc = np.ascontiguousarray(q)
            t0 = time.time()
            lib.sh_search(G, qc.ctypes.data, seeds.ctypes.data if seeds is not None else None, m, float(r2), 5, 0.25,
                          out_idx.ctypes.data, out_d2.ctypes.data, stats.ctypes.data, epq.ctypes.data)
            dt = time.time() - t0
            # exactness vs the oracle's kd-tree
            srows = np.zeros((m, 12), np.float32)
            srows[:, :3] = q
            oi, od = oracle.nn(tgt, srows, 1e9)
            within = od <= r2 / np.float32(1.0001)
            bad = np.count_nonzero((out_idx != oi) & within)
            badd = np.count_nonzero((out_d2 != od) & within)
            key = (it,)
            a = tot.setdefault(key, np.zeros(13))
            a[:9] += stats[:9].astype(np.float64)
            a[9] += m
            a[10] = max(a[10], float(stats[8]))
            a[11] += bad + badd
            print(f"  it{it}: m={m} probes/q={stats[0] / m:.2f} evals/q={stats[2] / m:.1f} "
                  f"expands/q={stats[3] / m:.2f} levels/q={stats[4] / m:.2f} "
                  f"seed probes/q={stats[6] / m:.2f} seed evals/q={stats[7] / m:.1f} max evals={stats[8]} "
                  f"p50/p90/p99 evals={np.percentile(epq, 50):.0f}/{np.percentile(epq, 90):.0f}/{np.percentile(epq, 99):.0f} "
                  f"mismatch idx={bad} d2={badd} [{dt * 1e6 / m:.2f} us/q host]")
            # emulate determine_corres' shrinking: matched within 2.5*thre and winner of the duplicate check
            matched = (out_idx >= 0) & (out_d2.astype(np.float64) <= float(np.float32(maxd)) ** 2)
            if m >= 500:
                order = np.arange(m)
                first = {}
                keep = np.zeros(m, bool)
                mi = out_idx[matched]
                mo = order[matched]
                _, firstpos = np.unique(mi, return_index=True)
                keep[mo[firstpos]] = True
                q = q[keep]
                seeds = np.ascontiguousarray(out_idx[keep])
            else:
                seeds = np.ascontiguousarray(out_idx)
            th = max(th / P.dis_thre_update_rate, P.dis_thre_min)
        lib.sh_free(G)
    print("== totals per iteration (all classes) ==")
    for (it,), a in sorted(tot.items()):
        m = a[9]
        print(f"it{it}: queries={int(m)} probes/q={a[0] / m:.2f} evals/q={a[2] / m:.1f} expands/q={a[3] / m:.2f} "
              f"levels/q={a[4] / m:.2f} seed probes/q={a[6] / m:.2f} seed evals/q={a[7] / m:.1f} mismatches={int(a[11])}")


if __name__ == "__main__":
    main()
