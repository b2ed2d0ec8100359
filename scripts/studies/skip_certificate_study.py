"""Study (CPU): how many queries of a late iteration can keep their match without a search — the certificate of
search_core.cuh (nn_search_walk's return value) carried from the last full search, the skip rule of k_search:
    sqrt(|p' - q|^2) + |p' - p_ref| (+ margins) < certificate
on the synthetic C2 pair, iteration by iteration (queries moved by the oracle's increments, shrunk by the duplicate check),
with the kept matches checked against the oracle's kd-tree.
    python scripts/studies/skip_certificate_study.py [seed] [config]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import search_stats as S
from mulls_b200 import synth
from oracle import oracle

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
config = sys.argv[2] if len(sys.argv) > 2 else "c2"
lib = S.load_harness()
lib.sh_search_cert.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float, C.c_int, C.c_float, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
pair = synth.make_pair(seed, config)
res, tr = oracle.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"], threads=8)
P = pair["params"]; n_it = tr["n_iter"]; h0 = 0.125
tot = {}
for c in range(5):
    tgt = pair["tgt"][c]; src = pair["src"][c]
    if len(tgt) < 3 or len(src) < 3: continue
    origin = (tgt[:, :3].min(0) - 2 * h0).astype(np.float32)
    rmax = 2.5 * P.dis_thre_unit * 1.0001
    L = 2
    while L < 12 and 0.999 * 0.5 * h0 * (1 << (L - 1)) < rmax: L += 1
    pts4 = np.ascontiguousarray(np.concatenate([tgt[:, :3], np.zeros((len(tgt), 1), np.float32)], axis=1))
    G = lib.sh_build(pts4.ctypes.data, len(tgt), float(origin[0]), float(origin[1]), float(origin[2]), h0, L, 32)
    T0 = pair["init_guess"]
    q = (src[:, :3].astype(np.float64) @ T0[:3, :3].T + T0[:3, 3]).astype(np.float32)
    th = P.dis_thre_unit
    m = len(q)
    prev = np.full(m, -1, np.int32); ref = q.copy(); cert = np.zeros(m, np.float32)
    for it in range(n_it):
        if it > 0:
            Ti = S.trans_a(tr["x"][it - 1]); q = (q.astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
        m = len(q)
        maxd = 2.5 * th
        r2 = np.float32(np.float32(maxd) * np.float32(maxd)) * np.float32(1.0001)
        # skip rule (float32 as the kernel would evaluate it)
        skip = np.zeros(m, bool)
        has = (prev >= 0) & (cert > 0)
        if has.any():
            tq = tgt[prev[has], :3]
            d = q[has] - tq; d1 = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
            e = q[has] - ref[has]; dl = np.sqrt((e[:, 0] * e[:, 0] + e[:, 1] * e[:, 1]) + e[:, 2] * e[:, 2])
            ok = (d1 + dl) * np.float32(1.0001) + np.float32(3e-5) < np.sqrt(cert[has]) * np.float32(0.9999)
            skip[np.flatnonzero(has)[ok]] = True
        need = ~skip
        oi = np.empty(m, np.int32); od = np.empty(m, np.float32); oc = np.empty(m, np.float32); st = np.zeros(12, np.uint64)
        st[10] = 1 if it >= 3 else 0
        qn = np.ascontiguousarray(q[need]); sn = np.ascontiguousarray(prev[need]); k = int(need.sum())
        oin = np.empty(k, np.int32); odn = np.empty(k, np.float32); ocn = np.empty(k, np.float32)
        if k:
            lib.sh_search_cert(G, qn.ctypes.data, sn.ctypes.data, k, float(r2), 5, 0.25, oin.ctypes.data, odn.ctypes.data, st.ctypes.data, None, ocn.ctypes.data)
        new_prev = prev.copy(); new_prev[need] = oin
        new_cert = cert.copy(); new_cert[need] = ocn
        new_ref = ref.copy(); new_ref[need] = q[need]
        # exactness of everything (skipped and searched) vs the oracle kd-tree
        srows = np.zeros((m, 12), np.float32); srows[:, :3] = q
        ki, kd = oracle.nn(tgt, srows, 1e9)
        within = kd <= r2 / np.float32(1.0001)
        bad = np.count_nonzero((new_prev != ki) & within)
        a = tot.setdefault(it, np.zeros(4)); a += [m, skip.sum(), bad, st[2]]
        # shrink
        d = q - tgt[np.maximum(new_prev, 0), :3]
        dd = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float64)
        matched = (new_prev >= 0) & (dd <= float(np.float32(maxd)) ** 2)
        if m >= 500:
            idx = np.flatnonzero(matched); _, first = np.unique(new_prev[idx], return_index=True)
            keep = np.zeros(m, bool); keep[idx[first]] = True
        else:
            keep = np.ones(m, bool)
        q, prev, cert, ref = q[keep], new_prev[keep], new_cert[keep], new_ref[keep]
        th = max(th / P.dis_thre_update_rate, P.dis_thre_min)
    lib.sh_free(G)
for it, a in sorted(tot.items()):
    print(f"it{it}: queries {int(a[0])}, kept without a search {int(a[1])} = {100 * a[1] / a[0]:.1f}%, mismatches vs the oracle {int(a[2])}, evals of the searched ones {int(a[3])}")
