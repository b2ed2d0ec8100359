"""Design study for round 2 (no GPU needed): would 2x2x2 blocks of LEVEL-0 cells (h0 = 0.125 m) be exact and cheaper for
the late-iteration queries of k_search whose seed (previous match) is closer than h0/2?

For every source point of a converged pair the script emulates, in float32 exactly as the kernel computes it, the
level-0 cell of the query, the half of the cell it lies in, the 8 cells of the block, and checks that the nearest
target INSIDE THE BLOCK is the true nearest target (scipy cKDTree on float64, FLANN float distance for the comparison)
whenever the seed distance is <= cover0 = 0.998 * h0 / 2. It also counts the candidate points and live cells a level-0
and a level-1 block present to such a query.

Usage: python scripts/studies/level0_block_study.py            (synthetic C2 pair + the real-data golden pair)
"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mulls_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

H0 = np.float32(0.125)


def flann_d2(p, q):
    d = (p - q).astype(np.float32)
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]


def study(name, tgt, src):
    tgt = tgt.astype(np.float32)
    src = src.astype(np.float32)
    org = (tgt.min(0) - np.float32(1e-3)).astype(np.float32)
    inv = np.float32(1.0) / H0
    # target cells (level 0), float32 arithmetic of k_make_keys
    ct = np.floor((tgt - org) * inv).astype(np.int64)
    key = (ct[:, 0] << 40) | (ct[:, 1] << 20) | ct[:, 2]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    uniq, start, cnt = np.unique(ks, return_index=True, return_counts=True)
    cell = {int(k): (int(s), int(c)) for k, s, c in zip(uniq, start, cnt)}
    tree = cKDTree(tgt.astype(np.float64))
    d_true, j_true = tree.query(src.astype(np.float64))
    cover0 = np.float32(0.998) * H0 * np.float32(0.5)
    sel = np.flatnonzero(d_true <= float(cover0) * 0.999)  # seed = the converged match itself (tightest possible seed)
    fq = (src - org) * inv
    c0 = np.floor(fq).astype(np.int64)
    side = np.where(fq - np.floor(fq) >= np.float32(0.5), 1, -1)
    wrong = 0
    n_pts = 0
    n_live = 0
    for i in sel:
        best, bj = np.float32(np.inf), -1
        p = src[i]
        bound = np.float32(d_true[i]) ** 2 * np.float32(1.0001) + np.float32(1e-12)
        for dx in (0, side[i, 0]):
            for dy in (0, side[i, 1]):
                for dz in (0, side[i, 2]):
                    cc = (c0[i, 0] + dx, c0[i, 1] + dy, c0[i, 2] + dz)
                    lo = org + np.array(cc, np.float32) * H0
                    gap = np.maximum(np.float32(0), np.maximum(lo - p, p - (lo + H0)))
                    if float((gap * gap).sum()) > float(bound):
                        continue  # not live: cannot hold anything closer than the seed
                    ent = cell.get((cc[0] << 40) | (cc[1] << 20) | cc[2])
                    n_live += 1
                    if ent is None:
                        continue
                    idx = order[ent[0]:ent[0] + ent[1]]
                    n_pts += len(idx)
                    d2 = flann_d2(tgt[idx], p[None, :])
                    k = int(np.argmin(d2))
                    if d2[k] < best or (d2[k] == best and idx[k] < bj):
                        best, bj = d2[k], int(idx[k])
        # the block's nearest must be AS CLOSE as the true nearest (ties on the float distance are allowed)
        true_d2 = flann_d2(tgt[j_true[i]][None, :], p[None, :])[0]
        if not (bj >= 0 and best <= true_d2):
            wrong += 1
    print(f"{name}: {len(src)} queries, {len(sel)} ({100.0 * len(sel) / len(src):.1f} %) with the match within "
          f"{float(cover0):.4f} m; level-0 block exact for all but {wrong}; per query {n_live / max(len(sel), 1):.2f} live cells, "
          f"{n_pts / max(len(sel), 1):.1f} candidate points")
    return wrong


def main():
    bad = 0
    pair = synth.make_pair(1001, "c2")
    o, _ = oracle.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"])
    T = np.array(o["T"]).reshape(4, 4)
    for c, nm in ((0, "ground"), (2, "facade"), (1, "pillar")):
        s = pair["src"][c][:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        sub = np.random.default_rng(c).choice(len(s), size=min(len(s), 12000), replace=False)
        bad += study(f"synthetic C2 {nm}", pair["tgt"][c][:, :3], s[sub])
    from conftest import load_golden_pair
    gp, exp = load_golden_pair(os.path.join(ROOT, "tests", "golden", "demo_pair.npz"))
    T = np.array(exp["T"]).reshape(4, 4)
    for c, nm in ((0, "ground"), (2, "facade")):
        s = gp["src"][c][:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        bad += study(f"real demo_pair {nm}", gp["tgt"][c][:, :3], s)
    print("ALL EXACT" if bad == 0 else f"{bad} QUERIES WRONG")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
