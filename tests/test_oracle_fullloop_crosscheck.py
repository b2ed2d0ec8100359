"""A second, independently written implementation of the WHOLE registration loop (SURVEY.md §8c item iii, VERDICT r1
"weak" 1): numpy float32/float64 arithmetic in the reference's operand types and order, scipy's cKDTree for the
neighbours, numpy.linalg for the 6x6 inverse, scipy's Rotation for the rotation angle. It runs every iteration on its
own results — threshold schedule, Huber weights from the fourth iteration on, duplicate check with permanent source
shrinking, convergence rule, posterior sigma and information matrix — and is compared with the C++ oracle's trace in
EVERY iteration: correspondence and source counts exactly, normal equations to 1e-11 of their scale (observed:
1e-15), the solved increment to 1e-8, the final pose to 1e-9.

Agreement means two separately written restatements of cregistration.hpp:1114-1440 / :1701-1967 / :1976-2275 /
:2518-2677 compute the same thing all the way through; it does not pin the reference binary (PARITY UNPINNED)."""
import math

import numpy as np
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation

from mulls_b200 import abi

F32, F64 = np.float32, np.float64
G, PL, F, B, R, V = abi.GROUND, abi.PILLAR, abi.FACADE, abi.BEAM, abi.ROOF, abi.VERTEX


def rigid(cloud, T):
    """pcl::transformPointCloudWithNormals: double products summed left to right, each component cast to float."""
    if len(cloud) == 0:
        return cloud
    out = cloud.copy()
    p, n = cloud[:, 0:3].astype(F64), cloud[:, 4:7].astype(F64)
    for r in range(3):
        out[:, r] = (((T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1]) + T[r, 2] * p[:, 2]) + T[r, 3]).astype(F32)
        out[:, 4 + r] = ((T[r, 0] * n[:, 0] + T[r, 1] * n[:, 1]) + T[r, 2] * n[:, 2]).astype(F32)
    return out


def l2_simple(p, q):
    d = p - q  # float32
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def nearest(tgt, src):
    """k = 1 neighbour under FLANN's float distance, ties to the lower index (cKDTree proposes, float32 decides)."""
    k = min(4, len(tgt))
    _, cand = cKDTree(tgt[:, 0:3].astype(F64)).query(src[:, 0:3].astype(F64), k=k)
    cand = cand.reshape(len(src), k)
    d2 = np.stack([l2_simple(src[:, 0:3], tgt[cand[:, j], 0:3]) for j in range(k)], axis=1)
    order = np.lexsort((cand, d2), axis=1)[:, 0]
    rows = np.arange(len(src))
    return cand[rows, order], d2[rows, order]


def shoot(tgt, src, max_distance):
    """CorrespondenceEstimationNormalShooting, k = 10 (:1732-1737): among the 10 nearest targets the one closest to the
    line through the source point along its normal; dropped if that squared line distance exceeds max_distance (not
    squared); the correspondence carries the squared NN distance of the chosen target."""
    k = min(10, len(tgt))
    _, cand = cKDTree(tgt[:, 0:3].astype(F64)).query(src[:, 0:3].astype(F64), k=k)
    cand = cand.reshape(len(src), k)
    d2 = np.stack([l2_simple(src[:, 0:3], tgt[cand[:, j], 0:3]) for j in range(k)], axis=1)
    order = np.lexsort((cand, d2), axis=1)  # the k-NN list in (distance, index) order: the first minimum wins ties
    cand, d2 = np.take_along_axis(cand, order, 1), np.take_along_axis(d2, order, 1)
    n = src[:, 4:7].astype(F64)
    line = np.empty((len(src), k))
    for j in range(k):
        v = (tgt[cand[:, j], 0:3] - src[:, 0:3]).astype(F64)  # float difference, widened
        c = np.cross(n, v)
        line[:, j] = c[:, 0] * c[:, 0] + (c[:, 1] * c[:, 1] + c[:, 2] * c[:, 2])
    best = np.argmin(line, axis=1)
    rows = np.arange(len(src))
    ok = line[rows, best] <= max_distance
    return np.where(ok, cand[rows, best], -1), d2[rows, best], ok


def correspondences(src, tgt, thre, normal_check, cos_thre, normal_shooting=False):
    """determine_corres (:1701-1835). Returns (shrunk source cloud, source rows, target rows, squared distances) or None."""
    if len(src) < 3 or len(tgt) < 3:
        return None
    max_distance = float(F32(2.5) * F32(thre))
    if normal_shooting:
        j, d2, keep = shoot(tgt, src, max_distance)
    else:
        j, d2 = nearest(tgt, src)
        keep = d2.astype(F64) <= max_distance * max_distance
    s_i, t_i, dd = np.flatnonzero(keep), j[keep], d2[keep]
    if len(src) >= 500:  # duplicate check: the first source in index order keeps the target; the cloud shrinks for good
        _, first = np.unique(t_i, return_index=True)
        first.sort()
        src = src[s_i[first]]
        s_i, t_i, dd = np.arange(len(first)), t_i[first], dd[first]
    ok = dd < F32(thre) * F32(thre)
    s_i, t_i, dd = s_i[ok], t_i[ok], dd[ok]
    if normal_check:
        a, b = src[s_i, 4:7].astype(F64), tgt[t_i, 4:7].astype(F64)
        dot = a[:, 0] * b[:, 0] + (a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2])
        ok = ~(np.abs(dot).astype(F32).astype(F64) < cos_thre)
        s_i, t_i, dd = s_i[ok], t_i[ok], dd[ok]
    return src, s_i, t_i, dd


def w_dist(q, it):
    dist = np.sqrt((q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2])
    b = F32(0.7) + F32(0.05) * F32(it)
    b = b if b < F32(1.3) else F32(1.3)
    w = (F64(b) + (1.0 - F64(b)) * dist.astype(F64) / 30.0).astype(F32)
    return np.where(w.astype(F64) > 0.01, w, F32(0.01))


def w_intensity(pi, qi):
    i1, i2 = (pi.astype(F64) + 0.0001).astype(F32), (qi.astype(F64) + 0.0001).astype(F32)
    ratio = np.abs(i1 - i2) / F32(255.0)
    return np.exp(-1.0 * ratio.astype(F64)).astype(F32)


def w_residual(res, window):
    window = F32(window)
    safe = np.where(res > window, res, F32(1.0))
    huber = ((F32(2.0) * safe * window + F32(-1.0) * (window * window)) / safe) / safe
    return np.where(res > window, huber, F32(1.0)).astype(F32)


def seq_sum(v):
    return float(np.cumsum(v.astype(F64))[-1]) if len(v) else 0.0


LOWER = [(r, c) for c in range(6) for r in range(c, 6)]


def plane_terms(S, T, s_i, t_i, it, weight, dist_w, resid_w, inten_w, window):
    """pt2pl_lls_summation (:2066-2156): float products, double sums."""
    p, q, n = S[s_i, 0:3], T[t_i, 0:3], T[t_i, 4:7]
    a = n[:, 2] * p[:, 1] - n[:, 1] * p[:, 2]
    b = n[:, 0] * p[:, 2] - n[:, 2] * p[:, 0]
    c = n[:, 1] * p[:, 0] - n[:, 0] * p[:, 1]
    d = ((((n[:, 0] * q[:, 0] + n[:, 1] * q[:, 1]) + n[:, 2] * q[:, 2]) - n[:, 0] * p[:, 0]) - n[:, 1] * p[:, 1]) - n[:, 2] * p[:, 2]
    w = np.full(len(p), F32(weight), F32)
    if dist_w:
        w = w * w_dist(q, it)
    if resid_w:
        w = w * w_residual(np.abs(d), window)
    if inten_w:
        w = w * w_intensity(S[s_i, 8], T[t_i, 8])
    cols = [n[:, 0], n[:, 1], n[:, 2], a, b, c]
    A, bv = np.zeros((6, 6)), np.zeros(6)
    for r, cc in LOWER:  # w * (the later coefficient) * (the earlier one) for the mixed terms, as the reference writes them
        first, second = (cols[r], cols[cc]) if r >= 3 and cc < 3 else (cols[cc], cols[r])
        A[r, cc] = seq_sum((w * first) * second)
    for r in range(6):
        bv[r] = seq_sum((w * d) * cols[r])
    return A, bv, w, (cols, d)


def line_terms(S, T, s_i, t_i, it, weight, dist_w, resid_w, inten_w, window):
    """pt2li_lls_pri_direction_summation (:2160-2275): only the diagonal survives the symmetrisation (Q1)."""
    p, q, v = S[s_i, 0:3], T[t_i, 0:3], T[t_i, 4:7]
    px, py, pz, vx, vy, vz = p[:, 0], p[:, 1], p[:, 2], v[:, 0], v[:, 1], v[:, 2]
    dx, dy, dz = px - q[:, 0], py - q[:, 1], pz - q[:, 2]
    z = np.zeros(len(p), F32)
    rows = [[z, -vz, vy, vy * py + vz * pz, -vy * px, -vz * px],
            [vz, z, -vx, -vx * py, vz * pz + vx * px, -vz * py],
            [-vy, vx, z, -vx * pz, -vy * pz, vx * px + vy * py]]
    bb = [-vy * dz + vz * dy, -vz * dx + vx * dz, -vx * dy + vy * dx]
    e = [np.abs(x) for x in bb]
    ed = np.sqrt((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2])
    w = np.full(len(p), F32(weight), F32)
    if dist_w:
        w = w * w_dist(q, it)
    if inten_w:
        w = w * w_intensity(S[s_i, 8], T[t_i, 8])
    if resid_w:
        w = w * w_residual(ed, window)
    sw = np.sqrt(w).astype(F64)
    Am = [[sw * x.astype(F64) for x in row] for row in rows]
    bm = [sw * x.astype(F64) for x in bb]
    A, bv = np.zeros((6, 6)), np.zeros(6)
    for j in range(6):
        A[j, j] = seq_sum(Am[0][j] * Am[0][j] + (Am[1][j] * Am[1][j] + Am[2][j] * Am[2][j]))
        bv[j] = seq_sum(Am[0][j] * bm[0] + (Am[1][j] * bm[1] + Am[2][j] * bm[2]))
    return A, bv, w


def point_terms(S, T, s_i, t_i, dd, it, weight, dist_w, resid_w, inten_w, window):
    """pt2pt_lls_summation (:1976-2063): float products, double sums; no weight is stored back into the correspondence,
    so the posterior reads the squared NN distance that shares its storage (Q2)."""
    p, q = S[s_i, 0:3], T[t_i, 0:3]
    px, py, pz = p[:, 0], p[:, 1], p[:, 2]
    dx, dy, dz = px - q[:, 0], py - q[:, 1], pz - q[:, 2]
    w = np.full(len(p), F32(weight), F32)
    if dist_w:
        w = w * w_dist(q, it)
    if resid_w:
        w = w * w_residual(np.sqrt((dx * dx + dy * dy) + dz * dz), window)
    if inten_w:
        w = w * w_intensity(S[s_i, 8], T[t_i, 8])
    A, bv = np.zeros((6, 6)), np.zeros(6)
    A[0, 0] = A[1, 1] = A[2, 2] = seq_sum(w)
    A[4, 0] = seq_sum(w * pz)
    A[5, 0] = seq_sum(-w * py)
    A[3, 1] = seq_sum(-w * pz)
    A[5, 1] = seq_sum(w * px)
    A[3, 2] = seq_sum(w * py)
    A[4, 2] = seq_sum(-w * px)
    A[3, 3] = seq_sum((w * pz) * pz + (w * py) * py)
    A[4, 3] = seq_sum((-w * px) * py)
    A[5, 3] = seq_sum((-w * px) * pz)
    A[4, 4] = seq_sum((w * pz) * pz + (w * px) * px)
    A[5, 4] = seq_sum((-w * py) * pz)
    A[5, 5] = seq_sum((w * py) * py + (w * px) * px)
    bv[0], bv[1], bv[2] = seq_sum(-w * dx), seq_sum(-w * dy), seq_sum(-w * dz)
    bv[3] = seq_sum((w * pz) * dy - (w * py) * dz)
    bv[4] = seq_sum((w * px) * dz - (w * pz) * dx)
    bv[5] = seq_sum((w * py) * dx - (w * px) * dy)
    return A, bv, dd  # (the "weight" the residual pass will read: the squared NN distance)


def increment_matrix(x):
    tx, ty, tz, al, be, ga = x
    T = np.eye(4)
    T[:3, :3] = (Rotation.from_euler("z", ga) * Rotation.from_euler("y", be) * Rotation.from_euler("x", al)).as_matrix()
    T[:3, 3] = (tx, ty, tz)
    return T


def euler_jacobian(e):
    sr, sp, sy = (F64(F32(math.sin(0.5 * a))) for a in e)
    cr, cp, cy = (F64(F32(math.cos(0.5 * a))) for a in e)
    f = lambda a, b, c: F64(F32(F32(a) * F32(b)) * F32(c))  # float products
    return 0.5 * np.array([[f(cr, cp, cy) + f(sr, sp, sy), -f(sr, sp, cy) - f(cr, cp, sy), -f(sr, cp, sy) - f(cr, sp, cy)],
                           [-f(sr, sp, cy) + f(cr, cp, sy), f(cr, cp, cy) - f(sr, sp, sy), -f(cr, sp, sy) + f(sr, cp, cy)],
                           [-f(sr, cp, sy) - f(cr, sp, cy), -f(cr, sp, sy) - f(sr, cp, cy), f(cr, cp, cy) + f(sr, sp, sy)]])


def undistort(cloud, Tinv):
    """CFilter::apply_motion_compensation (cfilter.hpp:496-516): p <- slerp(I, q(Tinv), s) p + s t(Tinv), s = the timestamp
    ratio kept in `curvature`; ratios outside [0, 1] leave the point alone; coordinates stored back as float."""
    out = cloud.copy()
    s = cloud[:, 9].astype(F64)
    use = ~((cloud[:, 9] < F32(0.0)) | (s > 1.0))
    rotvec = Rotation.from_matrix(Tinv[:3, :3]).as_rotvec()
    Rs = Rotation.from_rotvec(s[use, None] * rotvec[None, :])  # the geodesic from the identity = Eigen's slerp
    moved = Rs.apply(cloud[use, 0:3].astype(F64)) + s[use, None] * Tinv[:3, 3][None, :]
    out[use, 0:3] = moved.astype(F32)
    return out


def run_loop(pair):
    P = pair["params"]
    used = [P.used_feature_type[c:c + 1] == b"1" for c in range(6)]
    ws = P.weight_strategy
    guess = np.array(pair["init_guess"], F64).reshape(4, 4)
    tgt = [t.copy() for t in pair["tgt"]]
    src = [rigid(s, guess) for s in pair["src"]]
    motion_variant = bool(P.apply_motion_undistortion_while_registration)  # :1248-1258
    if P.apply_intersection_filter and not motion_variant:  # :2894-2922, utility.hpp:858-890, cfilter.hpp:950-981
        pts = np.concatenate([src[c][:, 0:3] for c in (G, PL, F)]).astype(F64)
        tb = np.array(P.target_bound[:])
        lo = np.maximum(tb[:3], pts.min(0)) - 1.0
        hi = np.minimum(tb[3:], pts.max(0)) + 1.0
        inside = lambda a: a[((a[:, 0:3].astype(F64) > lo) & (a[:, 0:3].astype(F64) < hi)).all(1)]
        tgt, src = [inside(t) for t in tgt], [inside(s) for s in src]
    n_feature = sum(len(src[c]) for c in (PL, F, B) if used[c])
    thre = [F32(P.dis_thre_unit)] * 6
    cos_thre = math.cos(float(P.normal_bearing) / 180.0 * math.pi)
    max_t = float(F32(2.0 * float(F32(P.dis_thre_unit))))
    conv_r = float(F32(float(P.converge_rotation_d) / 180.0 * math.pi))
    max_r = float(F32(float(P.max_bearable_rotation_d) / 180.0 * math.pi))
    inc, code, log = np.eye(4), 0, []
    sigma2, info, ratio = 1.0, np.eye(6), 1.0
    corr = {}
    for it in range(P.max_iter_num):
        if motion_variant and it == 0:
            # the delivered clouds are undistorted with the inverse initial guess and moved by the initial guess; the vertex
            # cloud is neither undistorted nor re-cloned, so it receives the initial guess a second time (as the reference)
            ginv = np.linalg.inv(guess)
            src = [undistort(pair["src"][c], ginv) for c in range(5)] + [src[V]]
            src = [rigid(s_, guess) for s_ in src]
        else:
            src = [rigid(s, inc) for s in src]
        corr = {}
        for c in range(6):
            if used[c] and len(src[c]) > 0:
                out = correspondences(src[c], tgt[c], thre[c], c != V, cos_thre,
                                      normal_shooting=bool(P.normal_shooting_on) and c in (G, F, R))
                if out is not None:
                    src[c], s_i, t_i, dd = out
                    corr[c] = (s_i, t_i, dd)
        cnt = [len(corr[c][0]) if c in corr else 0 for c in range(6)]
        entry = {"n_corr": cnt, "n_src": [len(s) for s in src]}
        log.append(entry)
        necessary = cnt[PL] + cnt[B] + cnt[F]
        ratio = float(F32(1.0 * necessary / n_feature)) if n_feature else float("inf")
        if sum(cnt) < 40 or necessary < 20 or ratio < float(P.min_neccessary_corr_ratio):
            code, inc = -2, np.eye(4)
            break
        thre = [F32(max(F64(t) / F64(F32(P.dis_thre_update_rate)), F64(F32(P.dis_thre_min)))) for t in thre]
        w_ground = F32(1.0)
        if ws[0:1] == b"1":
            m1, m2, m3, m4 = cnt[G] + cnt[R], cnt[F], cnt[PL], cnt[B]
            w_ground = F32(max(0.01, float(F32(P.z_xy_balanced_ratio) * F32(m2 + 2 * m3 - m4)) / (0.0001 + 2.0 * m1)))
        resid_w, dist_w, inten_w = ws[1:2] == b"1" and it > 2, ws[2:3] == b"1", ws[3:4] == b"1"
        A, b = np.zeros((6, 6)), np.zeros(6)
        weights = {}
        for c, wc in ((G, w_ground), (F, F32(1.0)), (R, w_ground)):
            if c in corr:
                a_, b_, weights[c], _ = plane_terms(src[c], tgt[c], corr[c][0], corr[c][1], it, wc, dist_w, resid_w, inten_w,
                                                    P.pt2pl_residual_window)
                A, b = A + a_, b + b_
        for c in (PL, B):
            if c in corr:
                a_, b_, weights[c] = line_terms(src[c], tgt[c], corr[c][0], corr[c][1], it, 1.0, dist_w, resid_w, inten_w,
                                                P.pt2li_residual_window)
                A, b = A + a_, b + b_
        if V in corr:
            a_, b_, weights[V] = point_terms(src[V], tgt[V], corr[V][0], corr[V][1], corr[V][2], it, 1.0, dist_w, resid_w, inten_w,
                                             P.pt2pt_residual_window)
            A, b = A + a_, b + b_
        A = np.tril(A) + np.tril(A, -1).T
        inv = np.linalg.inv(A)
        x = inv @ b
        entry.update(atpa=A, atpb=b, x=x)
        J = euler_jacobian(x[3:])
        cof = inv.copy()
        cof[3:, 3:] = J @ inv[3:, 3:] @ J.T
        cof[:3, 3:] = inv[:3, 3:] @ J.T
        cof[3:, :3] = J @ inv[3:, :3]
        inc = increment_matrix(x)
        ts, rs = float(np.linalg.norm(inc[:3, 3])), float(Rotation.from_matrix(inc[:3, :3]).magnitude())
        if ts > max_t or abs(rs) > max_r:
            code, inc = -1, np.eye(4)
            break
        if it == P.max_iter_num - 1 or (it > 2 and ts < float(P.converge_translation) and abs(rs) < conv_r):
            vtpv, nobs = 0.0, 0
            for c in (G, F, R):  # :2590-2628
                if c not in corr:
                    continue
                s_i, t_i, _ = corr[c]
                p, q, n = src[c][s_i, 0:3], tgt[c][t_i, 0:3], tgt[c][t_i, 4:7]
                a = n[:, 2] * p[:, 1] - n[:, 1] * p[:, 2]
                bb = n[:, 0] * p[:, 2] - n[:, 2] * p[:, 0]
                cc = n[:, 1] * p[:, 0] - n[:, 0] * p[:, 1]
                d = ((((n[:, 0] * q[:, 0] + n[:, 1] * q[:, 1]) + n[:, 2] * q[:, 2]) - n[:, 0] * p[:, 0]) - n[:, 1] * p[:, 1]) - n[:, 2] * p[:, 2]
                cols = [n[:, 0], n[:, 1], n[:, 2], a, bb, cc]
                r = cols[0].astype(F64) * x[0]
                for k in range(1, 6):
                    r = r + cols[k].astype(F64) * x[k]
                r = (r - d.astype(F64)).astype(F32)
                vtpv += seq_sum((weights[c] * r) * r)
                nobs += len(s_i)
            for c in (PL, B):  # :2631-2677
                if c not in corr:
                    continue
                s_i, t_i, _ = corr[c]
                p, q, v = src[c][s_i, 0:3], tgt[c][t_i, 0:3], tgt[c][t_i, 4:7]
                px, py, pz, vx, vy, vz = p[:, 0], p[:, 1], p[:, 2], v[:, 0], v[:, 1], v[:, 2]
                dx, dy, dz = px - q[:, 0], py - q[:, 1], pz - q[:, 2]
                z = np.zeros(len(p), F32)
                rows = [[z, vz, -vy, -vz * pz - vy * py, vy * px, vz * px],
                        [-vz, z, vx, vx * py, -vx * px - vz * pz, vz * py],
                        [vy, -vx, z, vx * pz, vy * pz, -vy * py - vx * px]]
                bb = [-vz * dy + vy * dz, -vx * dz + vz * dx, -vy * dx + vx * dy]
                tot = np.zeros(len(p))
                for k in range(3):
                    s = rows[k][0].astype(F64) * x[0]
                    for jj in range(1, 6):
                        s = s + rows[k][jj].astype(F64) * x[jj]
                    rk = s - bb[k].astype(F64)
                    tot = tot + rk * rk if k else rk * rk
                vtpv += seq_sum(weights[c].astype(F64) * tot)
                nobs += 3 * len(s_i)
            if V in corr:  # :2546-2588
                s_i, t_i, _ = corr[V]
                p, q = src[V][s_i, 0:3], tgt[V][t_i, 0:3]
                px, py, pz = (p[:, k].astype(F64) for k in range(3))
                d = [(p[:, k] - q[:, k]).astype(F64) for k in range(3)]
                one, zero = np.ones(len(p)), np.zeros(len(p))
                rows = [[one, zero, zero, zero, pz, -py], [zero, one, zero, -pz, zero, px], [zero, zero, one, py, -px, zero]]
                tot = np.zeros(len(p))
                for k in range(3):
                    sk = rows[k][0] * x[0]
                    for jj in range(1, 6):
                        sk = sk + rows[k][jj] * x[jj]
                    rk = sk - (-d[k])
                    tot = tot + rk * rk if k else rk * rk
                vtpv += seq_sum(weights[V].astype(F64) * tot)
                nobs += 3 * len(s_i)
            sigma2 = vtpv / (nobs - 6)
            code = 1 if math.sqrt(sigma2) < float(P.sigma_thre) else -3
            info = (1.0 / sigma2) * np.linalg.inv(cof)
            break
        guess = inc @ guess
    guess = inc @ guess
    return {"T": guess, "code": code, "iters": len(log), "sigma": math.sqrt(sigma2), "info": info, "confidence": ratio,
            "n_corr": log[-1]["n_corr"], "n_src": log[-1]["n_src"]}, log


def check_against_oracle(oracle_mod, pair, min_iters):
    res, tr = oracle_mod.icp_run(pair["tgt"], pair["src"], pair["params"], pair["init_guess"])
    mine, log = run_loop(pair)
    assert mine["code"] == res["code"] and mine["iters"] == res["iters"], (mine["code"], res["code"], mine["iters"], res["iters"])
    assert res["iters"] >= min_iters  # the Huber weights (iteration > 2) and the convergence rule (i > 2) have been exercised
    for it, e in enumerate(log):
        np.testing.assert_array_equal(e["n_corr"], tr["n_corr"][it], err_msg=f"correspondences, iteration {it}")
        np.testing.assert_array_equal(e["n_src"], tr["n_src"][it], err_msg=f"source sizes, iteration {it}")
        if "atpa" not in e:
            continue
        scale_a, scale_b = np.abs(tr["atpa"][it]).max(), np.abs(tr["atpb"][it]).max()
        np.testing.assert_allclose(e["atpa"], np.asarray(tr["atpa"][it]).reshape(6, 6), rtol=0, atol=1e-11 * scale_a, err_msg=f"ATPA, iteration {it}")
        np.testing.assert_allclose(e["atpb"], tr["atpb"][it], rtol=0, atol=1e-11 * scale_b, err_msg=f"ATPb, iteration {it}")
        np.testing.assert_allclose(e["x"], tr["x"][it], rtol=0, atol=1e-8 * max(1.0, np.abs(tr["x"][it]).max()), err_msg=f"x, iteration {it}")
    np.testing.assert_allclose(mine["T"], np.asarray(res["T"]).reshape(4, 4), rtol=0, atol=1e-9)
    np.testing.assert_allclose(mine["sigma"], res["sigma"], rtol=1e-6)
    np.testing.assert_allclose(mine["confidence"], res["confidence"], rtol=1e-6)
    if res["code"] in (1, -3):
        inf = np.asarray(res["info"]).reshape(6, 6)
        np.testing.assert_allclose(mine["info"], inf, rtol=0, atol=1e-6 * np.abs(inf).max())
    return res, log


def test_whole_loop_on_the_synthetic_pair(oracle_mod, small_pair):
    res, log = check_against_oracle(oracle_mod, small_pair, min_iters=5)
    assert res["code"] == 1
    assert log[0]["n_src"] != log[-1]["n_src"]  # the duplicate check did shrink the sources


def test_whole_loop_without_weights_and_without_filter(oracle_mod, small_pair):
    p = abi.IcpParams.from_buffer_copy(small_pair["params"])
    p.weight_strategy = b"0000"
    p.apply_intersection_filter = 0
    check_against_oracle(oracle_mod, dict(small_pair, params=p), min_iters=4)


def test_whole_loop_hits_the_iteration_limit(oracle_mod, small_pair):
    p = abi.IcpParams.from_buffer_copy(small_pair["params"])
    p.max_iter_num = 3
    res, _ = check_against_oracle(oracle_mod, dict(small_pair, params=p), min_iters=3)
    assert res["iters"] == 3


def test_whole_loop_on_real_data(oracle_mod, golden_dir):
    """The decimated demo_data pair (tests/golden/demo_pair.npz): real scans, label-derived classes."""
    import os

    from conftest import load_golden_pair

    pair, _ = load_golden_pair(os.path.join(golden_dir, "demo_pair.npz"))
    check_against_oracle(oracle_mod, pair, min_iters=4)


def test_whole_loop_with_the_point_to_point_metric(oracle_mod, small_pair):
    """vertex class on (off in the shipped configurations): pt2pt terms, and the posterior reading the squared NN distance
    where the other metrics stored a weight"""
    tgt, src = list(small_pair["tgt"]), list(small_pair["src"])
    tgt[V] = small_pair["tgt"][PL][::3].copy()
    src[V] = small_pair["src"][PL][::3].copy()
    p = abi.IcpParams.from_buffer_copy(small_pair["params"])
    p.used_feature_type = b"111111"
    _, log = check_against_oracle(oracle_mod, dict(small_pair, tgt=tgt, src=src, params=p), min_iters=5)
    assert log[0]["n_corr"][V] > 50


def test_whole_loop_with_normal_shooting(oracle_mod, small_pair):
    """normal_shooting_on: ground, facade and roof pick, among their 10 nearest targets, the one closest to the source
    normal's line (cregistration.hpp:1730-1739)"""
    p = abi.IcpParams.from_buffer_copy(small_pair["params"])
    p.normal_shooting_on = 1
    check_against_oracle(oracle_mod, dict(small_pair, params=p), min_iters=4)


def test_whole_loop_with_motion_undistortion(oracle_mod, small_pair):
    """apply_motion_undistortion_while_registration (cregistration.hpp:1248-1258): per-point slerp by the timestamp ratio in
    `curvature` (scipy's rotation-vector scaling here, Eigen's quaternion slerp restated in the oracle), no intersection
    filter, the vertex cloud moved twice"""
    rng = np.random.default_rng(7)
    src = [s_.copy() for s_ in small_pair["src"]]
    for s_ in src:
        s_[:, 9] = rng.uniform(-0.05, 1.05, len(s_)).astype(np.float32)  # a few ratios outside [0, 1]: left untouched
    tgt = list(small_pair["tgt"])
    tgt[V] = small_pair["tgt"][PL][::3].copy()
    src[V] = src[PL][::3].copy()
    p = abi.IcpParams.from_buffer_copy(small_pair["params"])
    p.apply_motion_undistortion_while_registration = 1
    p.used_feature_type = b"111111"
    init = np.eye(4)
    init[:3, :3] = Rotation.from_euler("xyz", [0.002, -0.001, 0.012]).as_matrix()
    init[:3, 3] = (0.9, 0.04, 0.01)
    check_against_oracle(oracle_mod, dict(small_pair, tgt=tgt, src=src, params=p, init_guess=init), min_iters=4)
