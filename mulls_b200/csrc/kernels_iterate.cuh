// Iteration phase: one pass of the loop body of mm_lls_icp (cregistration.hpp:1239-1401) as four kernels; the first
// three run over (pair, class, 128-source chunk) work items, the fourth over pairs:
//   k_search      I1+I2a  apply the previous increment to the source (:1260), exact radius-bounded 1-NN on the hashed
//                         multi-level grid (replaces the kd-tree query of :1745), claim the target
//   k_resolve     I2b     duplicate check (:1755-1792), distance rejector (:1794-1796), normal check (:1798-1830),
//                         per-class correspondence counts
//   k_accumulate  I3-I5   order-preserving source compaction (:1776-1789), 21+6 normal-equation terms per
//                         correspondence (:1976-2275), fixed-order block reduction -> one partial per chunk
//   k_solve       I5-I9   per pair: partials summed in chunk order, 6x6 solve, Euler/Jacobian, convergence and
//                         status logic (:1301-1400) — the iteration driver lives on the device
// After the loop: k_posterior + k_finalize (:2518-2677, :1386). Source-sharded registrations (k_shard_*) insert the
// caller's all-reduce between these phases.
#pragma once
#include "device_math.cuh"
#include "device_types.cuh"
#include "kernels_ingest.cuh"

namespace mulls {

// kernels recorded into the iteration graph are launched over the context's CAPACITY and take the double-buffer index
// from the device-side loop counter (buf < 0); the host loop passes exact grids and buf = it & 1
__device__ __forceinline__ bool chunk_in_run(const DeviceArrays &A) { return blockIdx.x < (unsigned)A.ctl->n_it_chunks; }
__device__ __forceinline__ int loop_buf(const DeviceArrays &A, int buf) { return buf >= 0 ? buf : (A.ctl->it & 1); }
__device__ __forceinline__ mulls_icp_trace *trace_of(const DeviceArrays &A, uint32_t pair) {
    return (A.trace && A.ctl->trace_on) ? &A.trace[pair] : nullptr;
}

// ------------------------------------------------------------------------------------------------
// exact 1-NN within radius on the multi-level hashed grid of one target class: search_core.cuh
// (__host__ __device__; the CPU suite runs the same functions against a brute-force scan)
// ------------------------------------------------------------------------------------------------
constexpr int kSearchRanges = 8;  // candidate ranges queued per thread between two scans (one block / one split)
constexpr int kSearchStack = 8;   // dense cells waiting to be split (overflow: the cell is scanned whole)

// ---- the warp-cooperative form of the search (see SoloCoop in search_core.cuh for the per-thread semantics) -----
// Traversal stays per lane (a few probes per query); the candidates it queues are examined by the whole warp as ONE
// flat list: per-query candidate counts are broad (median 16, p99 > 100), so a per-lane scan loop keeps ~8 of 32
// lanes busy, while equal shares of the flat list keep all of them busy whatever the split between the queries.
//   ranges    lane l queues its ranges straight into ent[0..nr)[l]; obase[l] = items queued by the lanes before l
//   shares    lane i examines items [i*C, (i+1)*C), C = ceil(T / 32): a 5-step search for the owner of its first
//             item, then the same lean loop a single thread would run, walking from range to range
//   result    per owner a 64-bit key (distance bits << 32 | target) in shared memory. A lane keeps the best of the
//             owner it is working for in registers (starting from the owner's current key, so that most candidates
//             fail one compare) and publishes an improvement when it moves on to another owner: compare-and-swap
//             loop; equal distances are settled by the ORIGINAL index, exactly like consider()
constexpr int kSoloScanLanes = 3; // at most this many lanes with queued ranges: they scan on their own

struct WarpShared {
    uint2 ent[kSearchRanges][32]; // [r][lane] = {start, count}
    uint32_t obase[33];
    uint32_t nr[32];
    float4 qp[32];
    unsigned long long best[32];
};

// per-thread scratch of the search in shared memory: the range queue lives in the warp's table, the stack of dense
// cells is interleaved by thread (conflict-free 8-byte accesses)
struct SmemScratch {
    uint2 *ranges; // &warp.ent[0][lane]
    uint2 *stk;    // &s_stack[0][threadIdx.x]
    __device__ __forceinline__ uint2 &range(int i) { return ranges[i * 32]; }
    __device__ __forceinline__ uint2 &stack(int i) { return stk[i * kIterBlock]; }
};

struct WarpCoop {
    WarpShared *w;
    __device__ __forceinline__ bool any(bool b) { return __any_sync(0xffffffffu, b); }

    __device__ __forceinline__ void offer(const GridView &g, int owner, uint32_t d2bits, int j) {
        unsigned long long old = *(volatile unsigned long long *)&w->best[owner];
        const unsigned long long mine = ((unsigned long long)d2bits << 32) | (unsigned long long)(uint32_t)j;
        while (true) {
            const uint32_t od2 = (uint32_t)(old >> 32);
            const int oj = (int)(uint32_t)old;
            bool win = d2bits < od2;
            if (!win && d2bits == od2 && oj != j)
                win = oj < 0 || __float_as_int(__ldg(&g.nrm[j]).w) < __float_as_int(__ldg(&g.nrm[oj]).w);
            if (!win) return;
            const unsigned long long prev = atomicCAS(&w->best[owner], old, mine);
            if (prev == old) return;
            old = prev;
        }
    }

    template <class Scratch, class Stats>
    __device__ __forceinline__ void scan(const GridView &g, float px, float py, float pz, Scratch &S, int &nr, float &best_d2,
                                         int &best_j, Stats &st) {
        const unsigned full = 0xffffffffu;
        const unsigned queued = __ballot_sync(full, nr > 0); // lanes that queued anything (the owners to walk through)
        if (queued == 0u) return;
        if (__popc(queued) <= kSoloScanLanes) { // a handful of stragglers: sharing their work costs more than it saves
            scan_ranges(g, px, py, pz, S, nr, best_d2, best_j, st);
            return;
        }
        const int lane = threadIdx.x & 31;
        uint32_t cnt = 0;
        for (int r = 0; r < nr; ++r) cnt += S.range(r).y;
        uint32_t incl = cnt; // inclusive prefix sum of the item counts over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t T = __shfl_sync(full, incl, 31);
        w->obase[lane] = incl - cnt;
        w->nr[lane] = (uint32_t)nr;
        w->qp[lane] = make_float4(px, py, pz, 0.0f);
        w->best[lane] = ((unsigned long long)__float_as_uint(best_d2) << 32) | (unsigned long long)(uint32_t)best_j;
        __syncwarp();
        const uint32_t C = (T + 31u) >> 5;
        const uint32_t s0 = (uint32_t)lane * C;
        if (s0 < T) {
            uint32_t n = min(C, T - s0);
            // owner of item s0: the last lane whose base is <= s0 (lanes with nothing queued share their successor's base)
            int owner = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (w->obase[owner + step] <= s0) owner += step;
            uint32_t r = 0, off = w->obase[owner];
            uint2 en = w->ent[0][owner];
            while (off + en.y <= s0) {
                off += en.y;
                en = w->ent[++r][owner];
            }
            const float4 *cur = g.pos + en.x + (s0 - off);
            uint32_t seg_left = en.y - (s0 - off);
            float4 o = w->qp[owner];
            unsigned long long key = *(volatile unsigned long long *)&w->best[owner];
            uint32_t lbd = (uint32_t)(key >> 32);
            int lbj = (int)(uint32_t)key;
            bool improved = false;
            while (true) {
                // the part of the current range that belongs to this lane's share: one tight loop
                const uint32_t m = min(n, seg_left);
                n -= m, seg_left -= m;
#pragma unroll 2
                for (uint32_t i = 0; i < m; ++i) {
                    const float4 q = __ldg(cur);
                    const uint32_t bits = __float_as_uint(flann_l2(o.x, o.y, o.z, q.x, q.y, q.z));
                    if (bits <= lbd) {
                        const int jj = (int)(cur - g.pos);
                        if (bits < lbd) {
                            lbd = bits, lbj = jj, improved = true;
                        } else if (jj != lbj &&
                                   (lbj < 0 || __float_as_int(__ldg(&g.nrm[jj]).w) < __float_as_int(__ldg(&g.nrm[lbj]).w))) {
                            lbj = jj, improved = true;
                        }
                    }
                    ++cur;
                }
                if (n == 0) break;
                // next range: of the same owner, or of the next lane that queued any
                if (++r == w->nr[owner]) {
                    if (improved) offer(g, owner, lbd, lbj);
                    owner = __ffs((int)(queued & (0xfffffffeu << owner))) - 1;
                    r = 0;
                    o = w->qp[owner];
                    key = *(volatile unsigned long long *)&w->best[owner];
                    lbd = (uint32_t)(key >> 32), lbj = (int)(uint32_t)key, improved = false;
                }
                en = w->ent[r][owner];
                cur = g.pos + en.x, seg_left = en.y;
            }
            if (improved) offer(g, owner, lbd, lbj);
        }
        __syncwarp();
        const unsigned long long b = w->best[lane];
        best_d2 = __uint_as_float((uint32_t)(b >> 32));
        best_j = (int)(uint32_t)b;
        nr = 0;
        __syncwarp(); // the next round's writes to ent / qp / best must not overtake these reads
    }
};

__device__ __forceinline__ GridView grid_of(const DeviceArrays &A, const PairConst &pc, const PairState &ps, int c, int leaf_count) {
    GridView g;
    g.table = A.hash + ps.hash_base[c];
    g.mask = ps.hash_mask[c];
    g.pos = A.tgt_pos + pc.tgt_base[c];
    g.nrm = A.tgt_nrm + pc.tgt_base[c];
    g.ox = ps.origin[0], g.oy = ps.origin[1], g.oz = ps.origin[2];
    g.h0 = ps.h0, g.inv_h0 = ps.inv_h0;
    g.n_levels = ps.n_levels;
    g.leaf_count = leaf_count;
    return g;
}

// squared distance from p to the (slightly inflated) box of cell (x,y,z) at a level with cell size hl
__device__ __forceinline__ float cell_dist2(const GridView &g, float px, float py, float pz, float hl, int x, int y,
                                            int z, float margin) {
    const float ax = slab_dist(g.ox + (float)x * hl, g.ox + (float)(x + 1) * hl, px, margin);
    const float ay = slab_dist(g.oy + (float)y * hl, g.oy + (float)(y + 1) * hl, py, margin);
    const float az = slab_dist(g.oz + (float)z * hl, g.oz + (float)(z + 1) * hl, pz, margin);
    return ax * ax + ay * ay + az * az;
}

// ------------------------------------------------------------------------------------------------
// exact k nearest targets (k = 10) for the normal-shooting correspondences of :1732-1737
// (pcl::registration::CorrespondenceEstimationNormalShooting): same hierarchy as nn_search, the pruning bound is
// the current k-th best distance and there is no search radius — the level pyramid of a pair that uses normal
// shooting goes up to a block that spans the whole grid, so the result is exact however far the targets are.
// Total order (d2, original index). A rarely used option: plain per-thread DFS with its stack in local memory,
// kept out of k_search (own kernel, k_search_shoot).
// ------------------------------------------------------------------------------------------------
constexpr int kShootK = 10;
constexpr int kShootStack = 48; // DFS entries: at most 7 stay behind per descended level

struct KnnList {
    float d2[kShootK];
    int j[kShootK];
    int n;
};
__device__ __forceinline__ float knn_bound(const KnnList &kl) { return kl.n < kShootK ? INFINITY : kl.d2[kShootK - 1]; }

__device__ __forceinline__ void knn_insert(const GridView &g, KnnList &kl, float d2, int j) {
    if (kl.n == kShootK) {
        const float w = kl.d2[kShootK - 1];
        if (d2 > w) return;
        if (d2 == w && __float_as_int(__ldg(&g.nrm[j]).w) >= __float_as_int(__ldg(&g.nrm[kl.j[kShootK - 1]]).w)) return;
    }
    for (int i = 0; i < kl.n; ++i)
        if (kl.j[i] == j) return; // a point is met again when the search ascends a level
    int pos = (kl.n < kShootK) ? kl.n : kShootK - 1;
    while (pos > 0) {
        const float dp = kl.d2[pos - 1];
        bool before = d2 < dp;
        if (d2 == dp) before = __float_as_int(__ldg(&g.nrm[j]).w) < __float_as_int(__ldg(&g.nrm[kl.j[pos - 1]]).w);
        if (!before) break;
        kl.d2[pos] = kl.d2[pos - 1];
        kl.j[pos] = kl.j[pos - 1];
        --pos;
    }
    kl.d2[pos] = d2;
    kl.j[pos] = j;
    if (kl.n < kShootK) ++kl.n;
}

__device__ __forceinline__ void knn_search(const GridView &g, float px, float py, float pz, int start_level, KnnList &kl) {
    kl.n = 0;
    const int c0x = (int)floorf((px - g.ox) * g.inv_h0);
    const int c0y = (int)floorf((py - g.oy) * g.inv_h0);
    const int c0z = (int)floorf((pz - g.oz) * g.inv_h0);
    const int L = g.n_levels;
    const float margin = 1e-3f * g.h0;
    uint2 st_cell[kShootStack]; // pack_cell(x, y, z, level, 0)
    float st_d2[kShootStack];
    for (int l = min(max(start_level, 1), L - 1);; ++l) {
        const float H = g.h0 * (float)(1 << l);
        const int ncell = (1 << kCoordBits) >> l;
        for (int k = 0; k < 8; ++k) { // own cell first, then the half-side neighbours
            int x = (c0x >> l) + ((k & 1) ? ((((c0x >> (l - 1)) & 1) ? 1 : -1)) : 0);
            int y = (c0y >> l) + ((k & 2) ? ((((c0y >> (l - 1)) & 1) ? 1 : -1)) : 0);
            int z = (c0z >> l) + ((k & 4) ? ((((c0z >> (l - 1)) & 1) ? 1 : -1)) : 0);
            if (ncell == 2) x = k & 1, y = (k >> 1) & 1, z = k >> 2; // top of the full pyramid: the 8 cells ARE the grid
            if (x < 0 || y < 0 || z < 0 || x >= ncell || y >= ncell || z >= ncell) continue;
            int sp = 0;
            st_cell[0] = pack_cell((uint32_t)x, (uint32_t)y, (uint32_t)z, l, 0u);
            st_d2[0] = cell_dist2(g, px, py, pz, H, x, y, z, margin);
            sp = 1;
            while (sp > 0) {
                --sp;
                if (st_d2[sp] > knn_bound(kl) * 1.0001f + 1e-12f) continue;
                const uint2 ce = st_cell[sp];
                const int lv = (int)((ce.y >> 4) & 0xfu);
                const int cx = (int)(ce.x & 0xfffu), cy = (int)((ce.x >> 12) & 0xfffu), cz = (int)((ce.x >> 24) | ((ce.y & 0xfu) << 8));
                uint32_t start, count, cmask;
                if (!probe_cell(g, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz, lv, start, count, cmask)) continue;
                if (count <= (uint32_t)g.leaf_count || lv == 0 || sp + 8 > kShootStack) {
                    for (uint32_t jj = start; jj < start + count; ++jj) {
                        const float4 q = __ldg(&g.pos[jj]);
                        knn_insert(g, kl, flann_l2(px, py, pz, q.x, q.y, q.z), (int)jj);
                    }
                } else {
                    const float hc = 0.5f * g.h0 * (float)(1 << lv);
                    for (int ch = 7; ch >= 0; --ch) {
                        if (!((cmask >> ch) & 1u)) continue;
                        const int x2 = 2 * cx + (ch & 1), y2 = 2 * cy + ((ch >> 1) & 1), z2 = 2 * cz + (ch >> 2);
                        const float d2c = cell_dist2(g, px, py, pz, hc, x2, y2, z2, margin);
                        if (d2c > knn_bound(kl) * 1.0001f + 1e-12f) continue;
                        st_cell[sp] = pack_cell((uint32_t)x2, (uint32_t)y2, (uint32_t)z2, lv - 1, 0u);
                        st_d2[sp] = d2c;
                        ++sp;
                    }
                }
            }
        }
        const float cover = 0.999f * 0.5f * H; // every target closer than this has been examined
        if (kl.n == kShootK && kl.d2[kShootK - 1] <= cover * cover) break;
        if (l == L - 1) break; // the top block spans the whole grid: everything has been examined
    }
}

// ---- k_search ----------------------------------------------------------------------------------
// what every search kernel does first: cregistration.hpp:1260 — incremental in-place update of the float source
// cloud by the previous iteration's TempTran (double math, float store, as pcl::transformPointCloudWithNormals)
__device__ __forceinline__ void load_and_advance(DeviceArrays &A, const PairState &ps, int buf, uint32_t gi, bool valid,
                                                 float4 &p, float4 &n) {
    p = A.src_pos[buf][gi];
    n = A.src_nrm[buf][gi];
    if (valid && ps.iter > 0) {
        const double *t = ps.T_inc;
        const double px = p.x, py = p.y, pz = p.z, qx = n.x, qy = n.y, qz = n.z;
        p.x = (float)(t[0] * px + t[1] * py + t[2] * pz + t[3]);
        p.y = (float)(t[4] * px + t[5] * py + t[6] * pz + t[7]);
        p.z = (float)(t[8] * px + t[9] * py + t[10] * pz + t[11]);
        n.x = (float)(t[0] * qx + t[1] * qy + t[2] * qz);
        n.y = (float)(t[4] * qx + t[5] * qy + t[6] * qz);
        n.z = (float)(t[8] * qx + t[9] * qy + t[10] * qz);
        A.src_pos[buf][gi] = p;
        A.src_nrm[buf][gi] = n;
    }
}

// shoot = 0: every class except the normal-shooting ones; shoot = 1 (k_search_shoot): only those
__device__ __forceinline__ bool shoots(const PairConst &pc, int c) {
    return pc.normal_shooting && (c == MULLS_GROUND || c == MULLS_FACADE || c == MULLS_ROOF);
}

__global__ void __launch_bounds__(kIterBlock, 10) k_search(DeviceArrays A, int buf, int start_level0, int leaf_count,
                                                          float reseed_cells, int dfs_until) {
    if (!chunk_in_run(A)) return;
    if (buf < 0 && A.ctl->it < dfs_until) return; // (iteration graph) this iteration is k_search_dfs's
    buf = loop_buf(A, buf);
    const ChunkDesc cd = A.it_chunks[blockIdx.x];
    const PairConst &pc = A.pc[cd.pair];
    const PairState &ps = A.ps[cd.pair];
    if (ps.status != kRunning || A.hash_used[1]) return;
    const int c = (int)cd.seg;
    const int ns = ps.n_src[c], nt = ps.n_tgt[c], nsg = ps.n_src_g[c];
    if ((int)cd.first >= ns) return; // block-uniform
    if (shoots(pc, c)) return;       // block-uniform: k_search_shoot's work
    const uint32_t local = cd.first + threadIdx.x;
    const bool valid = (int)local < ns;
    const uint32_t gi = pc.src_base[c] + (valid ? local : 0);
    float4 p, n;
    load_and_advance(A, ps, buf, gi, valid, p, n);
    // determine_corres needs >= 3 points on both sides (:1727-1728)
    if (!(pc.used[c] && nsg >= 3 && nt >= 3)) { // block-uniform
        if (valid) {
            A.nn_idx[gi] = -1;
            A.nn_d2[gi] = INFINITY;
        }
        return;
    }
    const GridView g = grid_of(A, pc, ps, c, leaf_count);
    // CorrespondenceEstimation keeps d2 <= (2.5*thre)^2, evaluated in double (:1745, PCL)
    const float max_distance_f = 2.5f * ps.thre;
    const double max_dist_sqr = (double)max_distance_f * (double)max_distance_f;
    const float r2_prune = (float)max_dist_sqr * 1.0001f;
    // seeds: the previous iteration's match (a real candidate, so the box-distance pruning bites from the first
    // cell on and the search only has to prove that nothing is closer); a stale or missing one is replaced by a
    // short walk through p's own cells when that is closer
    int best_j = -1;
    float best_d2 = INFINITY;
    NoStats st;
    if (valid) {
        const int pj = A.src_prevj[buf][gi];
        if (pj >= 0) {
            const float4 q = __ldg(&g.pos[pj]);
            best_d2 = flann_l2(p.x, p.y, p.z, q.x, q.y, q.z);
            best_j = pj;
        }
    }
    const float rs = reseed_cells * g.h0;
    __shared__ uint2 s_stack[kSearchStack][kIterBlock];
    __shared__ WarpShared s_warp[kIterBlock / 32];
    SmemScratch S{&s_warp[threadIdx.x >> 5].ent[0][threadIdx.x & 31], &s_stack[0][threadIdx.x]};
    WarpCoop co{&s_warp[threadIdx.x >> 5]};
    nn_search<kSearchRanges, kSearchStack>(g, valid, p.x, p.y, p.z, r2_prune, start_level0, rs * rs, best_d2, best_j, S, co, st);
    if (!valid) return;
    if (best_j >= 0 && !((double)best_d2 <= max_dist_sqr)) best_j = -1;
    if (best_j >= 0) {
        // duplicate_check_table as a claim: the lowest source index wins (:1762-1786, Q5)
        atomicMin(&A.claim[pc.tgt_base[c] + best_j], (unsigned)__float_as_int(n.w));
    }
    A.nn_idx[gi] = best_j;
    A.nn_d2[gi] = best_d2;
}

// ---- k_search_dfs: the same search, one independent depth-first walk per thread (nn_search_dfs): no cooperation,
//      no synchronisation, the tightest pruning; the warp-cooperative k_search keeps more lanes busy per instruction.
//      Which one serves an iteration is a tunable (see DESIGN.md for the measurements behind the default).
constexpr int kDfsRanges = 8, kDfsStack = 16;
struct DfsScratch {
    uint2 *base; // &s_scratch[0][threadIdx.x]
    __device__ __forceinline__ uint2 &range(int i) { return base[i * kIterBlock]; }
    __device__ __forceinline__ uint2 &stack(int i) { return base[(kDfsRanges + i) * kIterBlock]; }
};

__global__ void __launch_bounds__(kIterBlock, 9) k_search_dfs(DeviceArrays A, int buf, int start_level0, int leaf_count,
                                                             float reseed_cells, int defer_from_iter, int dfs_until) {
    if (!chunk_in_run(A)) return;
    if (buf < 0 && A.ctl->it >= dfs_until) return; // (iteration graph) this iteration is k_search's
    buf = loop_buf(A, buf);
    const ChunkDesc cd = A.it_chunks[blockIdx.x];
    const PairConst &pc = A.pc[cd.pair];
    const PairState &ps = A.ps[cd.pair];
    if (ps.status != kRunning || A.hash_used[1]) return;
    const int c = (int)cd.seg;
    const int ns = ps.n_src[c], nt = ps.n_tgt[c], nsg = ps.n_src_g[c];
    if ((int)cd.first >= ns) return; // block-uniform
    if (shoots(pc, c)) return;       // block-uniform: k_search_shoot's work
    const uint32_t local = cd.first + threadIdx.x;
    const bool valid = (int)local < ns;
    const uint32_t gi = pc.src_base[c] + (valid ? local : 0);
    float4 p, n;
    load_and_advance(A, ps, buf, gi, valid, p, n);
    if (!valid) return;
    if (!(pc.used[c] && nsg >= 3 && nt >= 3)) {
        A.nn_idx[gi] = -1;
        A.nn_d2[gi] = INFINITY;
        return;
    }
    const GridView g = grid_of(A, pc, ps, c, leaf_count);
    const float max_distance_f = 2.5f * ps.thre;
    const double max_dist_sqr = (double)max_distance_f * (double)max_distance_f;
    const float r2_prune = (float)max_dist_sqr * 1.0001f;
    __shared__ uint2 s_scratch[kDfsRanges + kDfsStack][kIterBlock];
    DfsScratch S{&s_scratch[0][threadIdx.x]};
    NoStats st;
    int best_j = -1;
    float best_d2 = INFINITY;
    {
        const int pj = A.src_prevj[buf][gi];
        if (pj >= 0) {
            const float4 q = __ldg(&g.pos[pj]);
            best_d2 = flann_l2(p.x, p.y, p.z, q.x, q.y, q.z);
            best_j = pj;
        }
        const float rs = reseed_cells * g.h0;
        if (best_j < 0 || best_d2 > rs * rs) {
            uint2 leaf;
            if (quick_locate(g, p.x, p.y, p.z, start_level0, leaf, st)) {
                int nr = 0;
                S.range(nr++) = leaf;
                scan_ranges(g, p.x, p.y, p.z, S, nr, best_d2, best_j, st);
            }
        }
    }
    nn_search_dfs<kDfsRanges, kDfsStack>(g, p.x, p.y, p.z, r2_prune, start_level0, ps.iter >= defer_from_iter, best_d2, best_j, S, st);
    if (best_j >= 0 && !((double)best_d2 <= max_dist_sqr)) best_j = -1;
    if (best_j >= 0) atomicMin(&A.claim[pc.tgt_base[c] + best_j], (unsigned)__float_as_int(n.w));
    A.nn_idx[gi] = best_j;
    A.nn_d2[gi] = best_d2;
}

// ---- k_search_walk: round 1's per-thread walk (depth-first stack of (cell, box distance) in local memory, small
//      cells examined where they are met or queued per block), on the coordinate keys of grid_key.cuh and with
//      level-0 blocks. Kept selectable: the yardstick the other two forms are measured against on the same box.
constexpr int kWalkStack = 48; // DFS entries: at most 7 stay behind per descended level
constexpr int kWalkQueue = 8;  // leaves of one block whose scan is deferred to the end of its traversal

__device__ __forceinline__ void walk_scan_leaf(const GridView &g, float px, float py, float pz, uint32_t start, uint32_t count,
                                               float &best_d2, int &best_j) {
    for (uint32_t jj = start; jj < start + count; ++jj) {
        const float4 q = __ldg(&g.pos[jj]);
        const float d2 = flann_l2(px, py, pz, q.x, q.y, q.z);
        if (d2 < best_d2) {
            best_d2 = d2;
            best_j = (int)jj;
        } else if (d2 == best_d2 && best_j >= 0 && (int)jj != best_j) {
            const int oj = __float_as_int(__ldg(&g.nrm[jj]).w);
            const int ob = __float_as_int(__ldg(&g.nrm[best_j]).w);
            if (oj < ob) best_j = (int)jj;
        }
    }
}

// walk greedily from p's own cell (first level, from `l` upwards, at which it exists) down through the nearest
// existing child to a leaf and take its best point as the seed
__device__ __forceinline__ void walk_greedy_seed(const GridView &g, float px, float py, float pz, int l, float &best_d2, int &best_j) {
    const int c0x = (int)floorf((px - g.ox) * g.inv_h0);
    const int c0y = (int)floorf((py - g.oy) * g.inv_h0);
    const int c0z = (int)floorf((pz - g.oz) * g.inv_h0);
    const int L = g.n_levels;
    l = min(max(l, 1), L - 1);
    for (int lr = l; lr < L && best_j < 0; ++lr) {
        const int ncell = (1 << kCoordBits) >> lr;
        int cx = c0x >> lr, cy = c0y >> lr, cz = c0z >> lr;
        if (!(cx >= 0 && cy >= 0 && cz >= 0 && cx < ncell && cy < ncell && cz < ncell)) break;
        for (int lv = lr;; --lv) {
            uint32_t start, count, cmask;
            if (!probe_cell(g, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz, lv, start, count, cmask)) break; // only possible at lv == lr
            if (count <= (uint32_t)g.leaf_count || lv == 0) {
                for (uint32_t jj = start; jj < start + count; ++jj) {
                    const float4 q = __ldg(&g.pos[jj]);
                    const float d2 = flann_l2(px, py, pz, q.x, q.y, q.z);
                    if (d2 < best_d2) best_d2 = d2, best_j = (int)jj;
                }
                break;
            }
            const float hl = g.h0 * (float)(1 << lv);
            const int ox = (px >= g.ox + ((float)cx + 0.5f) * hl) ? 1 : 0;
            const int oy = (py >= g.oy + ((float)cy + 0.5f) * hl) ? 1 : 0;
            const int oz = (pz >= g.oz + ((float)cz + 0.5f) * hl) ? 1 : 0;
            int ch = ox | (oy << 1) | (oz << 2);
            if (!((cmask >> ch) & 1u)) ch = __ffs((int)cmask) - 1; // any existing child still yields a valid seed
            if (ch < 0) break;
            cx = 2 * cx + (ch & 1), cy = 2 * cy + ((ch >> 1) & 1), cz = 2 * cz + (ch >> 2);
        }
    }
}

__device__ __forceinline__ float walk_axis_dist(float o, float H, int x, float p, float margin) {
    const float lo = o + (float)x * H - margin, hi = o + (float)(x + 1) * H + margin;
    return fmaxf(0.0f, fmaxf(lo - p, p - hi));
}

// stack storage of the walk: thread-local arrays (local memory, L1-cached) or shared memory interleaved by thread
struct WalkStackLocal {
    uint32_t cell[kWalkStack], meta[kWalkStack];
    float d2[kWalkStack];
    __device__ __forceinline__ void put(int i, uint32_t c, uint32_t m, float d) { cell[i] = c, meta[i] = m, d2[i] = d; }
    __device__ __forceinline__ float dist(int i) const { return d2[i]; }
    __device__ __forceinline__ void get(int i, uint32_t &c, uint32_t &m) const { c = cell[i], m = meta[i]; }
};
constexpr int kWalkSmemStack = 16;
struct WalkStackSmem { // uint4-free: three word arrays, [entry][thread]
    uint32_t *base;    // &s_walk[0][threadIdx.x]; entry i: words (3*i .. 3*i+2) * kIterBlock
    __device__ __forceinline__ void put(int i, uint32_t c, uint32_t m, float d) {
        base[(3 * i) * kIterBlock] = c, base[(3 * i + 1) * kIterBlock] = m, base[(3 * i + 2) * kIterBlock] = __float_as_uint(d);
    }
    __device__ __forceinline__ float dist(int i) const { return __uint_as_float(base[(3 * i + 2) * kIterBlock]); }
    __device__ __forceinline__ void get(int i, uint32_t &c, uint32_t &m) const { c = base[(3 * i) * kIterBlock], m = base[(3 * i + 1) * kIterBlock]; }
};

template <class Stack, int kDepth>
__device__ __forceinline__ void nn_search_walk(const GridView &g, float px, float py, float pz, float r2_prune, int start_level,
                                               bool defer_scan, float &best_d2, int &best_j, Stack &stk) {
    const float fx = (px - g.ox) * g.inv_h0, fy = (py - g.oy) * g.inv_h0, fz = (pz - g.oz) * g.inv_h0;
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    const int c0x = (int)flx, c0y = (int)fly, c0z = (int)flz;
    const int L = g.n_levels;
    const float margin = 1e-3f * g.h0;
    // stack entry: cell = x | y << 12 | (z & 0xff) << 24, meta = z >> 8 | level << 4, and the cell's box distance
    uint32_t q_start[kWalkQueue], q_count[kWalkQueue];
    int nq = 0;
    int l = min(max(start_level, 1), L - 1);
    if (best_j >= 0) { // seeded: the smallest level whose coverage reaches the seed (level 0: 0.998 * h0 / 2)
        const float need = 1.001f * sqrtf(best_d2);
        const float t = need / (0.999f * 0.5f * g.h0);
        if (t <= 1.0f) l = (need <= 0.998f * 0.5f * g.h0) ? 0 : 1;
        else l = ilogbf(t) + 1;
        l = min(l, L - 1);
    }
    for (;; ++l) {
        const float H = g.h0 * (float)(1 << l);
        const int ncell = (1 << kCoordBits) >> l;
        int xs[2], ys[2], zs[2];
        xs[0] = c0x >> l, ys[0] = c0y >> l, zs[0] = c0z >> l;
        if (l == 0) {
            xs[1] = xs[0] + (((fx - flx) >= 0.5f) ? 1 : -1);
            ys[1] = ys[0] + (((fy - fly) >= 0.5f) ? 1 : -1);
            zs[1] = zs[0] + (((fz - flz) >= 0.5f) ? 1 : -1);
        } else {
            xs[1] = xs[0] + (((c0x >> (l - 1)) & 1) ? 1 : -1);
            ys[1] = ys[0] + (((c0y >> (l - 1)) & 1) ? 1 : -1);
            zs[1] = zs[0] + (((c0z >> (l - 1)) & 1) ? 1 : -1);
        }
        float ex[2], ey[2], ez[2];
        bool vx[2], vy[2], vz[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            vx[i] = xs[i] >= 0 && xs[i] < ncell;
            vy[i] = ys[i] >= 0 && ys[i] < ncell;
            vz[i] = zs[i] >= 0 && zs[i] < ncell;
            ex[i] = walk_axis_dist(g.ox, H, xs[i], px, margin);
            ey[i] = walk_axis_dist(g.oy, H, ys[i], py, margin);
            ez[i] = walk_axis_dist(g.oz, H, zs[i], pz, margin);
            ex[i] *= ex[i], ey[i] *= ey[i], ez[i] *= ez[i];
        }
        uint32_t live = 0;
        {
            const float bound0 = fminf(best_d2, r2_prune) * 1.0001f + 1e-12f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = k & 1, j = (k >> 1) & 1, m = k >> 2;
                if (vx[i] && vy[j] && vz[m] && ex[i] + ey[j] + ez[m] <= bound0) live |= 1u << k;
            }
        }
#pragma unroll 1
        while (live) { // lowest bit first: k = 0 is p's own cell
            const int k = __ffs(live) - 1;
            live &= live - 1;
            const int i = k & 1, j = (k >> 1) & 1, m = k >> 2;
            int sp = 0;
            stk.put(0, cell_key_lo((uint32_t)xs[i], (uint32_t)ys[j], (uint32_t)zs[m]), ((uint32_t)zs[m] >> 8) | ((uint32_t)l << 4),
                    ex[i] + ey[j] + ez[m]);
            sp = 1;
            while (sp > 0) {
                --sp;
                // a cell farther than the best so far (or than the radius) cannot change the result
                if (stk.dist(sp) > fminf(best_d2, r2_prune) * 1.0001f + 1e-12f) continue;
                uint32_t cell, meta;
                stk.get(sp, cell, meta);
                const int lv = (int)((meta >> 4) & 0xfu);
                const int cx = (int)(cell & 0xfffu), cy = (int)((cell >> 12) & 0xfffu), cz = (int)((cell >> 24) | ((meta & 0xfu) << 8));
                uint32_t start, count, cmask;
                if (!probe_cell(g, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz, lv, start, count, cmask)) continue;
                if (count <= (uint32_t)g.leaf_count || lv == 0 || sp + 8 > kDepth) {
                    if (defer_scan && nq < kWalkQueue) { // scanned together with the block's other leaves
                        q_start[nq] = start;
                        q_count[nq] = count;
                        ++nq;
                    } else {
                        walk_scan_leaf(g, px, py, pz, start, count, best_d2, best_j);
                    }
                } else {
                    const float hc = 0.5f * g.h0 * (float)(1 << lv);
                    float ax[2], ay[2], az[2];
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        ax[b] = walk_axis_dist(g.ox, hc, 2 * cx + b, px, margin);
                        ay[b] = walk_axis_dist(g.oy, hc, 2 * cy + b, py, margin);
                        az[b] = walk_axis_dist(g.oz, hc, 2 * cz + b, pz, margin);
                        ax[b] *= ax[b], ay[b] *= ay[b], az[b] *= az[b];
                    }
                    const int near_child = (ax[1] < ax[0] ? 1 : 0) | (ay[1] < ay[0] ? 2 : 0) | (az[1] < az[0] ? 4 : 0);
                    const float bound = fminf(best_d2, r2_prune) * 1.0001f + 1e-12f;
                    uint32_t pass = 0;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch)
                        if (ax[ch & 1] + ay[(ch >> 1) & 1] + az[ch >> 2] <= bound) pass |= 1u << ch;
                    pass &= cmask;
                    if (near_child & 1) pass = ((pass & 0x55u) << 1) | ((pass & 0xaau) >> 1);
                    if (near_child & 2) pass = ((pass & 0x33u) << 2) | ((pass & 0xccu) >> 2);
                    if (near_child & 4) pass = ((pass & 0x0fu) << 4) | ((pass & 0xf0u) >> 4);
                    while (pass) {
                        const int c = 31 - __clz((int)pass);
                        pass ^= 1u << c;
                        const int ch = c ^ near_child;
                        const uint32_t x2 = (uint32_t)(2 * cx + (ch & 1)), y2 = (uint32_t)(2 * cy + ((ch >> 1) & 1)), z2 = (uint32_t)(2 * cz + (ch >> 2));
                        stk.put(sp, cell_key_lo(x2, y2, z2), (z2 >> 8) | ((uint32_t)(lv - 1) << 4), ax[ch & 1] + ay[(ch >> 1) & 1] + az[ch >> 2]);
                        ++sp;
                    }
                }
            }
        }
        for (int qi = 0; qi < nq; ++qi) walk_scan_leaf(g, px, py, pz, q_start[qi], q_count[qi], best_d2, best_j);
        nq = 0;
        const float cover = (l == 0) ? 0.998f * 0.5f * g.h0 : 0.999f * 0.5f * H;
        const float cover2 = cover * cover;
        if (best_d2 <= cover2) break;
        if (cover2 >= r2_prune) break;
        if (l == L - 1) break;
    }
}

template <int kMinBlocks, bool kSmem>
__global__ void __launch_bounds__(kIterBlock, kMinBlocks) k_search_walk(DeviceArrays A, int buf, int start_level0, int leaf_count,
                                                                       int defer_from_iter, float reseed_cells) {
    if (!chunk_in_run(A)) return;
    buf = loop_buf(A, buf);
    const ChunkDesc cd = A.it_chunks[blockIdx.x];
    const PairConst &pc = A.pc[cd.pair];
    const PairState &ps = A.ps[cd.pair];
    if (ps.status != kRunning || A.hash_used[1]) return;
    const int c = (int)cd.seg;
    const int ns = ps.n_src[c], nt = ps.n_tgt[c], nsg = ps.n_src_g[c];
    if ((int)cd.first >= ns) return;
    if (shoots(pc, c)) return;
    const uint32_t local = cd.first + threadIdx.x;
    const bool valid = (int)local < ns;
    const uint32_t gi = pc.src_base[c] + (valid ? local : 0);
    float4 p, n;
    load_and_advance(A, ps, buf, gi, valid, p, n);
    if (!valid) return;
    if (!(pc.used[c] && nsg >= 3 && nt >= 3)) {
        A.nn_idx[gi] = -1;
        A.nn_d2[gi] = INFINITY;
        return;
    }
    const GridView g = grid_of(A, pc, ps, c, leaf_count);
    const float max_distance_f = 2.5f * ps.thre;
    const double max_dist_sqr = (double)max_distance_f * (double)max_distance_f;
    const float r2_prune = (float)max_dist_sqr * 1.0001f;
    int best_j = -1;
    float best_d2 = INFINITY;
    const int pj = A.src_prevj[buf][gi];
    if (pj >= 0) {
        const float4 q = __ldg(&g.pos[pj]);
        best_d2 = flann_l2(p.x, p.y, p.z, q.x, q.y, q.z);
        best_j = pj;
    }
    {
        const float rs = reseed_cells * g.h0; // a match that the last increment left far away is challenged by a fresh seed
        if (best_j < 0 || best_d2 > rs * rs) {
            float d2 = INFINITY;
            int j = -1;
            walk_greedy_seed(g, p.x, p.y, p.z, start_level0, d2, j);
            if (j >= 0 && d2 < best_d2) best_d2 = d2, best_j = j;
        }
    }
    if (kSmem) {
        __shared__ uint32_t s_walk[3 * kWalkSmemStack][kIterBlock];
        WalkStackSmem stk{&s_walk[0][threadIdx.x]};
        nn_search_walk<WalkStackSmem, kWalkSmemStack>(g, p.x, p.y, p.z, r2_prune, start_level0, ps.iter >= defer_from_iter, best_d2, best_j, stk);
    } else {
        WalkStackLocal stk;
        nn_search_walk<WalkStackLocal, kWalkStack>(g, p.x, p.y, p.z, r2_prune, start_level0, ps.iter >= defer_from_iter, best_d2, best_j, stk);
    }
    if (best_j >= 0 && !((double)best_d2 <= max_dist_sqr)) best_j = -1;
    if (best_j >= 0) atomicMin(&A.claim[pc.tgt_base[c] + best_j], (unsigned)__float_as_int(n.w));
    A.nn_idx[gi] = best_j;
    A.nn_d2[gi] = best_d2;
}

// :1732-1737 normal shooting [PCL CorrespondenceEstimationNormalShooting, k = 10]: among the 10 nearest targets
// the one with the smallest squared distance to the line through the source point along its normal; dropped
// if that value exceeds max_distance (NOT squared); correspondence distance = its squared NN distance.
// Launched only when a pair of the batch asked for normal shooting.
__global__ void __launch_bounds__(kIterBlock) k_search_shoot(DeviceArrays A, int buf, int start_level0, int leaf_count) {
    if (!chunk_in_run(A)) return;
    buf = loop_buf(A, buf);
    const ChunkDesc cd = A.it_chunks[blockIdx.x];
    const PairConst &pc = A.pc[cd.pair];
    const PairState &ps = A.ps[cd.pair];
    if (ps.status != kRunning || A.hash_used[1]) return;
    const int c = (int)cd.seg;
    const int ns = ps.n_src[c], nt = ps.n_tgt[c], nsg = ps.n_src_g[c];
    if ((int)cd.first >= ns || !shoots(pc, c)) return; // block-uniform
    const uint32_t local = cd.first + threadIdx.x;
    const bool valid = (int)local < ns;
    const uint32_t gi = pc.src_base[c] + (valid ? local : 0);
    float4 p, n;
    load_and_advance(A, ps, buf, gi, valid, p, n);
    if (!valid) return;
    if (!(pc.used[c] && nsg >= 3 && nt >= 3)) {
        A.nn_idx[gi] = -1;
        A.nn_d2[gi] = INFINITY;
        return;
    }
    const GridView g = grid_of(A, pc, ps, c, leaf_count);
    const float max_distance_f = 2.5f * ps.thre;
    int sj = -1;
    float sd2 = INFINITY;
    KnnList kl;
    knn_search(g, p.x, p.y, p.z, start_level0, kl);
    double min_dist = 1.7976931348623157e308;
    for (int t = 0; t < kl.n; ++t) {
        const float4 q = __ldg(&g.pos[kl.j[t]]);
        const float ptx = q.x - p.x, pty = q.y - p.y, ptz = q.z - p.z;
        const double Nx = n.x, Ny = n.y, Nz = n.z, Vx = ptx, Vy = pty, Vz = ptz;
        const double Cx = Ny * Vz - Nz * Vy, Cy = Nz * Vx - Nx * Vz, Cz = Nx * Vy - Ny * Vx;
        const double dist = Cx * Cx + (Cy * Cy + Cz * Cz);
        if (dist < min_dist) {
            min_dist = dist;
            sj = kl.j[t];
            sd2 = kl.d2[t];
        }
    }
    if (sj >= 0 && min_dist > (double)max_distance_f) sj = -1;
    if (sj >= 0) atomicMin(&A.claim[pc.tgt_base[c] + sj], (unsigned)__float_as_int(n.w));
    A.nn_idx[gi] = sj;
    A.nn_d2[gi] = sd2;
}

// ---- k_resolve ---------------------------------------------------------------------------------
// kFused: the body runs inside k_finish, where data produced by OTHER blocks of the SAME launch is read — such loads
// bypass L1 (ld.global.cg): a line cached by an earlier block of this SM may predate the producer's store.
template <bool kFused, typename T>
__device__ __forceinline__ T ld_x(const T *p) {
    if (kFused) return __ldcg(p);
    return *p;
}

template <bool kFused>
__device__ __forceinline__ void resolve_body(DeviceArrays &A, int buf, uint32_t chunk) {
    const ChunkDesc cd = A.it_chunks[chunk];
    const PairConst &pc = A.pc[cd.pair];
    PairState &ps = A.ps[cd.pair];
    if (ps.status != kRunning) return;
    const int c = (int)cd.seg;
    const int ns = ps.n_src[c], nt = ps.n_tgt[c], nsg = ps.n_src_g[c];
    const uint32_t local = cd.first + threadIdx.x;
    const bool valid = (int)local < ns;
    const bool active = pc.used[c] && nsg >= 3 && nt >= 3; // determine_corres ran for this class
    const bool dedup = active && nsg >= kDedupMinSrc;
    bool kept = false, pass = false;
    if (valid) {
        const uint32_t gi = pc.src_base[c] + local;
        const int j = A.nn_idx[gi];
        const bool matched = active && j >= 0;
        bool corr = matched;
        kept = true;
        if (dedup) {
            const float4 n = A.src_nrm[buf][gi];
            const bool winner = matched && A.claim[pc.tgt_base[c] + j] == (unsigned)__float_as_int(n.w);
            kept = winner;
            corr = winner;
        }
        if (corr) {
            // CorrespondenceRejectorDistance: distance < thre*thre, both float (:1794-1796, PCL)
            const float d2 = A.nn_d2[gi];
            pass = d2 < ps.thre * ps.thre;
            if (pass && c != MULLS_VERTEX) {
                const float4 n = A.src_nrm[buf][gi];
                const float4 m = A.tgt_nrm[pc.tgt_base[c] + j];
                const double dot = (double)n.x * (double)m.x + (double)n.y * (double)m.y + (double)n.z * (double)m.z;
                const float cos_angle = (float)fabs(dot);
                if ((double)cos_angle < pc.cos_thre) pass = false;
            }
        }
        A.flags[gi] = (uint8_t)((kept ? 1 : 0) | (pass ? 2 : 0));
    }
    const unsigned kb = __ballot_sync(0xffffffffu, kept);
    const unsigned pb = __ballot_sync(0xffffffffu, pass);
    __shared__ unsigned s_kept[kIterBlock / 32], s_pass[kIterBlock / 32];
    if ((threadIdx.x & 31) == 0) {
        s_kept[threadIdx.x >> 5] = __popc(kb);
        s_pass[threadIdx.x >> 5] = __popc(pb);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned k = 0, p = 0;
        for (int w = 0; w < kIterBlock / 32; ++w) {
            k += s_kept[w];
            p += s_pass[w];
        }
        A.blk_kept[chunk] = k;
        if (p) atomicAdd(&ps.n_corr[c], p);
    }
}
__global__ void __launch_bounds__(kIterBlock) k_resolve(DeviceArrays A, int buf) {
    if (chunk_in_run(A)) resolve_body<false>(A, loop_buf(A, buf), blockIdx.x);
}

// ------------------------------------------------------------------------------------------------
// per-correspondence normal-equation terms. Layout of the kTerms doubles of a partial:
//   [0..20]  lower triangle of ATPA, column by column: (0,0)(1,0)..(5,0)(1,1)(2,1)..(5,5)
//   [21..26] ATPb
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void terms_pt2pl(const float4 p, const float pi, const float4 q, const float4 qn,
                                            float weight, int iter_num, bool dist_w, bool resid_w, bool inten_w,
                                            float window, double *t, float &w_out) {
    // cregistration.hpp:2080-2151
    const float px = p.x, py = p.y, pz = p.z, qx = q.x, qy = q.y, qz = q.z;
    const float ntx = qn.x, nty = qn.y, ntz = qn.z;
    float w = weight;
    const float a = ntz * py - nty * pz;
    const float b = ntx * pz - ntz * px;
    const float c = nty * px - ntx * py;
    const float d = ntx * qx + nty * qy + ntz * qz - ntx * px - nty * py - ntz * pz;
    const float dist = sqrtf(qx * qx + qy * qy + qz * qz);
    if (dist_w) w = w * weight_by_dist_adaptive(dist, iter_num);
    if (resid_w) w = w * weight_by_residual(fabsf(d), window);
    if (inten_w) w = w * weight_by_intensity((float)((double)pi + 0.0001), (float)((double)q.w + 0.0001));
    w_out = w;
    t[0] = w * ntx * ntx;
    t[1] = w * ntx * nty;
    t[2] = w * ntx * ntz;
    t[3] = w * a * ntx;
    t[4] = w * b * ntx;
    t[5] = w * c * ntx;
    t[6] = w * nty * nty;
    t[7] = w * nty * ntz;
    t[8] = w * a * nty;
    t[9] = w * b * nty;
    t[10] = w * c * nty;
    t[11] = w * ntz * ntz;
    t[12] = w * a * ntz;
    t[13] = w * b * ntz;
    t[14] = w * c * ntz;
    t[15] = w * a * a;
    t[16] = w * a * b;
    t[17] = w * a * c;
    t[18] = w * b * b;
    t[19] = w * b * c;
    t[20] = w * c * c;
    t[21] = w * d * ntx;
    t[22] = w * d * nty;
    t[23] = w * d * ntz;
    t[24] = w * d * a;
    t[25] = w * d * b;
    t[26] = w * d * c;
}

// diagonal index of column j in the lower-triangle layout
__device__ __forceinline__ int diag_index(int j) {
    const int d[6] = {0, 6, 11, 15, 18, 20};
    return d[j];
}

__device__ __forceinline__ void terms_pt2li(const float4 p, const float pi, const float4 q, const float4 qv,
                                            float weight, int iter_num, bool dist_w, bool resid_w, bool inten_w,
                                            float window, double *t, float &w_out) {
    // cregistration.hpp:2174-2271; only the diagonal of this block survives the symmetrisation (Q1)
    const float px = p.x, py = p.y, pz = p.z, qx = q.x, qy = q.y, qz = q.z;
    const float vx = qv.x, vy = qv.y, vz = qv.z;
    const float dx = px - qx, dy = py - qy, dz = pz - qz;
    double Am[3][6], bv[3];
    Am[0][0] = 0;
    Am[0][1] = (double)(-vz);
    Am[0][2] = (double)vy;
    Am[0][3] = (double)(vy * py + vz * pz);
    Am[0][4] = (double)(-vy * px);
    Am[0][5] = (double)(-vz * px);
    Am[1][0] = (double)vz;
    Am[1][1] = 0;
    Am[1][2] = (double)(-vx);
    Am[1][3] = (double)(-vx * py);
    Am[1][4] = (double)(vz * pz + vx * px);
    Am[1][5] = (double)(-vz * py);
    Am[2][0] = (double)(-vy);
    Am[2][1] = (double)vx;
    Am[2][2] = 0;
    Am[2][3] = (double)(-vx * pz);
    Am[2][4] = (double)(-vy * pz);
    Am[2][5] = (double)(vx * px + vy * py);
    bv[0] = (double)(-vy * dz + vz * dy);
    bv[1] = (double)(-vz * dx + vx * dz);
    bv[2] = (double)(-vx * dy + vy * dx);
    const float ex = (float)fabs(bv[0]), ey = (float)fabs(bv[1]), ez = (float)fabs(bv[2]);
    const float ed = sqrtf(ex * ex + ey * ey + ez * ez);
    float wx = weight;
    const float dist = sqrtf(qx * qx + qy * qy + qz * qz);
    if (dist_w) wx = wx * weight_by_dist_adaptive(dist, iter_num);
    if (inten_w) wx = wx * weight_by_intensity((float)((double)pi + 0.0001), (float)((double)q.w + 0.0001));
    if (resid_w) wx = wx * weight_by_residual(ed, window);
    w_out = wx;
    const double sw = (double)sqrtf(wx);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) Am[r][cc] = sw * Am[r][cc];
        bv[r] = sw * bv[r];
    }
#pragma unroll
    for (int k = 0; k < 21; ++k) t[k] = 0.0;
    const int dg[6] = {0, 6, 11, 15, 18, 20};
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        t[dg[j]] = Am[0][j] * Am[0][j] + (Am[1][j] * Am[1][j] + Am[2][j] * Am[2][j]);
        t[21 + j] = Am[0][j] * bv[0] + (Am[1][j] * bv[1] + Am[2][j] * bv[2]);
    }
}

__device__ __forceinline__ void terms_pt2pt(const float4 p, const float pi, const float4 q, float weight,
                                            int iter_num, bool dist_w, bool resid_w, bool inten_w, float window,
                                            double *t) {
    // cregistration.hpp:1991-2058
    const float px = p.x, py = p.y, pz = p.z, qx = q.x, qy = q.y, qz = q.z;
    const float dx = px - qx, dy = py - qy, dz = pz - qz;
    float wx = weight;
    const float dist = sqrtf(qx * qx + qy * qy + qz * qz);
    if (dist_w) wx = wx * weight_by_dist_adaptive(dist, iter_num);
    if (resid_w) wx = wx * weight_by_residual(sqrtf(dx * dx + dy * dy + dz * dz), window);
    if (inten_w) wx = wx * weight_by_intensity((float)((double)pi + 0.0001), (float)((double)q.w + 0.0001));
    const float wy = wx, wz = wx;
#pragma unroll
    for (int k = 0; k < 27; ++k) t[k] = 0.0;
    t[0] = wx;
    t[4] = wx * pz;
    t[5] = (-wx * py);
    t[6] = wy;
    t[8] = (-wy * pz);
    t[10] = wy * px;
    t[11] = wz;
    t[12] = wz * py;
    t[13] = (-wz * px);
    t[15] = wy * pz * pz + wz * py * py;
    t[16] = (-wz * px * py);
    t[17] = (-wy * px * pz);
    t[18] = wx * pz * pz + wz * px * px;
    t[19] = (-wx * py * pz);
    t[20] = wx * py * py + wy * px * px;
    t[21] = (-wx * dx);
    t[22] = (-wy * dy);
    t[23] = (-wz * dz);
    t[24] = wy * pz * dy - wz * py * dz;
    t[25] = wz * px * dz - wx * pz * dx;
    t[26] = wx * py * dx - wy * px * dy;
}

// w_ground of cregistration.hpp:1892-1900 from the per-class correspondence counts
__device__ __forceinline__ float balanced_ground_weight(const PairConst &pc, const uint32_t *n_corr) {
    if (!pc.w_balance) return 1.0f;
    const int m1 = (int)(n_corr[MULLS_GROUND] + n_corr[MULLS_ROOF]);
    const int m2 = (int)n_corr[MULLS_FACADE], m3 = (int)n_corr[MULLS_PILLAR], m4 = (int)n_corr[MULLS_BEAM];
    const float num = pc.z_xy_ratio * (float)(m2 + 2 * m3 - m4);
    const double v = (double)num / (0.0001 + 2.0 * (double)m1);
    return (float)((0.01 > v) ? 0.01 : v);
}

// :1301-1305 — too few correspondences?
__device__ __forceinline__ bool too_few(const PairConst &pc, const PairState &ps, const uint32_t *n_corr, float &ratio) {
    int total = 0;
    for (int c = 0; c < kNumClasses; ++c) total += (int)n_corr[c];
    const int nec = (int)(n_corr[MULLS_PILLAR] + n_corr[MULLS_BEAM] + n_corr[MULLS_FACADE]);
    ratio = (float)(1.0 * (double)nec / (double)ps.source_feature_points_count);
    return total < 40 || nec < 20 || ratio < pc.min_ratio;
}

// A pair stops iterating: publish the number of pairs still running to the host's launch loop.
__device__ __forceinline__ void pair_left_running(DeviceArrays &A) {
    const int left = atomicSub(A.running, 1) - 1;
    *A.h_running = left;
    __threadfence_system();
}

// Solve + state update of one pair; executed by thread 0 of the last block of k_accumulate
// (cregistration.hpp:1301-1400 after the summations). S = per-class sums [6][kTerms] in shared memory.
__device__ __noinline__ void solve_and_advance(DeviceArrays &A, uint32_t pair, const double *S, double *sm /*>= 150 doubles*/,
                                  int buf_written) {
    const PairConst &pc = A.pc[pair];
    PairState &ps = A.ps[pair];
    const int i = ps.iter;
    ps.iters_entered = i + 1;
    for (int c = 0; c < kNumClasses; ++c) ps.n_corr_last[c] = ps.n_corr[c];
    // bytes touched by this iteration's correspondence search: 28 B per active source and target point
    {
        uint64_t pts = 0;
        for (int c = 0; c < kNumClasses; ++c)
            if (pc.used[c]) pts += (uint64_t)ps.n_tgt[c];
        ps.alg_bytes += 28ull * pts; // sources added by the caller of this function (pre-compaction counts)
    }
    mulls_icp_trace *tr = trace_of(A, pair);
    if (tr && i < MULLS_MAX_TRACE_ITERS) {
        tr->n_iter = i + 1;
        for (int c = 0; c < kNumClasses; ++c) tr->n_corr[i][c] = ps.n_corr[c];
        for (int k = 0; k < 36; ++k) tr->atpa[i][k] = 0.0;
        for (int k = 0; k < 6; ++k) tr->atpb[i][k] = tr->x[i][k] = 0.0;
    }
    float ratio;
    const bool few = too_few(pc, ps, ps.n_corr, ratio);
    ps.confidence = ratio;
    if (few) {
        ps.code = -2;
        ps.status = kDone;
        pair_left_running(A);
        return; // TempTran = identity: T_total stays (:1307-1310, :1403)
    }
    // :1314-1315 threshold update
    {
        const double t = 1.0 * (double)ps.thre / (double)pc.thre_rate;
        ps.thre = (t > (double)pc.thre_min) ? (float)t : pc.thre_min;
    }
    // ATPA/ATPb: classes in the order of :1914-1921 (ground, facade, roof, pillar, beam, vertex)
    double *ATPA = sm;       // 36
    double *ATPb = sm + 36;  // 6
    double *inv = sm + 42;   // 36
    double *lu = sm + 78;    // 36
    double *Tmp = sm + 114;  // 16
    double low[21];
    for (int k = 0; k < 21; ++k) low[k] = 0.0;
    for (int k = 0; k < 6; ++k) ATPb[k] = 0.0;
    const int order[6] = {MULLS_GROUND, MULLS_FACADE, MULLS_ROOF, MULLS_PILLAR, MULLS_BEAM, MULLS_VERTEX};
    for (int o = 0; o < 6; ++o) {
        const double *s = S + order[o] * kTerms;
        for (int k = 0; k < 21; ++k) low[k] += s[k];
        for (int k = 0; k < 6; ++k) ATPb[k] += s[21 + k];
    }
    {
        int k = 0;
        for (int col = 0; col < 6; ++col)
            for (int row = col; row < 6; ++row, ++k) {
                ATPA[6 * row + col] = low[k];
                ATPA[6 * col + row] = low[k]; // :1924-1938 lower -> upper
            }
    }
    inverse6(ATPA, inv, lu);
    double x[6];
    for (int r = 0; r < 6; ++r) {
        double s = inv[6 * r] * ATPb[0];
        for (int cc = 1; cc < 6; ++cc) s = s + inv[6 * r + cc] * ATPb[cc];
        x[r] = s;
        ps.x[r] = s;
    }
    if (tr && i < MULLS_MAX_TRACE_ITERS) {
        for (int k = 0; k < 36; ++k) tr->atpa[i][k] = ATPA[k];
        for (int k = 0; k < 6; ++k) {
            tr->atpb[i][k] = ATPb[k];
            tr->x[i][k] = x[k];
        }
    }
    // :1953-1964 cofactor with the Euler->quaternion Jacobian (half-angle sines/cosines in FLOAT, :2797)
    {
        const float sr = (float)sin(0.5 * x[3]), sp = (float)sin(0.5 * x[4]), sy = (float)sin(0.5 * x[5]);
        const float cr = (float)cos(0.5 * x[3]), cp = (float)cos(0.5 * x[4]), cy = (float)cos(0.5 * x[5]);
        double J[3][3];
        J[0][0] = 0.5 * (double)(cr * cp * cy + sr * sp * sy);
        J[0][1] = 0.5 * (double)(-sr * sp * cy - cr * cp * sy);
        J[0][2] = 0.5 * (double)(-sr * cp * sy - cr * sp * cy);
        J[1][0] = 0.5 * (double)(-sr * sp * cy + cr * cp * sy);
        J[1][1] = 0.5 * (double)(cr * cp * cy - sr * sp * sy);
        J[1][2] = 0.5 * (double)(-cr * sp * sy + sr * cp * cy);
        J[2][0] = 0.5 * (double)(-sr * cp * sy - cr * sp * cy);
        J[2][1] = 0.5 * (double)(-cr * sp * sy - sr * cp * cy);
        J[2][2] = 0.5 * (double)(cr * cp * cy + sr * sp * sy);
        double *cof = ps.cofactor;
        for (int k = 0; k < 36; ++k) cof[k] = inv[k];
        double tmp[3][3];
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc)
                tmp[r][cc] = J[r][0] * inv[6 * 3 + 3 + cc] + J[r][1] * inv[6 * 4 + 3 + cc] + J[r][2] * inv[6 * 5 + 3 + cc];
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc)
                cof[6 * (3 + r) + 3 + cc] = tmp[r][0] * J[cc][0] + tmp[r][1] * J[cc][1] + tmp[r][2] * J[cc][2];
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc)
                cof[6 * r + 3 + cc] = inv[6 * r + 3] * J[cc][0] + inv[6 * r + 4] * J[cc][1] + inv[6 * r + 5] * J[cc][2];
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc)
                cof[6 * (3 + r) + cc] = J[r][0] * inv[6 * 3 + cc] + J[r][1] * inv[6 * 4 + cc] + J[r][2] * inv[6 * 5 + cc];
    }
    // :1333 TempTran
    double *Tt = ps.T_inc;
    construct_trans_a(x, Tt);
    const double ts_norm = sqrt(Tt[3] * Tt[3] + Tt[7] * Tt[7] + Tt[11] * Tt[11]);
    const double rs_angle = fabs(rotation_angle(Tt));
    if (ts_norm > (double)pc.max_t || rs_angle > (double)pc.max_r) { // :1348-1354
        ps.code = -1;
        ps.status = kDone;
        pair_left_running(A);
        return;
    }
    // :1400 / :1403 — the increment is always folded into the accumulated transform
    mat4_mul(Tt, ps.T_total, Tmp);
    for (int k = 0; k < 16; ++k) ps.T_total[k] = Tmp[k];
    if (i == pc.max_iter - 1 || (i > 2 && ts_norm < (double)pc.conv_t && rs_angle < (double)pc.conv_r)) { // :1357
        ps.status = kNeedPosterior;
        ps.final_buf = buf_written;
        pair_left_running(A);
        return;
    }
    ps.iter = i + 1;
}

// ---- k_accumulate ------------------------------------------------------------------------------
template <bool kFused>
__device__ __forceinline__ void accumulate_body(DeviceArrays &A, int buf, uint32_t chunk) {
    const ChunkDesc cd = A.it_chunks[chunk];
    const PairConst &pc = A.pc[cd.pair];
    PairState &ps = A.ps[cd.pair];
    if (ps.status != kRunning) return;
    const int c = (int)cd.seg;
    const int ns = ps.n_src[c];
    const uint32_t local = cd.first + threadIdx.x;
    const bool valid = (int)local < ns;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kIterBlock / 32;
    __shared__ double s_red[kWarps][kTerms];
    __shared__ uint32_t s_off[kWarps + 1];
    __shared__ uint32_t s_base;

    // blocks entirely past the live part of the class have nothing to contribute (k_solve skips them)
    if ((int)cd.first >= ns) return;
    uint32_t dst_local = 0;
    uint8_t fl = 0;
    uint32_t gi = 0;
    bool kept = false, pass = false;
    double t[32];
    // (1) destination of the kept sources: blocks before this one in the same (pair, class)
    {
        uint32_t acc = 0;
        const uint32_t first_chunk = pc.class_chunk_begin[c];
        for (uint32_t b = first_chunk + threadIdx.x; b < chunk; b += kIterBlock) acc += ld_x<kFused>(&A.blk_kept[b]);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_off[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tt = 0;
            for (int w = 0; w < kWarps; ++w) tt += s_off[w];
            s_base = tt;
        }
        __syncthreads();
    }
    if (valid) {
        gi = pc.src_base[c] + local;
        fl = ld_x<kFused>(&A.flags[gi]);
    }
    kept = (fl & 1) != 0, pass = (fl & 2) != 0;
    const unsigned kb = __ballot_sync(0xffffffffu, kept);
    if (lane == 0) s_off[warp] = __popc(kb);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t tt = s_off[w];
            s_off[w] = run;
            run += tt;
        }
        s_off[kWarps] = run;
    }
    __syncthreads();
    dst_local = s_base + s_off[warp] + __popc(kb & ((1u << lane) - 1u));

    // (2) terms of the surviving correspondences
#pragma unroll
    for (int k = 0; k < 32; ++k) t[k] = 0.0;
    float w_store = 0.0f;
    int j = -1;
    float4 p = make_float4(0, 0, 0, 0), n = make_float4(0, 0, 0, 0);
    float d2 = 0.0f;
    if (valid) {
        p = A.src_pos[buf][gi];
        n = A.src_nrm[buf][gi];
        j = A.nn_idx[gi];
        d2 = A.nn_d2[gi];
        if (j >= 0) A.claim[pc.tgt_base[c] + j] = kClaimFree; // reset the table for the next iteration
    }
    uint32_t n_corr[kNumClasses]; // complete since every k_resolve block of the pair has finished
#pragma unroll
    for (int k = 0; k < kNumClasses; ++k) n_corr[k] = ld_x<kFused>(&ps.n_corr[k]);
    float ratio_unused;
    const bool few = too_few(pc, ps, n_corr, ratio_unused);
    if (pass && !few) {
        const float4 q = A.tgt_pos[pc.tgt_base[c] + j];
        const float4 qn = A.tgt_nrm[pc.tgt_base[c] + j];
        const int it = ps.iter;
        const bool resid_w = pc.w_residual && it > 2; // :1905-1907
        const bool dist_w = pc.w_dist != 0, inten_w = pc.w_intensity != 0;
        if (c == MULLS_GROUND || c == MULLS_FACADE || c == MULLS_ROOF) {
            const float wc = (c == MULLS_FACADE) ? 1.0f : balanced_ground_weight(pc, n_corr);
            terms_pt2pl(p, p.w, q, qn, wc, it, dist_w, resid_w, inten_w, pc.win_pt2pl, t, w_store);
        } else if (c == MULLS_PILLAR || c == MULLS_BEAM) {
            terms_pt2li(p, p.w, q, qn, 1.0f, it, dist_w, resid_w, inten_w, pc.win_pt2li, t, w_store);
        } else {
            terms_pt2pt(p, p.w, q, 1.0f, it, dist_w, resid_w, inten_w, pc.win_pt2pt, t);
            w_store = d2; // pt2pt never stores a weight: the posterior reads the squared NN distance (Q2)
        }
    }
    // (3) compaction into the other buffer (order preserved: :1776-1789)
    if (kept) {
        const uint32_t gd = pc.src_base[c] + dst_local;
        A.src_pos[buf ^ 1][gd] = p;
        A.src_nrm[buf ^ 1][gd] = n;
        A.src_prevj[buf ^ 1][gd] = j;
        A.corr_j[gd] = pass ? j : -1;
        A.corr_w[gd] = w_store;
    }
    // (4) block reduction in a fixed order. Warp level: reduce-scatter butterfly — at every step a lane
    // keeps half of the terms and receives the partner's sums of that half, so after 5 steps lane i holds
    // the warp total of term i (31 shuffles instead of 27 x 5).
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const double send = upper ? t[k] : t[k + half];
            const double keep = upper ? t[k + half] : t[k];
            t[k] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    if (lane < kTerms) s_red[warp][lane] = t[0];
    __syncthreads();
    if (threadIdx.x < 27) {
        double v = 0.0;
        for (int w = 0; w < kWarps; ++w) v += s_red[w][threadIdx.x];
        A.partials[(size_t)chunk * kTerms + threadIdx.x] = v;
    }
}
__global__ void __launch_bounds__(kIterBlock) k_accumulate(DeviceArrays A, int buf) {
    if (chunk_in_run(A)) accumulate_body<false>(A, loop_buf(A, buf), blockIdx.x);
}

// ---- solve: one block per pair, after every k_accumulate block of the pair. Sums the per-chunk partials of every
//      class in chunk order (fixed order => bit-reproducible), then one thread solves and advances the pair state.
//      Any block size that is a multiple of 32: the warps take the classes in turn.
template <bool kFused>
__device__ __forceinline__ void solve_body(DeviceArrays &A, int buf, uint32_t pair) {
    const PairConst &pc = A.pc[pair];
    PairState &ps = A.ps[pair];
    if (ps.status != kRunning) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    __shared__ double s_S[kNumClasses][kTerms];
    __shared__ double s_scratch[160];
    __shared__ int s_newn[kNumClasses];
    __shared__ uint32_t s_ncorr[kNumClasses];
    for (int cc = warp; cc < kNumClasses; cc += n_warps) {
        const uint32_t b0 = pc.class_chunk_begin[cc];
        // only the chunks that held live sources this iteration wrote a partial
        const uint32_t live = (uint32_t)((ps.n_src[cc] + kIterBlock - 1) / kIterBlock);
        const uint32_t b1 = min(pc.class_chunk_begin[cc + 1], b0 + live);
        // lane = term; chunks in order, four independent accumulators combined in a fixed order
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (lane < 27) {
            uint32_t b = b0;
            for (; b + 4 <= b1; b += 4) {
                a0 += ld_x<kFused>(&A.partials[(size_t)(b + 0) * kTerms + lane]);
                a1 += ld_x<kFused>(&A.partials[(size_t)(b + 1) * kTerms + lane]);
                a2 += ld_x<kFused>(&A.partials[(size_t)(b + 2) * kTerms + lane]);
                a3 += ld_x<kFused>(&A.partials[(size_t)(b + 3) * kTerms + lane]);
            }
            for (; b < b1; ++b) a0 += ld_x<kFused>(&A.partials[(size_t)b * kTerms + lane]);
        }
        if (lane < kTerms) s_S[cc][lane] = (lane < 27) ? ((a0 + a1) + (a2 + a3)) : 0.0;
        uint32_t acc = 0; // kept sources of the class = its new size
        for (uint32_t b = b0 + lane; b < b1; b += 32) acc += ld_x<kFused>(&A.blk_kept[b]);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            s_newn[cc] = (int)acc;
            s_ncorr[cc] = ld_x<kFused>(&ps.n_corr[cc]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t srcpts = 0;
        for (int cc = 0; cc < kNumClasses; ++cc)
            if (pc.used[cc]) srcpts += (uint64_t)ps.n_src_g[cc];
        ps.alg_bytes += 28ull * srcpts;
        if (pc.sharded) {
            // source-sharded registration: publish this rank's per-class sums; the all-reduce and
            // k_shard_solve (identical on every rank) finish the iteration
            for (int cc = 0; cc < kNumClasses; ++cc) {
                ps.n_src[cc] = s_newn[cc];
                for (int k = 0; k < kTerms; ++k) A.xch_f64[cc * kTerms + k] = (k < 27) ? s_S[cc][k] : 0.0;
            }
            return;
        }
        mulls_icp_trace *tr = trace_of(A, pair);
        for (int cc = 0; cc < kNumClasses; ++cc) {
            ps.n_src[cc] = s_newn[cc]; // classes that skipped determine_corres keep everything (k_resolve)
            ps.n_src_g[cc] = s_newn[cc];
            ps.n_corr[cc] = s_ncorr[cc];
            if (tr && ps.iter < MULLS_MAX_TRACE_ITERS) tr->n_src[ps.iter][cc] = (uint32_t)ps.n_src[cc];
        }
        solve_and_advance(A, pair, &s_S[0][0], s_scratch, buf ^ 1);
        for (int cc = 0; cc < kNumClasses; ++cc) ps.n_corr[cc] = 0;
    }
}
constexpr int kSolveThreads = kNumClasses * 32; // one warp per feature class
// loop_handle != 0: the launch is the last kernel of the iteration graph's WHILE body — the block that finishes last
// advances the loop counter and tells the graph whether another iteration is needed (pairs still running)
__global__ void __launch_bounds__(kSolveThreads) k_solve(DeviceArrays A, int buf, unsigned long long loop_handle) {
    LoopCtl &ctl = *A.ctl;
    if (blockIdx.x >= (unsigned)ctl.n_pairs) return;
    solve_body<false>(A, loop_buf(A, buf), blockIdx.x);
    if (loop_handle == 0ull) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&ctl.solved, 1u) == (unsigned)ctl.n_pairs - 1u) {
            ctl.solved = 0;
            const int it = ctl.it + 1;
            ctl.it = it;
            __threadfence();
            const bool again = *(volatile int *)A.running > 0 && it < ctl.max_iter;
            cudaGraphSetConditional((cudaGraphConditionalHandle)loop_handle, again ? 1u : 0u);
        }
    }
}

// ---- k_finish: everything of one ICP iteration after the search, in ONE launch (2 x chunks blocks). Blocks take
//      tickets: the first `n_chunks` tickets resolve a chunk (duplicate check, rejectors, counts), the next `n_chunks`
//      accumulate a chunk — after waiting for every resolve block of THEIR PAIR (the class weights and the
//      compaction offsets need the pair's complete counts) — and the block that finishes a pair's last chunk solves
//      its 6x6 system and advances its state. A waiting block only ever waits for tickets handed out BEFORE its own,
//      i.e. for blocks that are already running and never wait themselves: no deadlock, whatever the residency.
//      Pairs progress independently: one pair's accumulation overlaps another's resolution.
struct FinishSync {
    unsigned ticket, done, stuck, _pad;
};
__global__ void __launch_bounds__(kIterBlock) k_finish(DeviceArrays A, int buf, uint32_t n_chunks) {
    __shared__ uint32_t s_ticket;
    __shared__ int s_last;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&A.fsync->ticket, 1u);
    __syncthreads();
    const uint32_t t = s_ticket;
    const bool second = t >= n_chunks;
    const uint32_t chunk = second ? t - n_chunks : t;
    const uint32_t pair = A.it_chunks[chunk].pair;
    const uint32_t pair_chunks = A.pc[pair].chunk_end - A.pc[pair].chunk_begin;
    unsigned *resolved = &A.pair_sync[2 * pair], *accumulated = &A.pair_sync[2 * pair + 1];
    if (!second) {
        resolve_body<true>(A, buf, chunk);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(resolved, 1u);
        }
    } else {
        if (threadIdx.x == 0) {
            unsigned spins = 0;
            while (*(volatile unsigned *)resolved < pair_chunks) {
                __nanosleep(64);
                if (++spins > (1u << 24)) { // (cannot happen; never hang the device on a logic error)
                    atomicExch(&A.fsync->stuck, 1u);
                    break;
                }
            }
            __threadfence();
        }
        __syncthreads();
        accumulate_body<true>(A, buf, chunk);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = atomicAdd(accumulated, 1u) == pair_chunks - 1;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            solve_body<true>(A, buf, pair);
            if (threadIdx.x == 0) *resolved = 0, *accumulated = 0; // every block of the pair is past its wait
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&A.fsync->done, 1u) == 2 * n_chunks - 1) A.fsync->ticket = 0, A.fsync->done = 0;
    }
}

// ---- sharded mode (mulls_icp_run_sharded, BASELINE config 5): one pair, the target replicated, the source
//      classes split over ranks in contiguous index ranges. Three tiny exchange steps per iteration, each an
//      all-reduce supplied by the caller (NCCL): claim table (min), counts (sum), per-class sums (sum).
// after ingest: global class sizes and the global bbox of source ground/pillar/facade
__global__ void k_shard_pack_setup(DeviceArrays A, int phase) {
    PairState &ps = A.ps[0];
    if (phase == 0) { // bbox: min over [min_xyz, -max_xyz] in the ordered-int encoding
        for (int d = 0; d < 3; ++d) {
            A.xch_i32[d] = ps.bb_src[d];
            A.xch_i32[3 + d] = ~ps.bb_src[3 + d]; // max(x) = ~min(~x), no overflow for INT_MIN
        }
    } else if (phase == 1) {
        for (int d = 0; d < 3; ++d) {
            ps.bb_src[d] = A.xch_i32[d];
            ps.bb_src[3 + d] = ~A.xch_i32[3 + d];
        }
    } else if (phase == 2) {
        for (int c = 0; c < kNumClasses; ++c) A.xch_i32[c] = ps.n_src[c];
    } else {
        const PairConst &pc = A.pc[0];
        int cnt = 0;
        for (int c = 0; c < kNumClasses; ++c) ps.n_src_g[c] = A.xch_i32[c];
        if (pc.used[MULLS_PILLAR]) cnt += ps.n_src_g[MULLS_PILLAR];
        if (pc.used[MULLS_FACADE]) cnt += ps.n_src_g[MULLS_FACADE];
        if (pc.used[MULLS_BEAM]) cnt += ps.n_src_g[MULLS_BEAM];
        ps.source_feature_points_count = cnt;
    }
}
// after k_resolve: this rank's correspondence and kept-source counts -> exchange buffer; and back
__global__ void __launch_bounds__(kIterBlock) k_shard_counts(DeviceArrays A, int phase) {
    const PairConst &pc = A.pc[0];
    PairState &ps = A.ps[0];
    if (ps.status != kRunning) {
        if (phase == 0 && threadIdx.x < 2 * kNumClasses) A.xch_i32[threadIdx.x] = 0;
        return;
    }
    if (phase == 0) {
        __shared__ uint32_t s_w[kIterBlock / 32];
        for (int cc = 0; cc < kNumClasses; ++cc) {
            const uint32_t b0 = pc.class_chunk_begin[cc];
            const uint32_t live = (uint32_t)((ps.n_src[cc] + kIterBlock - 1) / kIterBlock);
            const uint32_t b1 = min(pc.class_chunk_begin[cc + 1], b0 + live);
            uint32_t acc = 0;
            for (uint32_t b = b0 + threadIdx.x; b < b1; b += kIterBlock) acc += A.blk_kept[b];
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t tot = 0;
                for (int w = 0; w < kIterBlock / 32; ++w) tot += s_w[w];
                A.xch_i32[kNumClasses + cc] = (int)tot;
                A.xch_i32[cc] = (int)ps.n_corr[cc];
            }
        }
    } else if (threadIdx.x == 0) {
        for (int cc = 0; cc < kNumClasses; ++cc) {
            ps.n_corr[cc] = (uint32_t)A.xch_i32[cc];
            ps.n_src_g_next[cc] = A.xch_i32[kNumClasses + cc]; // the class size after this iteration's shrinking
        }
    }
}
// after the all-reduce of the per-class sums: every rank solves the same system and advances identically
__global__ void k_shard_solve(DeviceArrays A, int buf, int it_flag) {
    PairState &ps = A.ps[0];
    if (threadIdx.x != 0) return;
    if (ps.status != kRunning) {
        A.h_running_iter[it_flag] = *A.running;
        __threadfence_system();
        return;
    }
    __shared__ double s_scratch[160];
    mulls_icp_trace *tr = trace_of(A, 0);
    for (int cc = 0; cc < kNumClasses; ++cc) {
        ps.n_src_g[cc] = ps.n_src_g_next[cc];
        if (tr && ps.iter < MULLS_MAX_TRACE_ITERS) tr->n_src[ps.iter][cc] = (uint32_t)ps.n_src_g[cc];
    }
    solve_and_advance(A, 0, A.xch_f64, s_scratch, buf ^ 1);
    for (int cc = 0; cc < kNumClasses; ++cc) ps.n_corr[cc] = 0;
    A.h_running_iter[it_flag] = *A.running; // what the launch loop of every rank reads two iterations later
    __threadfence_system();
}
// posterior in sharded mode: VTPV / n_obs of this rank -> exchange buffer
__global__ void k_shard_post(DeviceArrays A, int phase) {
    const PairConst &pc = A.pc[0];
    PairState &ps = A.ps[0];
    if (threadIdx.x != 0) return;
    if (phase == 0) {
        double VTPV = 0.0, nob = 0.0;
        if (ps.status == kNeedPosterior) {
            const int order[6] = {MULLS_GROUND, MULLS_FACADE, MULLS_ROOF, MULLS_PILLAR, MULLS_BEAM, MULLS_VERTEX};
            for (int o = 0; o < 6; ++o) {
                const uint32_t b0 = pc.class_chunk_begin[order[o]];
                const uint32_t live = (uint32_t)((ps.n_src[order[o]] + kIterBlock - 1) / kIterBlock);
                const uint32_t b1 = min(pc.class_chunk_begin[order[o] + 1], b0 + live);
                for (uint32_t b = b0; b < b1; ++b) {
                    VTPV += A.post_partials[2 * (size_t)b];
                    nob += A.post_partials[2 * (size_t)b + 1];
                }
            }
        }
        A.xch_f64[0] = VTPV;
        A.xch_f64[1] = nob;
    } else if (ps.status == kNeedPosterior) {
        const double sigma2 = A.xch_f64[0] / (double)((int)A.xch_f64[1] - 6);
        ps.sigma2 = sigma2;
        ps.code = (sqrt(sigma2) < pc.sigma_thre) ? 1 : -3;
        double inv[36], lu[36];
        inverse6(ps.cofactor, inv, lu);
        for (int k = 0; k < 36; ++k) ps.info[k] = (1.0 / sigma2) * inv[k];
        ps.status = kDone;
    }
}

// ---- k_posterior -------------------------------------------------------------------------------
// cregistration.hpp:2518-2544: VTPV and observation count over the correspondences of the converged
// iteration with its estimate x; sigma^2, code 1 / -3, information matrix (:1386).
__global__ void __launch_bounds__(kIterBlock) k_posterior(DeviceArrays A) {
    if (!chunk_in_run(A)) return;
    const ChunkDesc cd = A.it_chunks[blockIdx.x];
    const PairConst &pc = A.pc[cd.pair];
    PairState &ps = A.ps[cd.pair];
    if (ps.status != kNeedPosterior) return;
    const int c = (int)cd.seg;
    const int buf = ps.final_buf;
    if ((int)cd.first >= ps.n_src[c]) return; // k_finalize only sums the live chunks
    const uint32_t local = cd.first + threadIdx.x;
    const bool valid = (int)local < ps.n_src[c];
    double vtpv = 0.0;
    int nobs = 0;
    if (valid) {
        const uint32_t gi = pc.src_base[c] + local;
        const int j = A.corr_j[gi];
        if (j >= 0) {
            const float4 p = A.src_pos[buf][gi];
            const float4 q = A.tgt_pos[pc.tgt_base[c] + j];
            const float4 qn = A.tgt_nrm[pc.tgt_base[c] + j];
            const float w = A.corr_w[gi];
            const double *x = ps.x;
            const float px = p.x, py = p.y, pz = p.z, qx = q.x, qy = q.y, qz = q.z;
            if (c == MULLS_GROUND || c == MULLS_FACADE || c == MULLS_ROOF) { // :2602-2623
                const float ntx = qn.x, nty = qn.y, ntz = qn.z;
                const float a = ntz * py - nty * pz;
                const float b = ntx * pz - ntz * px;
                const float cc = nty * px - ntx * py;
                const float d = ntx * qx + nty * qy + ntz * qz - ntx * px - nty * py - ntz * pz;
                const float residual = (float)((double)ntx * x[0] + (double)nty * x[1] + (double)ntz * x[2] +
                                               (double)a * x[3] + (double)b * x[4] + (double)cc * x[5] - (double)d);
                vtpv = (double)(w * residual * residual);
                nobs = 1;
            } else {
                const float dx = px - qx, dy = py - qy, dz = pz - qz;
                double Am[3][6], bv[3];
                if (c == MULLS_PILLAR || c == MULLS_BEAM) { // :2643-2673
                    const float vx = qn.x, vy = qn.y, vz = qn.z;
                    Am[0][0] = 0, Am[0][1] = (double)vz, Am[0][2] = (double)(-vy), Am[0][3] = (double)(-vz * pz - vy * py),
                    Am[0][4] = (double)(vy * px), Am[0][5] = (double)(vz * px);
                    Am[1][0] = (double)(-vz), Am[1][1] = 0, Am[1][2] = (double)vx, Am[1][3] = (double)(vx * py),
                    Am[1][4] = (double)(-vx * px - vz * pz), Am[1][5] = (double)(vz * py);
                    Am[2][0] = (double)vy, Am[2][1] = (double)(-vx), Am[2][2] = 0, Am[2][3] = (double)(vx * pz),
                    Am[2][4] = (double)(vy * pz), Am[2][5] = (double)(-vy * py - vx * px);
                    bv[0] = (double)(-vz * dy + vy * dz);
                    bv[1] = (double)(-vx * dz + vz * dx);
                    bv[2] = (double)(-vy * dx + vx * dy);
                } else { // :2559-2583
                    Am[0][0] = 1, Am[0][1] = 0, Am[0][2] = 0, Am[0][3] = 0, Am[0][4] = (double)pz, Am[0][5] = (double)(-py);
                    Am[1][0] = 0, Am[1][1] = 1, Am[1][2] = 0, Am[1][3] = (double)(-pz), Am[1][4] = 0, Am[1][5] = (double)px;
                    Am[2][0] = 0, Am[2][1] = 0, Am[2][2] = 1, Am[2][3] = (double)py, Am[2][4] = (double)(-px), Am[2][5] = 0;
                    bv[0] = (double)(-dx), bv[1] = (double)(-dy), bv[2] = (double)(-dz);
                }
                double r[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double s = Am[k][0] * x[0];
#pragma unroll
                    for (int jj = 1; jj < 6; ++jj) s = s + Am[k][jj] * x[jj];
                    r[k] = s - bv[k];
                }
                vtpv = (double)w * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
                nobs = 3;
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kIterBlock / 32;
    __shared__ double s_v[kWarps];
    __shared__ int s_n[kWarps];
    for (int o = 16; o > 0; o >>= 1) {
        vtpv += __shfl_xor_sync(0xffffffffu, vtpv, o);
        nobs += __shfl_xor_sync(0xffffffffu, nobs, o);
    }
    if (lane == 0) {
        s_v[warp] = vtpv;
        s_n[warp] = nobs;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        int n = 0;
        for (int w = 0; w < kWarps; ++w) {
            v += s_v[w];
            n += s_n[w];
        }
        A.post_partials[2 * (size_t)blockIdx.x] = v;
        A.post_partials[2 * (size_t)blockIdx.x + 1] = (double)n;
    }
}

// ---- k_finalize: one thread per pair: sigma^2 = VTPV/(n-6) (:2536), code 1 / -3 (:2540-2543), information
//      matrix = cofactor^-1 / sigma^2 (:1386). Partials are summed in the class order of :2529-2534.
__global__ void k_finalize(DeviceArrays A, int n_pairs) {
    if (n_pairs < 0) n_pairs = A.ctl->n_pairs; // recorded into the iteration graph: launched over the capacity
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    const PairConst &pc = A.pc[pair];
    PairState &ps = A.ps[pair];
    if (ps.status != kNeedPosterior) return;
    const int order[6] = {MULLS_GROUND, MULLS_FACADE, MULLS_ROOF, MULLS_PILLAR, MULLS_BEAM, MULLS_VERTEX};
    double VTPV = 0.0;
    long long nob = 0;
    for (int o = 0; o < 6; ++o) {
        const uint32_t b0 = pc.class_chunk_begin[order[o]];
        const uint32_t live = (uint32_t)((ps.n_src[order[o]] + kIterBlock - 1) / kIterBlock);
        const uint32_t b1 = min(pc.class_chunk_begin[order[o] + 1], b0 + live);
        for (uint32_t b = b0; b < b1; ++b) {
            VTPV += A.post_partials[2 * (size_t)b];
            nob += (long long)A.post_partials[2 * (size_t)b + 1];
        }
    }
    const double sigma2 = VTPV / (double)((int)nob - 6);
    ps.sigma2 = sigma2;
    ps.code = (sqrt(sigma2) < pc.sigma_thre) ? 1 : -3;
    double inv[36], lu[36];
    inverse6(ps.cofactor, inv, lu);
    for (int k = 0; k < 36; ++k) ps.info[k] = (1.0 / sigma2) * inv[k];
    ps.status = kDone;
}

// ---- k_nn_query: mulls_nn_query — exact 1-NN of arbitrary query points in one target class of pair 0, on the grid
//      the last registration built (what block1->tree_*->nearestKSearch(p, 1) answers in the reference)
__global__ void __launch_bounds__(kIterBlock) k_nn_query(DeviceArrays A, int cls, const float *xyz, uint32_t n, int start_level0,
                                                        int leaf_count, int *out_idx, float *out_d2) {
    const uint32_t i = blockIdx.x * kIterBlock + threadIdx.x;
    if (i >= n) return;
    const PairConst &pc = A.pc[0];
    const PairState &ps = A.ps[0];
    int best_j = -1;
    float best_d2 = INFINITY;
    if (ps.n_tgt[cls] > 0 && !A.hash_used[1]) {
        const GridView g = grid_of(A, pc, ps, cls, leaf_count);
        const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
        const float rmax = 2.5f * pc.thre_unit;
        const float r2 = rmax * rmax * 1.0001f;
        walk_greedy_seed(g, px, py, pz, start_level0, best_d2, best_j);
        WalkStackLocal stk;
        nn_search_walk<WalkStackLocal, kWalkStack>(g, px, py, pz, r2, start_level0, false, best_d2, best_j, stk);
        if (best_j >= 0 && !((double)best_d2 <= (double)rmax * (double)rmax)) best_j = -1;
        if (best_j >= 0) best_j = __float_as_int(__ldg(&g.nrm[best_j]).w);
    }
    out_idx[i] = best_j;
    out_d2[i] = best_j >= 0 ? best_d2 : INFINITY;
}

// ---- k_state_init: reset the per-pair accumulators that the ingest kernels update atomically
__global__ void k_state_init(DeviceArrays A, int n_pairs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    PairState &ps = A.ps[p];
    for (int d = 0; d < 3; ++d) {
        ps.bb_src[d] = ps.bb_tgt[d] = 0x7fffffff;
        ps.bb_src[3 + d] = ps.bb_tgt[3 + d] = (int)0x80000000;
    }
    for (int s = 0; s < kNumSegs; ++s) ps.seg_count[s] = ps.seg_start[s] = 0;
    for (int c = 0; c < kNumClasses; ++c) ps.hash_entries[c] = ps.n_corr[c] = 0;
    ps.status = kRunning;
    if (p == 0) {
        *A.running = n_pairs;
        *A.h_running = n_pairs;
    }
}

// ---- k_collect: pair state -> mulls_icp_result (device copy, then one D2H)
__global__ void k_collect(DeviceArrays A, int n_pairs, mulls_icp_result *out) {
    if (n_pairs < 0) n_pairs = A.ctl->n_pairs;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const PairState &ps = A.ps[p];
    mulls_icp_result &r = out[p];
    for (int k = 0; k < 16; ++k) r.T[k] = ps.T_total[k];
    for (int k = 0; k < 36; ++k) r.info[k] = ps.info[k];
    r.sigma = (float)sqrt(ps.sigma2);
    r.confidence = ps.confidence;
    r.code = ps.code;
    r.iters = ps.iters_entered;
    for (int c = 0; c < kNumClasses; ++c) {
        r.n_corr[c] = ps.n_corr_last[c];
        r.n_src[c] = (uint32_t)ps.n_src_g[c];
    }
}

} // namespace mulls
