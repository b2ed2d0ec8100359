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

// ------------------------------------------------------------------------------------------------
// exact 1-NN within radius on the multi-level hashed grid of one target class
// ------------------------------------------------------------------------------------------------
struct GridView {
    const HashEntry *table;
    uint32_t mask;
    const float4 *pos; // class slice
    const float4 *nrm; // class slice (w = original index, for tie-breaks)
    float ox, oy, oz, h0, inv_h0;
    int n_levels;
    int leaf_count; // cells with at most this many points are scanned, larger ones are descended
    int defer_scan; // queue the leaves of a block and scan them together after its traversal
};

__device__ __forceinline__ bool probe(const GridView &g, uint64_t key, uint32_t &start, uint32_t &count,
                                      uint32_t &child_mask) {
    uint32_t slot = hash_key(key) & g.mask;
    const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
    while (true) {
        const uint4 e = __ldg(reinterpret_cast<const uint4 *>(&g.table[slot]));
        if (e.x == klo && (e.y & kKeyHiMask) == khi) {
            start = e.z;
            count = e.w;
            child_mask = e.y >> 16;
            return true;
        }
        if (e.x == 0u && e.y == 0u) return false;
        slot = (slot + 1) & g.mask;
    }
}

// distance along one axis from p to the (slightly inflated) extent of cell x at a level with cell size H
__device__ __forceinline__ float axis_dist(float o, float H, int x, float p, float margin) {
    const float lo = o + (float)x * H - margin, hi = o + (float)(x + 1) * H + margin;
    return fmaxf(0.0f, fmaxf(lo - p, p - hi));
}

constexpr int kLeafQueue = 8;   // leaves of one block whose scan is deferred to the end of its traversal

// examine the points [start, start+count) of a leaf: FLANN distance, total order (d2, original index)
__device__ __forceinline__ void scan_leaf(const GridView &g, float px, float py, float pz, uint32_t start, uint32_t count,
                                          float &best_d2, int &best_j) {
    for (uint32_t jj = start; jj < start + count; ++jj) {
        const float4 q = __ldg(&g.pos[jj]);
        const float d2 = flann_l2(px, py, pz, q.x, q.y, q.z);
        if (d2 < best_d2) {
            best_d2 = d2;
            best_j = (int)jj;
        } else if (d2 == best_d2 && (int)jj != best_j) {
            const int oj = __float_as_int(__ldg(&g.nrm[jj]).w);
            const int ob = __float_as_int(__ldg(&g.nrm[best_j]).w);
            if (oj < ob) best_j = (int)jj;
        }
    }
}

// No candidate yet (first iteration): walk greedily from p's own cell (first level, from `l` upwards, at which it
// exists) down through the nearest existing child to a leaf and take its best point as the seed. A handful of
// probes, and the exact search that follows has a tight bound from its first cell on instead of stacking every
// sibling within the (large) search radius. Ties are settled by the exact search.
__device__ __forceinline__ void greedy_seed(const GridView &g, float px, float py, float pz, int l, float &best_d2,
                                            int &best_j) {
    const int c0x = (int)floorf((px - g.ox) * g.inv_h0);
    const int c0y = (int)floorf((py - g.oy) * g.inv_h0);
    const int c0z = (int)floorf((pz - g.oz) * g.inv_h0);
    const int L = g.n_levels;
    l = min(max(l, 1), L - 1);
    for (int lr = l; lr < L && best_j < 0; ++lr) {
        const int ncell = (1 << kCoordBits) >> lr;
        int cx = c0x >> lr, cy = c0y >> lr, cz = c0z >> lr;
        if (!(cx >= 0 && cy >= 0 && cz >= 0 && cx < ncell && cy < ncell && cz < ncell)) break;
        uint64_t code = morton36((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
        for (int lv = lr;; --lv) {
            uint32_t start, count, cmask;
            if (!probe(g, cell_key(lv, code), start, count, cmask)) break; // only possible at lv == lr
            if (count <= (uint32_t)g.leaf_count || lv == 0) {
                for (uint32_t jj = start; jj < start + count; ++jj) {
                    const float4 q = __ldg(&g.pos[jj]);
                    const float d2 = flann_l2(px, py, pz, q.x, q.y, q.z);
                    if (d2 < best_d2) {
                        best_d2 = d2;
                        best_j = (int)jj;
                    }
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
            code = (code << 3) | (uint64_t)ch;
            cx = 2 * cx + (ch & 1), cy = 2 * cy + ((ch >> 1) & 1), cz = 2 * cz + (ch >> 2);
        }
    }
}

// squared distance from p to the (slightly inflated) box of cell (x,y,z) at a level with cell size hl
__device__ __forceinline__ float cell_dist2(const GridView &g, float px, float py, float pz, float hl, int x, int y,
                                            int z, float margin) {
    const float ax = axis_dist(g.ox, hl, x, px, margin), ay = axis_dist(g.oy, hl, y, py, margin),
                az = axis_dist(g.oz, hl, z, pz, margin);
    return ax * ax + ay * ay + az * az;
}

constexpr int kPacketStack = 96; // warp-shared DFS entries of the packet search (3 words each, in shared memory)

// Packet search: the 32 queries of a warp (Morton-adjacent source points, each with a seed) are resolved by ONE
// warp-uniform traversal. The warp walks the cells overlapping the union of the lanes' search balls; a cell is
// visited if ANY lane can still improve inside it, its points are then tested by the lanes that need it. No lane
// waits for another lane's traversal (the cause of the 11/32 thread efficiency of the per-thread search), at the
// price of testing the union of the candidate sets. Exact for every lane: each lane sees a superset of the cells
// its own ball overlaps. Returns false (nothing done) when the packet is too spread out for this to pay.
__device__ __forceinline__ bool packet_search(const GridView &g, bool valid, float px, float py, float pz, float r2_prune,
                                              float max_ext, float &best_d2, int &best_j, uint32_t *stk) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const float margin = 1e-3f * g.h0;
    const float bound0 = fminf(best_d2, r2_prune);
    const float r = valid ? sqrtf(bound0) * 1.0001f + 2.0f * margin : 0.0f;
    float lo[3] = {valid ? px - r : INFINITY, valid ? py - r : INFINITY, valid ? pz - r : INFINITY};
    float hi[3] = {valid ? px + r : -INFINITY, valid ? py + r : -INFINITY, valid ? pz + r : -INFINITY};
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(full, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(full, hi[d], o));
        }
    if (!(hi[0] >= lo[0])) return true; // no valid lane
    const float ext = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
    if (!(ext <= max_ext)) return false; // spread-out packet (or an unseeded lane): per-thread search instead
    // level whose cells are at least as wide as the union box: it overlaps at most 2 (3 with rounding) cells per axis
    const int L = g.n_levels;
    int l = 0;
    while (l < L - 1 && g.h0 * (float)(1 << l) < ext) ++l;
    const int ncell = (1 << kCoordBits) >> l;
    int clo[3], chi[3];
    {
        const float o3[3] = {g.ox, g.oy, g.oz};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            clo[d] = max(((int)floorf((lo[d] - o3[d]) * g.inv_h0)) >> l, 0);
            chi[d] = min(((int)floorf((hi[d] - o3[d]) * g.inv_h0)) >> l, ncell - 1);
        }
    }
    if ((chi[0] - clo[0] + 1) * (chi[1] - clo[1] + 1) * (chi[2] - clo[2] + 1) > 27) return false;
    // warp-uniform DFS; the stack lives in shared memory, written by lane 0
    int sp = 0;
    for (int z = clo[2]; z <= chi[2]; ++z)
        for (int y = clo[1]; y <= chi[1]; ++y)
            for (int x = clo[0]; x <= chi[0]; ++x) {
                if (lane == 0) {
                    const uint64_t code = morton36((uint32_t)x, (uint32_t)y, (uint32_t)z);
                    stk[3 * sp + 0] = (uint32_t)code;
                    stk[3 * sp + 1] = (uint32_t)(code >> 32) | ((uint32_t)l << 4) | ((uint32_t)x << 8) | ((uint32_t)y << 20);
                    stk[3 * sp + 2] = (uint32_t)z;
                }
                ++sp;
            }
    __syncwarp();
    while (sp > 0) {
        --sp;
        const uint32_t w0 = stk[3 * sp + 0], meta = stk[3 * sp + 1], w2 = stk[3 * sp + 2];
        __syncwarp(); // everyone has read the entry before lane 0 may overwrite the slot
        const uint64_t code = (uint64_t)w0 | ((uint64_t)(meta & 0xf) << 32);
        const int lv = (int)((meta >> 4) & 0xf), cx = (int)((meta >> 8) & 0xfff), cy = (int)(meta >> 20), cz = (int)w2;
        const float hl = g.h0 * (float)(1 << lv);
        const bool need = valid && cell_dist2(g, px, py, pz, hl, cx, cy, cz, margin) <= fminf(best_d2, r2_prune) * 1.0001f + 1e-12f;
        if (!__any_sync(full, need)) continue;
        uint32_t start, count, cmask;
        if (!probe(g, cell_key(lv, code), start, count, cmask)) continue; // uniform key: one broadcast load
        if (count <= (uint32_t)g.leaf_count || lv == 0 || sp + 8 > kPacketStack) {
            if (need) scan_leaf(g, px, py, pz, start, count, best_d2, best_j);
        } else {
            if (lane == 0) {
                int k = sp;
                for (int ch = 0; ch < 8; ++ch)
                    if ((cmask >> ch) & 1u) {
                        const uint64_t cc = (code << 3) | (uint64_t)ch;
                        stk[3 * k + 0] = (uint32_t)cc;
                        stk[3 * k + 1] = (uint32_t)(cc >> 32) | ((uint32_t)(lv - 1) << 4) | ((uint32_t)(2 * cx + (ch & 1)) << 8) |
                                         ((uint32_t)(2 * cy + ((ch >> 1) & 1)) << 20);
                        stk[3 * k + 2] = (uint32_t)(2 * cz + (ch >> 2));
                        ++k;
                    }
            }
            sp += __popc(cmask & 0xffu);
            __syncwarp();
        }
    }
    return true;
}

constexpr int kStackDepth = 48; // DFS entries: at most 7 stay behind per descended level

// Exact nearest target (index within the class slice) under the total order (d2, original index) among
// all targets with d2 <= r2_prune.
//
// Ascend: at level l >= 1 the 2x2x2 block of cells that contains p's own cell and, along every axis, the
// neighbour on the side of the half of the cell p lies in, covers the 3x3x3 block of level l-1 around p.
// Every target closer than cover_l = 0.999 * h_(l-1) is therefore inside the block (the 0.1% absorbs the
// float rounding of the cell assignment), and the search stops at the first level whose block has been
// examined with best <= cover_l or cover_l >= radius. 8 probes per level instead of 27.
// Descend: a cell holding more than leaf_count points is not scanned but split: its entry carries the
// mask of existing children, the per-axis distances to the two child slabs are computed once, and only
// children that exist and can still beat the best distance are pushed (nearest octant last = popped first;
// Morton code = parent code << 3 | child) — the octree analogue of the kd-tree descent it replaces.
// `budget` > 0 bounds the number of cell visits: when it runs out the function returns false with the best
// candidate found so far (a valid seed for a second, unbounded call) — used to keep the 32 traversals of a
// warp from waiting on a few expensive queries (they are regrouped and finished together, see k_search).
__device__ __forceinline__ bool nn_search(const GridView &g, float px, float py, float pz, float r2_prune,
                                          int start_level, float &best_d2, int &best_j, int budget) {
    // best_d2 / best_j come in seeded: (INFINITY, -1) or a real candidate (the previous iteration's match)
    int steps_left = (budget > 0) ? budget : 0x7fffffff;
    const int c0x = (int)floorf((px - g.ox) * g.inv_h0);
    const int c0y = (int)floorf((py - g.oy) * g.inv_h0);
    const int c0z = (int)floorf((pz - g.oz) * g.inv_h0);
    const int L = g.n_levels;
    const float margin = 1e-3f * g.h0; // covers the float rounding of the cell assignment
    uint32_t st_code[kStackDepth], st_meta[kStackDepth], st_z[kStackDepth];
    float st_d2[kStackDepth];
    uint32_t q_start[kLeafQueue], q_count[kLeafQueue];
    int nq = 0;
    int l = min(max(start_level, 1), L - 1);
    if (best_j >= 0) { // seeded: the smallest level whose coverage 0.999 * h0 * 2^(l-1) reaches the seed
        const float need = 1.001f * sqrtf(best_d2) / (0.999f * 0.5f * g.h0);
        l = min(max((need <= 1.0f) ? 0 : (ilogbf(need) + 1), 1), L - 1);
    }
    for (;; ++l) {
        const float H = g.h0 * (float)(1 << l);
        const int ncell = (1 << kCoordBits) >> l;
        int xs[2], ys[2], zs[2];
        xs[0] = c0x >> l, ys[0] = c0y >> l, zs[0] = c0z >> l;
        xs[1] = xs[0] + (((c0x >> (l - 1)) & 1) ? 1 : -1);
        ys[1] = ys[0] + (((c0y >> (l - 1)) & 1) ? 1 : -1);
        zs[1] = zs[0] + (((c0z >> (l - 1)) & 1) ? 1 : -1);
        float ex[2], ey[2], ez[2];
        uint64_t sx[2], sy[2], sz[2];
        bool vx[2], vy[2], vz[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            vx[i] = xs[i] >= 0 && xs[i] < ncell;
            vy[i] = ys[i] >= 0 && ys[i] < ncell;
            vz[i] = zs[i] >= 0 && zs[i] < ncell;
            ex[i] = axis_dist(g.ox, H, xs[i], px, margin);
            ey[i] = axis_dist(g.oy, H, ys[i], py, margin);
            ez[i] = axis_dist(g.oz, H, zs[i], pz, margin);
            ex[i] *= ex[i], ey[i] *= ey[i], ez[i] *= ez[i];
            sx[i] = spread12((uint32_t)xs[i]);
            sy[i] = spread12((uint32_t)ys[i]) << 1;
            sz[i] = spread12((uint32_t)zs[i]) << 2;
        }
        // Live cells of the block as a bit mask, then one loop trip per LIVE cell: the 32 lanes of a warp run
        // their i-th live cell together instead of idling through each other's pruned slots.
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
            {
                const uint64_t code = sx[i] | sy[j] | sz[m];
                st_code[0] = (uint32_t)code;
                st_meta[0] = (uint32_t)(code >> 32) | ((uint32_t)l << 4) | ((uint32_t)xs[i] << 8) | ((uint32_t)ys[j] << 20);
                st_z[0] = (uint32_t)zs[m];
                st_d2[0] = ex[i] + ey[j] + ez[m];
                sp = 1;
            }
            while (sp > 0) {
                --sp;
                // a cell farther than the best so far (or than the radius) cannot change the result
                if (st_d2[sp] > fminf(best_d2, r2_prune) * 1.0001f + 1e-12f) continue;
                if (--steps_left < 0) return false;
                const uint32_t meta = st_meta[sp];
                const uint64_t code = (uint64_t)st_code[sp] | ((uint64_t)(meta & 0xf) << 32);
                const int lv = (int)((meta >> 4) & 0xf);
                uint32_t start, count, cmask;
                if (!probe(g, cell_key(lv, code), start, count, cmask)) continue;
                if (count <= (uint32_t)g.leaf_count || lv == 0 || sp + 8 > kStackDepth) {
                    if (g.defer_scan && nq < kLeafQueue) { // scanned together with the block's other leaves
                        q_start[nq] = start;
                        q_count[nq] = count;
                        ++nq;
                    } else {
                        scan_leaf(g, px, py, pz, start, count, best_d2, best_j);
                    }
                } else {
                    const int cx = (int)((meta >> 8) & 0xfff), cy = (int)(meta >> 20), cz = (int)st_z[sp];
                    const float hc = 0.5f * g.h0 * (float)(1 << lv);
                    float ax[2], ay[2], az[2];
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        ax[b] = axis_dist(g.ox, hc, 2 * cx + b, px, margin);
                        ay[b] = axis_dist(g.oy, hc, 2 * cy + b, py, margin);
                        az[b] = axis_dist(g.oz, hc, 2 * cz + b, pz, margin);
                        ax[b] *= ax[b], ay[b] *= ay[b], az[b] *= az[b];
                    }
                    // octant of p relative to the cell centre: the child with zero (or least) distance
                    const int near_child = (ax[1] < ax[0] ? 1 : 0) | (ay[1] < ay[0] ? 2 : 0) | (az[1] < az[0] ? 4 : 0);
                    const float bound = fminf(best_d2, r2_prune) * 1.0001f + 1e-12f;
                    // children that exist and can still beat the bound, as a bit mask (no 8-trip loop) ...
                    uint32_t pass = 0;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch)
                        if (ax[ch & 1] + ay[(ch >> 1) & 1] + az[ch >> 2] <= bound) pass |= 1u << ch;
                    pass &= cmask;
                    // ... re-indexed by c = ch ^ near_child (bit permutation by conditional swaps), so that the
                    // highest set bit is the farthest octant: pushed first, the nearest one last (popped first)
                    if (near_child & 1) pass = ((pass & 0x55u) << 1) | ((pass & 0xaau) >> 1);
                    if (near_child & 2) pass = ((pass & 0x33u) << 2) | ((pass & 0xccu) >> 2);
                    if (near_child & 4) pass = ((pass & 0x0fu) << 4) | ((pass & 0xf0u) >> 4);
                    while (pass) {
                        const int c = 31 - __clz((int)pass);
                        pass ^= 1u << c;
                        const int ch = c ^ near_child;
                        const uint64_t cc = (code << 3) | (uint64_t)ch;
                        st_code[sp] = (uint32_t)cc;
                        st_meta[sp] = (uint32_t)(cc >> 32) | ((uint32_t)(lv - 1) << 4) |
                                      ((uint32_t)(2 * cx + (ch & 1)) << 8) | ((uint32_t)(2 * cy + ((ch >> 1) & 1)) << 20);
                        st_z[sp] = (uint32_t)(2 * cz + (ch >> 2));
                        st_d2[sp] = ax[ch & 1] + ay[(ch >> 1) & 1] + az[ch >> 2];
                        ++sp;
                    }
                }
            }
        }
        // the queued leaves of this block: the lanes of a warp scan them at the same time
        for (int qi = 0; qi < nq; ++qi) scan_leaf(g, px, py, pz, q_start[qi], q_count[qi], best_d2, best_j);
        nq = 0;
        const float cover = 0.999f * 0.5f * H; // every target closer than this has been examined
        const float cover2 = cover * cover;
        if (best_d2 <= cover2) break;  // the best found is the global nearest
        if (cover2 >= r2_prune) break; // whole search radius examined
        if (l == L - 1) break;         // (n_levels is chosen so that the line above fires first)
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// exact k nearest targets (k = 10) for the normal-shooting correspondences of :1732-1737
// (pcl::registration::CorrespondenceEstimationNormalShooting): same hierarchy as nn_search, the pruning bound is
// the current k-th best distance and there is no search radius — the level pyramid of a pair that uses normal
// shooting goes up to a block that spans the whole grid, so the result is exact however far the targets are.
// Total order (d2, original index).
// ------------------------------------------------------------------------------------------------
constexpr int kShootK = 10;

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
    uint32_t st_code[kStackDepth], st_meta[kStackDepth], st_z[kStackDepth];
    float st_d2[kStackDepth];
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
            {
                const uint64_t code = morton36((uint32_t)x, (uint32_t)y, (uint32_t)z);
                st_code[0] = (uint32_t)code;
                st_meta[0] = (uint32_t)(code >> 32) | ((uint32_t)l << 4) | ((uint32_t)x << 8) | ((uint32_t)y << 20);
                st_z[0] = (uint32_t)z;
                st_d2[0] = cell_dist2(g, px, py, pz, H, x, y, z, margin);
                sp = 1;
            }
            while (sp > 0) {
                --sp;
                if (st_d2[sp] > knn_bound(kl) * 1.0001f + 1e-12f) continue;
                const uint32_t meta = st_meta[sp];
                const uint64_t code = (uint64_t)st_code[sp] | ((uint64_t)(meta & 0xf) << 32);
                const int lv = (int)((meta >> 4) & 0xf);
                uint32_t start, count, cmask;
                if (!probe(g, cell_key(lv, code), start, count, cmask)) continue;
                if (count <= (uint32_t)g.leaf_count || lv == 0 || sp + 8 > kStackDepth) {
                    for (uint32_t jj = start; jj < start + count; ++jj) {
                        const float4 q = __ldg(&g.pos[jj]);
                        knn_insert(g, kl, flann_l2(px, py, pz, q.x, q.y, q.z), (int)jj);
                    }
                } else {
                    const int cx = (int)((meta >> 8) & 0xfff), cy = (int)(meta >> 20), cz = (int)st_z[sp];
                    const float hc = 0.5f * g.h0 * (float)(1 << lv);
                    for (int ch = 7; ch >= 0; --ch) {
                        if (!((cmask >> ch) & 1u)) continue;
                        const int x2 = 2 * cx + (ch & 1), y2 = 2 * cy + ((ch >> 1) & 1), z2 = 2 * cz + (ch >> 2);
                        const float d2c = cell_dist2(g, px, py, pz, hc, x2, y2, z2, margin);
                        if (d2c > knn_bound(kl) * 1.0001f + 1e-12f) continue;
                        const uint64_t cc = (code << 3) | (uint64_t)ch;
                        st_code[sp] = (uint32_t)cc;
                        st_meta[sp] = (uint32_t)(cc >> 32) | ((uint32_t)(lv - 1) << 4) | ((uint32_t)x2 << 8) | ((uint32_t)y2 << 20);
                        st_z[sp] = (uint32_t)z2;
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
// start level for a search seeded with a candidate at squared distance d2: the smallest level whose
// guaranteed coverage 0.999 * h0 * 2^(l-1) reaches that distance
__device__ __forceinline__ int level_for_distance(const GridView &g, float d2) {
    const float need = 1.001f * sqrtf(d2) / (0.999f * 0.5f * g.h0);
    return (need <= 1.0f) ? 0 : (ilogbf(need) + 1);
}

__global__ void __launch_bounds__(kIterBlock) k_search(DeviceArrays A, int buf, int start_level0, int leaf_count,
                                                      int budget, int defer_scan, float packet_max_ext) {
    const ChunkDesc cd = A.it_chunks[blockIdx.x];
    const PairConst &pc = A.pc[cd.pair];
    const PairState &ps = A.ps[cd.pair];
    if (ps.status != kRunning || A.hash_used[1]) return;
    const int c = (int)cd.seg;
    const int ns = ps.n_src[c], nt = ps.n_tgt[c], nsg = ps.n_src_g[c];
    if ((int)cd.first >= ns) return; // block-uniform
    const uint32_t local = cd.first + threadIdx.x;
    const bool valid = (int)local < ns;
    const uint32_t gi = pc.src_base[c] + (valid ? local : 0);
    float4 p = A.src_pos[buf][gi];
    float4 n = A.src_nrm[buf][gi];
    if (valid && ps.iter > 0) {
        // cregistration.hpp:1260 — incremental in-place update of the float source cloud
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
    // determine_corres needs >= 3 points on both sides (:1727-1728)
    if (!(pc.used[c] && nsg >= 3 && nt >= 3)) { // block-uniform
        if (valid) {
            A.nn_idx[gi] = -1;
            A.nn_d2[gi] = INFINITY;
        }
        return;
    }
    GridView g;
    g.table = A.hash + ps.hash_base[c];
    g.mask = ps.hash_mask[c];
    g.pos = A.tgt_pos + pc.tgt_base[c];
    g.nrm = A.tgt_nrm + pc.tgt_base[c];
    g.ox = ps.origin[0], g.oy = ps.origin[1], g.oz = ps.origin[2];
    g.h0 = ps.h0, g.inv_h0 = ps.inv_h0;
    g.n_levels = ps.n_levels;
    g.leaf_count = leaf_count;
    // deferring pays once the seeds are good (from the third iteration on: the big first corrections are applied)
    g.defer_scan = (defer_scan == 2) ? (ps.iter >= 2) : defer_scan;
    // CorrespondenceEstimation keeps d2 <= (2.5*thre)^2, evaluated in double (:1745, PCL)
    const float max_distance_f = 2.5f * ps.thre;
    const double max_dist_sqr = (double)max_distance_f * (double)max_distance_f;
    const float r2_prune = (float)max_dist_sqr * 1.0001f;

    if (pc.normal_shooting && (c == MULLS_GROUND || c == MULLS_FACADE || c == MULLS_ROOF)) { // block-uniform
        // :1732-1737 normal shooting [PCL CorrespondenceEstimationNormalShooting, k = 10]: among the 10 nearest targets
        // the one with the smallest squared distance to the line through the source point along its normal; dropped
        // if that value exceeds max_distance (NOT squared); correspondence distance = its squared NN distance.
        int sj = -1;
        float sd2 = INFINITY;
        if (valid) {
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
        return;
    }
    // seeds: the previous iteration's match (a real candidate, so the box-distance pruning bites from the first
    // cell on and the search only has to prove that nothing is closer), else a greedy descent
    int best_j = -1;
    float best_d2 = INFINITY;
    if (valid) {
        const int pj = A.src_prevj[buf][gi];
        if (pj >= 0) {
            const float4 q = __ldg(&g.pos[pj]);
            best_d2 = flann_l2(p.x, p.y, p.z, q.x, q.y, q.z);
            best_j = pj;
        } else {
            greedy_seed(g, p.x, p.y, p.z, start_level0, best_d2, best_j);
        }
    }
    // packet search: one warp-uniform traversal for the 32 queries of the warp, when they are close together
    __shared__ uint32_t s_pstack[kIterBlock / 32][3 * kPacketStack];
    bool finished = false;
    if (packet_max_ext > 0.0f)
        finished = packet_search(g, valid, p.x, p.y, p.z, r2_prune, packet_max_ext, best_d2, best_j, s_pstack[threadIdx.x >> 5]);
    // per-thread search (pass 1, optionally with a bounded number of cell visits)
    if (!finished) {
        finished = true;
        if (valid) finished = nn_search(g, p.x, p.y, p.z, r2_prune, start_level0, best_d2, best_j, budget);
    }
    // pass 2: the unfinished (expensive) queries of the block are regrouped into the first threads and
    // finished there, seeded with what pass 1 found — warps of similar cost instead of 31 idle lanes
    __shared__ float4 s_q[kIterBlock];   // x y z best_d2
    __shared__ int s_qj[kIterBlock];     // best_j
    __shared__ int s_qi[kIterBlock];     // thread that owns the query
    __shared__ int s_cnt;
    if (budget > 0) {
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (!finished) {
            const int slot = atomicAdd(&s_cnt, 1);
            s_q[slot] = make_float4(p.x, p.y, p.z, best_d2);
            s_qj[slot] = best_j;
            s_qi[slot] = (int)threadIdx.x;
        }
        __syncthreads();
        const int m = s_cnt;
        if ((int)threadIdx.x < m) {
            const float4 q = s_q[threadIdx.x];
            float d2 = q.w;
            int j = s_qj[threadIdx.x];
            const int sl = (j >= 0) ? level_for_distance(g, d2) : start_level0;
            nn_search(g, q.x, q.y, q.z, r2_prune, sl, d2, j, 0);
            s_q[threadIdx.x].w = d2;
            s_qj[threadIdx.x] = j;
        }
        __syncthreads();
        if (!finished) { // fetch the result back (slot order is arbitrary: find my slot)
            for (int t = 0; t < m; ++t)
                if (s_qi[t] == (int)threadIdx.x) {
                    best_d2 = s_q[t].w;
                    best_j = s_qj[t];
                }
        }
    }
    if (!valid) return;
    if (best_j >= 0 && !((double)best_d2 <= max_dist_sqr)) best_j = -1;
    if (best_j >= 0) {
        // duplicate_check_table as a claim: the lowest source index wins (:1762-1786, Q5)
        atomicMin(&A.claim[pc.tgt_base[c] + best_j], (unsigned)__float_as_int(n.w));
    }
    A.nn_idx[gi] = best_j;
    A.nn_d2[gi] = best_d2;
}

// ---- k_resolve ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kIterBlock) k_resolve(DeviceArrays A, int buf) {
    const ChunkDesc cd = A.it_chunks[blockIdx.x];
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
        A.blk_kept[blockIdx.x] = k;
        if (p) atomicAdd(&ps.n_corr[c], p);
    }
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
__device__ void solve_and_advance(DeviceArrays &A, uint32_t pair, const double *S, double *sm /*>= 150 doubles*/,
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
    mulls_icp_trace *tr = A.trace ? &A.trace[pair] : nullptr;
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
__global__ void __launch_bounds__(kIterBlock) k_accumulate(DeviceArrays A, int buf) {
    const ChunkDesc cd = A.it_chunks[blockIdx.x];
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
        for (uint32_t b = first_chunk + threadIdx.x; b < blockIdx.x; b += kIterBlock) acc += A.blk_kept[b];
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
        fl = A.flags[gi];
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
    float ratio_unused;
    const bool few = too_few(pc, ps, ps.n_corr, ratio_unused);
    if (pass && !few) {
        const float4 q = A.tgt_pos[pc.tgt_base[c] + j];
        const float4 qn = A.tgt_nrm[pc.tgt_base[c] + j];
        const int it = ps.iter;
        const bool resid_w = pc.w_residual && it > 2; // :1905-1907
        const bool dist_w = pc.w_dist != 0, inten_w = pc.w_intensity != 0;
        if (c == MULLS_GROUND || c == MULLS_FACADE || c == MULLS_ROOF) {
            const float wc = (c == MULLS_FACADE) ? 1.0f : balanced_ground_weight(pc, ps.n_corr);
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
        A.partials[(size_t)blockIdx.x * kTerms + threadIdx.x] = v;
    }
}

// ---- k_solve: one block per pair, after k_accumulate. Sums the per-chunk partials of every class in chunk
//      order (fixed order => bit-reproducible), then one thread solves and advances the pair state.
constexpr int kSolveThreads = kNumClasses * 32; // one warp per feature class
__global__ void __launch_bounds__(kSolveThreads) k_solve(DeviceArrays A, int buf) {
    const uint32_t pair = blockIdx.x;
    const PairConst &pc = A.pc[pair];
    PairState &ps = A.ps[pair];
    if (ps.status != kRunning) return;
    const int lane = threadIdx.x & 31, cc = threadIdx.x >> 5; // warp = class
    __shared__ double s_S[kNumClasses][kTerms];
    __shared__ double s_scratch[160];
    __shared__ int s_newn[kNumClasses];
    {
        const uint32_t b0 = pc.class_chunk_begin[cc];
        // only the chunks that held live sources this iteration wrote a partial
        const uint32_t live = (uint32_t)((ps.n_src[cc] + kIterBlock - 1) / kIterBlock);
        const uint32_t b1 = min(pc.class_chunk_begin[cc + 1], b0 + live);
        // lane = term; chunks in order, four independent accumulators combined in a fixed order
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (lane < 27) {
            uint32_t b = b0;
            for (; b + 4 <= b1; b += 4) {
                a0 += A.partials[(size_t)(b + 0) * kTerms + lane];
                a1 += A.partials[(size_t)(b + 1) * kTerms + lane];
                a2 += A.partials[(size_t)(b + 2) * kTerms + lane];
                a3 += A.partials[(size_t)(b + 3) * kTerms + lane];
            }
            for (; b < b1; ++b) a0 += A.partials[(size_t)b * kTerms + lane];
        }
        if (lane < kTerms) s_S[cc][lane] = (lane < 27) ? ((a0 + a1) + (a2 + a3)) : 0.0;
        uint32_t acc = 0; // kept sources of the class = its new size
        for (uint32_t b = b0 + lane; b < b1; b += 32) acc += A.blk_kept[b];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_newn[cc] = (int)acc;
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
        mulls_icp_trace *tr = A.trace ? &A.trace[pair] : nullptr;
        for (int cc = 0; cc < kNumClasses; ++cc) {
            ps.n_src[cc] = s_newn[cc]; // classes that skipped determine_corres keep everything (k_resolve)
            ps.n_src_g[cc] = s_newn[cc];
            if (tr && ps.iter < MULLS_MAX_TRACE_ITERS) tr->n_src[ps.iter][cc] = (uint32_t)ps.n_src[cc];
        }
        solve_and_advance(A, pair, &s_S[0][0], s_scratch, buf ^ 1);
        for (int cc = 0; cc < kNumClasses; ++cc) ps.n_corr[cc] = 0;
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
__global__ void k_shard_solve(DeviceArrays A, int buf) {
    PairState &ps = A.ps[0];
    if (ps.status != kRunning || threadIdx.x != 0) return;
    __shared__ double s_scratch[160];
    mulls_icp_trace *tr = A.trace ? &A.trace[0] : nullptr;
    for (int cc = 0; cc < kNumClasses; ++cc) {
        ps.n_src_g[cc] = ps.n_src_g_next[cc];
        if (tr && ps.iter < MULLS_MAX_TRACE_ITERS) tr->n_src[ps.iter][cc] = (uint32_t)ps.n_src_g[cc];
    }
    solve_and_advance(A, 0, A.xch_f64, s_scratch, buf ^ 1);
    for (int cc = 0; cc < kNumClasses; ++cc) ps.n_corr[cc] = 0;
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
