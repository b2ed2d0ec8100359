// Exact radius-bounded nearest neighbour on the multi-level hashed grid of one target class — the replacement of
// the kd-tree query of cregistration.hpp:1742-1745 (pcl::registration::CorrespondenceEstimation ->
// KdTreeFLANN::nearestKSearch, k = 1) — written as __host__ __device__ code: k_search (kernels_iterate.cuh) runs it
// on the device, tests/harness/search_host.cu instantiates the very same functions on the CPU, where the CPU suite
// checks them against a brute-force scan (tests/test_search_core.py). The product never runs the host instantiation.
//
// Result: the target j (index inside the Morton-sorted class slice) that minimises the total order
// (FLANN L2_Simple float distance, original index) among all targets with d2 <= r2_prune; the caller applies the
// reference's keep test in double.
//
// Structure of one query (thread):
//   seed     a real candidate: the previous iteration's match, or a short climb/descent through p's own cells
//   level    the smallest level l whose block guarantees coverage of the seed distance (cover_l below)
//   block    the 2x2x2 cells made of p's cell and, per axis, the neighbour on the side of the half-cell p lies in
//            (level 0: by the fractional position inside the cell). The block contains every target closer than
//            cover_l = 0.999 * h_l / 2 (level 0: 0.998 * h0 / 2 — the margins absorb the float rounding of the cell
//            assignment, <= 4096 * 2^-23 cells). The eight cells are tested against the current bound and probed
//            with independent loads; small cells become candidate RANGES, dense cells go to a stack
//   descent  a dense cell is split: its entry carries the mask of existing children, the children that can still
//            beat the bound are probed (again independent loads) and become ranges or stack entries
//   scan     all queued ranges are examined in ONE flat loop (the only loop whose trip count is the number of
//            candidates), which keeps the lanes of a warp together far better than a loop nest per cell
//   stop     best <= cover_l^2 (the best found is the global nearest) or cover_l^2 >= r2_prune; else next level
// The traversal stack and the range queue are 8-byte entries in shared memory (device) — no local-memory frame.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#include "grid_key.cuh"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#else
#error "compile with nvcc (host instantiation: nvcc -x cu, host code only)"
#endif

namespace mulls {

struct HashEntry { // 16 B, one 128-bit load per probe
    uint32_t key_lo, key_hi, start, count;
};

struct GridView {
    const HashEntry *table;
    uint32_t mask;
    const float4 *pos; // class slice
    const float4 *nrm; // class slice (w = original index, for tie-breaks)
    float ox, oy, oz, h0, inv_h0;
    int n_levels;
    int leaf_count; // cells with at most this many points are scanned, larger ones are split
};

struct NoStats {
    MULLS_HD void probe(int) {}
    MULLS_HD void eval(int) {}
    MULLS_HD void expand() {}
    MULLS_HD void level() {}
    MULLS_HD void flush() {}
    MULLS_HD void seed_probe() {}
    MULLS_HD void seed_eval(int) {}
};

MULLS_HD uint4 ld_entry(const HashEntry *e) {
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const uint4 *>(e));
#else
    return *reinterpret_cast<const uint4 *>(e);
#endif
}
MULLS_HD float4 ld_point(const float4 *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
MULLS_HD int f2i_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    union {
        float f;
        int i;
    } u;
    u.f = f;
    return u.i;
#endif
}

// index of the lowest set bit (v != 0)
MULLS_HD int lowest_bit(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// finish a probe whose first slot was loaded by the caller: follow the chain (rare at load factor <= 0.5)
MULLS_HD bool probe_finish(const GridView &g, uint32_t slot, uint4 e, uint32_t klo, uint32_t khi, uint32_t &start,
                           uint32_t &count, uint32_t &cmask) {
    while (true) {
        if (e.x == klo && (e.y & kKeyHiMask) == khi) {
            start = e.z;
            count = e.w;
            cmask = (e.y >> 16) & 0xffu;
            return true;
        }
        if (e.x == 0u && e.y == 0u) return false;
        slot = (slot + 1) & g.mask;
        e = ld_entry(&g.table[slot]);
    }
}

MULLS_HD bool probe_cell(const GridView &g, uint32_t x, uint32_t y, uint32_t z, int level, uint32_t &start, uint32_t &count,
                         uint32_t &cmask) {
    const uint32_t klo = cell_key_lo(x, y, z), khi = cell_key_hi(z, level);
    const uint32_t slot = cell_hash(klo, khi) & g.mask;
    return probe_finish(g, slot, ld_entry(&g.table[slot]), klo, khi, start, count, cmask);
}

// one candidate under the total order (d2, original index)
MULLS_HD void consider(const GridView &g, float d2, uint32_t jj, float &best_d2, int &best_j) {
    if (d2 < best_d2) {
        best_d2 = d2;
        best_j = (int)jj;
    } else if (d2 == best_d2 && best_j >= 0 && (int)jj != best_j) {
        const int oj = f2i_bits(ld_point(&g.nrm[jj]).w);
        const int ob = f2i_bits(ld_point(&g.nrm[best_j]).w);
        if (oj < ob) best_j = (int)jj;
    }
}

// examine every queued range in one flat loop
template <class Scratch, class Stats>
MULLS_HD void scan_ranges(const GridView &g, float px, float py, float pz, Scratch &S, int &nr, float &best_d2, int &best_j,
                          Stats &st) {
    if (nr == 0) return;
    st.flush();
    int ri = 0;
    uint32_t cur = 0, end = 0;
    for (;;) {
        if (cur == end) {
            if (ri == nr) break;
            const uint2 r = S.range(ri++);
            cur = r.x;
            end = r.x + r.y;
            st.eval((int)r.y);
        }
        const float4 q = ld_point(&g.pos[cur]);
        consider(g, flann_l2(px, py, pz, q.x, q.y, q.z), cur, best_d2, best_j);
        ++cur;
    }
    nr = 0;
}

// No (good) candidate yet: climb from p's own level-1 cell to the first level at which it exists and walk down
// through the nearest existing child to a small cell. That cell is queued as a candidate range — its best point seeds
// the exact search (a handful of probes; ties are settled by the search). Returns false if p's cells are all empty.
template <class Stats>
MULLS_HD bool quick_locate(const GridView &g, float px, float py, float pz, int max_level, uint2 &leaf, Stats &st) {
    const int c0x = (int)floorf((px - g.ox) * g.inv_h0);
    const int c0y = (int)floorf((py - g.oy) * g.inv_h0);
    const int c0z = (int)floorf((pz - g.oz) * g.inv_h0);
    const int L = g.n_levels;
    const int top = (max_level < L - 1) ? max_level : L - 1;
    for (int lr = 1; lr <= top; ++lr) {
        const int ncell = 4096 >> lr;
        int cx = c0x >> lr, cy = c0y >> lr, cz = c0z >> lr;
        if (!(cx >= 0 && cy >= 0 && cz >= 0 && cx < ncell && cy < ncell && cz < ncell)) continue;
        uint32_t start, count, cmask;
        st.seed_probe();
        if (!probe_cell(g, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz, lr, start, count, cmask)) continue;
        for (int lv = lr;;) {
            if (count <= (uint32_t)g.leaf_count || lv == 0) {
                st.seed_eval((int)count);
                leaf = make_uint2(start, count);
                return true;
            }
            const float hl = g.h0 * (float)(1 << lv);
            const int ox = (px >= g.ox + ((float)cx + 0.5f) * hl) ? 1 : 0;
            const int oy = (py >= g.oy + ((float)cy + 0.5f) * hl) ? 1 : 0;
            const int oz = (pz >= g.oz + ((float)cz + 0.5f) * hl) ? 1 : 0;
            int ch = ox | (oy << 1) | (oz << 2);
            if (!((cmask >> ch) & 1u)) { // any existing child still yields a valid seed: the one sharing most octant bits
                int bestc = -1, bests = -1;
                for (int k = 0; k < 8; ++k)
                    if ((cmask >> k) & 1u) {
                        const int same = 3 - (((k ^ ch) & 1) + (((k ^ ch) >> 1) & 1) + (((k ^ ch) >> 2) & 1));
                        if (same > bests) bests = same, bestc = k;
                    }
                ch = bestc;
            }
            if (ch < 0) return false;
            cx = 2 * cx + (ch & 1), cy = 2 * cy + ((ch >> 1) & 1), cz = 2 * cz + (ch >> 2);
            --lv;
            st.seed_probe();
            if (!probe_cell(g, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz, lv, start, count, cmask)) return false; // (cannot happen)
        }
    }
    return false;
}

// distance along one axis from p to the slab [lo - margin, hi + margin]
MULLS_HD float slab_dist(float lo, float hi, float p, float margin) {
    return fmaxf(0.0f, fmaxf((lo - margin) - p, p - (hi + margin)));
}

// stack entry of a dense cell: .x = key_lo (x | y<<12 | (z&0xff)<<24), .y = z>>8 | level<<4 | child mask<<8
MULLS_HD uint2 pack_cell(uint32_t x, uint32_t y, uint32_t z, int lv, uint32_t cmask) {
    return make_uint2(cell_key_lo(x, y, z), (z >> 8) | ((uint32_t)lv << 4) | (cmask << 8));
}

// How the lanes of a warp cooperate. The per-thread form (host instantiation, and the reference semantics of the
// device form): nothing is shared. The device form (WarpCoop, kernels_iterate.cuh) keeps the 32 lanes of a warp in
// step through the phases of the search and examines the queued candidates of ALL lanes as one flat list.
struct SoloCoop {
    MULLS_HD bool any(bool b) { return b; }
    template <class Scratch, class Stats>
    MULLS_HD void scan(const GridView &g, float px, float py, float pz, Scratch &S, int &nr, float &best_d2, int &best_j,
                       Stats &st) {
        scan_ranges(g, px, py, pz, S, nr, best_d2, best_j, st);
    }
};

// geometry of one query, fixed for the whole search
struct QueryFrame {
    float fx, fy, fz, flx, fly, flz; // position in level-0 cell units and its floor
    int c0x, c0y, c0z;
    float margin;
};

// the 2x2x2 block of level l around p: cells that can still beat the bound are probed (eight independent loads);
// small cells are queued as candidate ranges, dense ones go to the stack
template <int kStack, class Scratch, class Stats>
MULLS_HD void block_phase(const GridView &g, const QueryFrame &f, float px, float py, float pz, int l, float bound0,
                          Scratch &S, int &nr, int &sp, Stats &st) {
    const float H = g.h0 * (float)(1 << l);
    const int ncell = 4096 >> l;
    const int cx = f.c0x >> l, cy = f.c0y >> l, cz = f.c0z >> l;
    int sx, sy, sz; // side of the half-cell p lies in
    if (l == 0) {
        sx = (f.fx - f.flx) >= 0.5f, sy = (f.fy - f.fly) >= 0.5f, sz = (f.fz - f.flz) >= 0.5f;
    } else {
        sx = (f.c0x >> (l - 1)) & 1, sy = (f.c0y >> (l - 1)) & 1, sz = (f.c0z >> (l - 1)) & 1;
    }
    const int nx = cx + (sx ? 1 : -1), ny = cy + (sy ? 1 : -1), nz = cz + (sz ? 1 : -1);
    // squared distance from p to the neighbour slab along each axis (p is inside its own slab: 0)
    float ex = sx ? ((g.ox + (float)(cx + 1) * H) - f.margin) - px : px - ((g.ox + (float)cx * H) + f.margin);
    float ey = sy ? ((g.oy + (float)(cy + 1) * H) - f.margin) - py : py - ((g.oy + (float)cy * H) + f.margin);
    float ez = sz ? ((g.oz + (float)(cz + 1) * H) - f.margin) - pz : pz - ((g.oz + (float)cz * H) + f.margin);
    ex = fmaxf(ex, 0.0f), ey = fmaxf(ey, 0.0f), ez = fmaxf(ez, 0.0f);
    ex *= ex, ey *= ey, ez *= ez;
    const bool vx0 = cx >= 0 && cx < ncell, vx1 = nx >= 0 && nx < ncell;
    const bool vy0 = cy >= 0 && cy < ncell, vy1 = ny >= 0 && ny < ncell;
    const bool vz0 = cz >= 0 && cz < ncell, vz1 = nz >= 0 && nz < ncell;
    // cells that can still beat the bound, as a mask; then one probe per live cell (k = 0 is p's own cell)
    uint32_t live = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool i = k & 1, j = (k >> 1) & 1, m = k >> 2;
        const float d2c = (i ? ex : 0.0f) + (j ? ey : 0.0f) + (m ? ez : 0.0f);
        if ((i ? vx1 : vx0) && (j ? vy1 : vy0) && (m ? vz1 : vz0) && d2c <= bound0) live |= 1u << k;
    }
    while (live) {
        const int k = lowest_bit(live);
        live &= live - 1;
        const uint32_t x = (uint32_t)((k & 1) ? nx : cx), y = (uint32_t)((k & 2) ? ny : cy), z = (uint32_t)((k & 4) ? nz : cz);
        uint32_t start, count, cmask;
        st.probe(0);
        if (!probe_cell(g, x, y, z, l, start, count, cmask)) continue;
        if (count <= (uint32_t)g.leaf_count || l == 0 || sp == kStack) S.range(nr++) = make_uint2(start, count);
        else S.stack(sp++) = pack_cell(x, y, z, l, cmask);
    }
}

// split the next dense cell of the stack that can still beat the bound: its existing children within the bound are
// probed (independent loads); small ones are queued as ranges, dense ones pushed (nearest octant last = next to pop)
template <int kStack, class Scratch, class Stats>
MULLS_HD void expand_one(const GridView &g, const QueryFrame &f, float px, float py, float pz, float bound, Scratch &S,
                         int &nr, int &sp, Stats &st) {
    while (sp > 0) {
        const uint2 ce = S.stack(--sp);
        const int lv = (int)((ce.y >> 4) & 0xfu);
        const uint32_t cmask = (ce.y >> 8) & 0xffu;
        const int x = (int)(ce.x & 0xfffu), y = (int)((ce.x >> 12) & 0xfffu), z = (int)((ce.x >> 24) | ((ce.y & 0xfu) << 8));
        const float hc = 0.5f * g.h0 * (float)(1 << lv); // child size
        // per-axis squared distances to the two child slabs
        float ax0, ax1, ay0, ay1, az0, az1;
        {
            const float lox = g.ox + (float)(2 * x) * hc, mdx = g.ox + (float)(2 * x + 1) * hc, hix = g.ox + (float)(2 * x + 2) * hc;
            const float loy = g.oy + (float)(2 * y) * hc, mdy = g.oy + (float)(2 * y + 1) * hc, hiy = g.oy + (float)(2 * y + 2) * hc;
            const float loz = g.oz + (float)(2 * z) * hc, mdz = g.oz + (float)(2 * z + 1) * hc, hiz = g.oz + (float)(2 * z + 2) * hc;
            ax0 = slab_dist(lox, mdx, px, f.margin), ax1 = slab_dist(mdx, hix, px, f.margin);
            ay0 = slab_dist(loy, mdy, py, f.margin), ay1 = slab_dist(mdy, hiy, py, f.margin);
            az0 = slab_dist(loz, mdz, pz, f.margin), az1 = slab_dist(mdz, hiz, pz, f.margin);
            ax0 *= ax0, ax1 *= ax1, ay0 *= ay0, ay1 *= ay1, az0 *= az0, az1 *= az1;
        }
        // the cell itself may have fallen behind the bound since it was pushed
        if (fminf(ax0, ax1) + fminf(ay0, ay1) + fminf(az0, az1) > bound) continue;
        st.expand();
        const int near_child = (ax1 < ax0 ? 1 : 0) | (ay1 < ay0 ? 2 : 0) | (az1 < az0 ? 4 : 0);
        // existing children within the bound, re-indexed by c = ch ^ near_child so that the lowest bit is the
        // farthest octant: pushed first, the nearest one last (popped first)
        uint32_t pass = 0;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
            if (((ch & 1) ? ax1 : ax0) + ((ch & 2) ? ay1 : ay0) + ((ch & 4) ? az1 : az0) <= bound) pass |= 1u << ch;
        pass &= cmask;
        if (near_child & 1) pass = ((pass & 0x55u) << 1) | ((pass & 0xaau) >> 1);
        if (near_child & 2) pass = ((pass & 0x33u) << 2) | ((pass & 0xccu) >> 2);
        if (near_child & 4) pass = ((pass & 0x0fu) << 4) | ((pass & 0xf0u) >> 4);
        pass = ((pass & 0x55u) << 1) | ((pass & 0xaau) >> 1); // reverse the 8 bits: farthest (c = 7) becomes bit 0
        pass = ((pass & 0x33u) << 2) | ((pass & 0xccu) >> 2);
        pass = ((pass & 0x0fu) << 4) | ((pass & 0xf0u) >> 4);
        while (pass) {
            const int b = lowest_bit(pass);
            pass &= pass - 1;
            const int ch = (7 - b) ^ near_child;
            const uint32_t x2 = (uint32_t)(2 * x + (ch & 1)), y2 = (uint32_t)(2 * y + ((ch >> 1) & 1)), z2 = (uint32_t)(2 * z + (ch >> 2));
            uint32_t start, count, cm2;
            st.probe(1);
            if (!probe_cell(g, x2, y2, z2, lv - 1, start, count, cm2)) continue;
            if (count <= (uint32_t)g.leaf_count || lv - 1 == 0 || sp == kStack) S.range(nr++) = make_uint2(start, count);
            else S.stack(sp++) = pack_cell(x2, y2, z2, lv - 1, cm2);
        }
        return; // one split per round: what it queued is examined before the next dense cell is opened
    }
}

// `active`: this lane holds a query (all lanes of a warp call the function; see Coop). kRanges >= 8: one block or one
// split queues at most eight ranges between two scans.
// reseed_d2: a candidate farther than this (or none at all) is challenged by the small cell quick_locate finds.
template <int kRanges, int kStack, class Scratch, class Coop, class Stats>
MULLS_HD void nn_search(const GridView &g, bool active, float px, float py, float pz, float r2_prune, int start_level,
                        float reseed_d2, float &best_d2, int &best_j, Scratch &S, Coop &co, Stats &st) {
    static_assert(kRanges >= 8, "a block or a split queues up to eight ranges");
    // best_d2 / best_j come in seeded: (INFINITY, -1) or a real candidate (the previous iteration's match)
    int nr = 0, sp = 0;
    if (active && (best_j < 0 || best_d2 > reseed_d2)) {
        uint2 leaf;
        if (quick_locate(g, px, py, pz, start_level, leaf, st)) S.range(nr++) = leaf;
    }
    co.scan(g, px, py, pz, S, nr, best_d2, best_j, st); // the seed cells of all lanes, examined together
    QueryFrame f;
    f.fx = (px - g.ox) * g.inv_h0, f.fy = (py - g.oy) * g.inv_h0, f.fz = (pz - g.oz) * g.inv_h0;
    f.flx = floorf(f.fx), f.fly = floorf(f.fy), f.flz = floorf(f.fz);
    f.c0x = (int)f.flx, f.c0y = (int)f.fly, f.c0z = (int)f.flz;
    f.margin = 1e-3f * g.h0; // covers the float rounding of the cell assignment
    const int L = g.n_levels;
    int l;
    if (best_j >= 0) { // seeded: the smallest level whose coverage reaches the seed
        const float need = 1.001f * sqrtf(best_d2);
        const float t = need / (0.999f * 0.5f * g.h0);
        if (t <= 1.0f) l = (need <= 0.998f * 0.5f * g.h0) ? 0 : 1;
        else l = ilogbf(t) + 1;
        l = (l < L - 1) ? l : L - 1;
    } else {
        l = (start_level < 1) ? 1 : ((start_level < L - 1) ? start_level : L - 1);
    }
    while (co.any(active)) {
        if (active) {
            st.level();
            block_phase<kStack>(g, f, px, py, pz, l, fminf(best_d2, r2_prune) * 1.0001f + 1e-12f, S, nr, sp, st);
        }
        co.scan(g, px, py, pz, S, nr, best_d2, best_j, st);
        // descent through the dense cells of the block, one split per round
        while (co.any(active && sp > 0)) {
            if (active && sp > 0) expand_one<kStack>(g, f, px, py, pz, fminf(best_d2, r2_prune) * 1.0001f + 1e-12f, S, nr, sp, st);
            co.scan(g, px, py, pz, S, nr, best_d2, best_j, st);
        }
        if (active) {
            const float H = g.h0 * (float)(1 << l);
            const float cover = (l == 0) ? 0.998f * 0.5f * g.h0 : 0.999f * 0.5f * H; // every closer target has been examined
            const float cover2 = cover * cover;
            if (best_d2 <= cover2) active = false;       // the best found is the global nearest
            else if (cover2 >= r2_prune) active = false; // whole search radius examined
            else if (l == L - 1) active = false;         // (n_levels is chosen so that the line above fires first)
            else ++l;
        }
    }
}

// ---- the per-thread depth-first form -------------------------------------------------------------------------
// One query per thread, no cooperation: the block's live cells are walked depth first, nearest octant first, and a
// small cell is examined the moment it is met, so that every later cell is pruned against the tightest bound. Costs
// SIMT efficiency (the 32 walks of a warp diverge) but no synchronisation and the fewest candidates. `defer`: the
// small cells of one block are queued (kRanges) and examined together in one flat loop — pays once the seeds are
// good (late iterations), costs candidates while they are not.
template <int kRanges, int kStack, class Scratch, class Stats>
MULLS_HD void nn_search_dfs(const GridView &g, float px, float py, float pz, float r2_prune, int start_level, bool defer,
                            float &best_d2, int &best_j, Scratch &S, Stats &st) {
    QueryFrame f;
    f.fx = (px - g.ox) * g.inv_h0, f.fy = (py - g.oy) * g.inv_h0, f.fz = (pz - g.oz) * g.inv_h0;
    f.flx = floorf(f.fx), f.fly = floorf(f.fy), f.flz = floorf(f.fz);
    f.c0x = (int)f.flx, f.c0y = (int)f.fly, f.c0z = (int)f.flz;
    f.margin = 1e-3f * g.h0;
    const int L = g.n_levels;
    int l;
    if (best_j >= 0) {
        const float need = 1.001f * sqrtf(best_d2);
        const float t = need / (0.999f * 0.5f * g.h0);
        if (t <= 1.0f) l = (need <= 0.998f * 0.5f * g.h0) ? 0 : 1;
        else l = ilogbf(t) + 1;
        l = (l < L - 1) ? l : L - 1;
    } else {
        l = (start_level < 1) ? 1 : ((start_level < L - 1) ? start_level : L - 1);
    }
    int nr = 0;
    for (;; ++l) {
        st.level();
        const float H = g.h0 * (float)(1 << l);
        const int ncell = 4096 >> l;
        const int cx = f.c0x >> l, cy = f.c0y >> l, cz = f.c0z >> l;
        int sx, sy, sz;
        if (l == 0) {
            sx = (f.fx - f.flx) >= 0.5f, sy = (f.fy - f.fly) >= 0.5f, sz = (f.fz - f.flz) >= 0.5f;
        } else {
            sx = (f.c0x >> (l - 1)) & 1, sy = (f.c0y >> (l - 1)) & 1, sz = (f.c0z >> (l - 1)) & 1;
        }
        const int nx = cx + (sx ? 1 : -1), ny = cy + (sy ? 1 : -1), nz = cz + (sz ? 1 : -1);
        float ex = sx ? ((g.ox + (float)(cx + 1) * H) - f.margin) - px : px - ((g.ox + (float)cx * H) + f.margin);
        float ey = sy ? ((g.oy + (float)(cy + 1) * H) - f.margin) - py : py - ((g.oy + (float)cy * H) + f.margin);
        float ez = sz ? ((g.oz + (float)(cz + 1) * H) - f.margin) - pz : pz - ((g.oz + (float)cz * H) + f.margin);
        ex = fmaxf(ex, 0.0f), ey = fmaxf(ey, 0.0f), ez = fmaxf(ez, 0.0f);
        ex *= ex, ey *= ey, ez *= ez;
        const bool vx0 = cx >= 0 && cx < ncell, vx1 = nx >= 0 && nx < ncell;
        const bool vy0 = cy >= 0 && cy < ncell, vy1 = ny >= 0 && ny < ncell;
        const bool vz0 = cz >= 0 && cz < ncell, vz1 = nz >= 0 && nz < ncell;
        uint32_t live = 0;
        {
            const float bound0 = fminf(best_d2, r2_prune) * 1.0001f + 1e-12f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool i = k & 1, j = (k >> 1) & 1, m = k >> 2;
                const float d2c = (i ? ex : 0.0f) + (j ? ey : 0.0f) + (m ? ez : 0.0f);
                if ((i ? vx1 : vx0) && (j ? vy1 : vy0) && (m ? vz1 : vz0) && d2c <= bound0) live |= 1u << k;
            }
        }
        while (live) { // lowest bit first: k = 0 is p's own cell
            const int k = lowest_bit(live);
            live &= live - 1;
            {
                const float d2c = ((k & 1) ? ex : 0.0f) + ((k & 2) ? ey : 0.0f) + ((k & 4) ? ez : 0.0f);
                if (d2c > fminf(best_d2, r2_prune) * 1.0001f + 1e-12f) continue; // the bound has moved since the mask was built
            }
            int sp = 0;
            S.stack(sp++) = pack_cell((uint32_t)((k & 1) ? nx : cx), (uint32_t)((k & 2) ? ny : cy), (uint32_t)((k & 4) ? nz : cz), l, 0u);
            bool root = true;
            while (sp > 0) {
                const uint2 ce = S.stack(--sp);
                const int lv = (int)((ce.y >> 4) & 0xfu);
                const int x = (int)(ce.x & 0xfffu), y = (int)((ce.x >> 12) & 0xfffu), z = (int)((ce.x >> 24) | ((ce.y & 0xfu) << 8));
                const float hl = g.h0 * (float)(1 << lv);
                if (!root) { // a pushed child may have fallen behind the bound since
                    const float ax = slab_dist(g.ox + (float)x * hl, g.ox + (float)(x + 1) * hl, px, f.margin);
                    const float ay = slab_dist(g.oy + (float)y * hl, g.oy + (float)(y + 1) * hl, py, f.margin);
                    const float az = slab_dist(g.oz + (float)z * hl, g.oz + (float)(z + 1) * hl, pz, f.margin);
                    if (ax * ax + ay * ay + az * az > fminf(best_d2, r2_prune) * 1.0001f + 1e-12f) continue;
                }
                root = false;
                uint32_t start, count, cmask;
                st.probe(lv == l ? 0 : 1);
                if (!probe_cell(g, (uint32_t)x, (uint32_t)y, (uint32_t)z, lv, start, count, cmask)) continue;
                if (count <= (uint32_t)g.leaf_count || lv == 0 || sp + 8 > kStack) {
                    S.range(nr++) = make_uint2(start, count);
                    if (!defer || nr == kRanges) scan_ranges(g, px, py, pz, S, nr, best_d2, best_j, st);
                    continue;
                }
                st.expand();
                const float hc = 0.5f * hl;
                float ax0, ax1, ay0, ay1, az0, az1;
                {
                    const float lox = g.ox + (float)(2 * x) * hc, mdx = g.ox + (float)(2 * x + 1) * hc, hix = g.ox + (float)(2 * x + 2) * hc;
                    const float loy = g.oy + (float)(2 * y) * hc, mdy = g.oy + (float)(2 * y + 1) * hc, hiy = g.oy + (float)(2 * y + 2) * hc;
                    const float loz = g.oz + (float)(2 * z) * hc, mdz = g.oz + (float)(2 * z + 1) * hc, hiz = g.oz + (float)(2 * z + 2) * hc;
                    ax0 = slab_dist(lox, mdx, px, f.margin), ax1 = slab_dist(mdx, hix, px, f.margin);
                    ay0 = slab_dist(loy, mdy, py, f.margin), ay1 = slab_dist(mdy, hiy, py, f.margin);
                    az0 = slab_dist(loz, mdz, pz, f.margin), az1 = slab_dist(mdz, hiz, pz, f.margin);
                    ax0 *= ax0, ax1 *= ax1, ay0 *= ay0, ay1 *= ay1, az0 *= az0, az1 *= az1;
                }
                const float bound = fminf(best_d2, r2_prune) * 1.0001f + 1e-12f;
                const int near_child = (ax1 < ax0 ? 1 : 0) | (ay1 < ay0 ? 2 : 0) | (az1 < az0 ? 4 : 0);
                uint32_t pass = 0;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    if (((ch & 1) ? ax1 : ax0) + ((ch & 2) ? ay1 : ay0) + ((ch & 4) ? az1 : az0) <= bound) pass |= 1u << ch;
                pass &= cmask;
                // re-index by c = ch ^ near_child and reverse: the lowest bit is the farthest octant (pushed first)
                if (near_child & 1) pass = ((pass & 0x55u) << 1) | ((pass & 0xaau) >> 1);
                if (near_child & 2) pass = ((pass & 0x33u) << 2) | ((pass & 0xccu) >> 2);
                if (near_child & 4) pass = ((pass & 0x0fu) << 4) | ((pass & 0xf0u) >> 4);
                pass = ((pass & 0x55u) << 1) | ((pass & 0xaau) >> 1);
                pass = ((pass & 0x33u) << 2) | ((pass & 0xccu) >> 2);
                pass = ((pass & 0x0fu) << 4) | ((pass & 0xf0u) >> 4);
                while (pass) {
                    const int b = lowest_bit(pass);
                    pass &= pass - 1;
                    const int ch = (7 - b) ^ near_child;
                    S.stack(sp++) = pack_cell((uint32_t)(2 * x + (ch & 1)), (uint32_t)(2 * y + ((ch >> 1) & 1)), (uint32_t)(2 * z + (ch >> 2)), lv - 1, 0u);
                }
            }
        }
        scan_ranges(g, px, py, pz, S, nr, best_d2, best_j, st); // what the block deferred
        const float cover = (l == 0) ? 0.998f * 0.5f * g.h0 : 0.999f * 0.5f * H;
        const float cover2 = cover * cover;
        if (best_d2 <= cover2) break;
        if (cover2 >= r2_prune) break;
        if (l == L - 1) break;
    }
}

} // namespace mulls
