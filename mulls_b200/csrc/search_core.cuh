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

// No candidate yet: climb from p's own level-1 cell to the first level at which it exists, walk down through the
// nearest existing child to a small cell and take its best point. A handful of probes; ties are settled by the
// exact search that follows.
template <class Stats>
MULLS_HD void quick_seed(const GridView &g, float px, float py, float pz, int max_level, float &best_d2, int &best_j,
                         Stats &st) {
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
                for (uint32_t jj = start; jj < start + count; ++jj) {
                    const float4 q = ld_point(&g.pos[jj]);
                    const float d2 = flann_l2(px, py, pz, q.x, q.y, q.z);
                    if (d2 < best_d2) {
                        best_d2 = d2;
                        best_j = (int)jj;
                    }
                }
                return;
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
            if (ch < 0) return;
            cx = 2 * cx + (ch & 1), cy = 2 * cy + ((ch >> 1) & 1), cz = 2 * cz + (ch >> 2);
            --lv;
            st.seed_probe();
            if (!probe_cell(g, (uint32_t)cx, (uint32_t)cy, (uint32_t)cz, lv, start, count, cmask)) return; // (cannot happen)
        }
    }
}

// distance along one axis from p to the slab [lo - margin, hi + margin]
MULLS_HD float slab_dist(float lo, float hi, float p, float margin) {
    return fmaxf(0.0f, fmaxf((lo - margin) - p, p - (hi + margin)));
}

// stack entry of a dense cell: .x = key_lo (x | y<<12 | (z&0xff)<<24), .y = z>>8 | level<<4 | child mask<<8
MULLS_HD uint2 pack_cell(uint32_t x, uint32_t y, uint32_t z, int lv, uint32_t cmask) {
    return make_uint2(cell_key_lo(x, y, z), (z >> 8) | ((uint32_t)lv << 4) | (cmask << 8));
}

template <int kRanges, int kStack, class Scratch, class Stats>
MULLS_HD void nn_search(const GridView &g, float px, float py, float pz, float r2_prune, int start_level, float &best_d2,
                        int &best_j, Scratch &S, Stats &st) {
    // best_d2 / best_j come in seeded: (INFINITY, -1) or a real candidate
    const float fx = (px - g.ox) * g.inv_h0, fy = (py - g.oy) * g.inv_h0, fz = (pz - g.oz) * g.inv_h0;
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    const int c0x = (int)flx, c0y = (int)fly, c0z = (int)flz;
    const int L = g.n_levels;
    const float margin = 1e-3f * g.h0; // covers the float rounding of the cell assignment
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
    int nr = 0, sp = 0;
    for (;; ++l) {
        st.level();
        const float H = g.h0 * (float)(1 << l);
        const int ncell = 4096 >> l;
        const int cx = c0x >> l, cy = c0y >> l, cz = c0z >> l;
        int sx, sy, sz; // side of the half-cell p lies in
        if (l == 0) {
            sx = (fx - flx) >= 0.5f, sy = (fy - fly) >= 0.5f, sz = (fz - flz) >= 0.5f;
        } else {
            sx = (c0x >> (l - 1)) & 1, sy = (c0y >> (l - 1)) & 1, sz = (c0z >> (l - 1)) & 1;
        }
        const int nx = cx + (sx ? 1 : -1), ny = cy + (sy ? 1 : -1), nz = cz + (sz ? 1 : -1);
        // squared distance from p to the neighbour slab along each axis (p is inside its own slab: 0)
        float ex = sx ? ((g.ox + (float)(cx + 1) * H) - margin) - px : px - ((g.ox + (float)cx * H) + margin);
        float ey = sy ? ((g.oy + (float)(cy + 1) * H) - margin) - py : py - ((g.oy + (float)cy * H) + margin);
        float ez = sz ? ((g.oz + (float)(cz + 1) * H) - margin) - pz : pz - ((g.oz + (float)cz * H) + margin);
        ex = fmaxf(ex, 0.0f), ey = fmaxf(ey, 0.0f), ez = fmaxf(ez, 0.0f);
        ex *= ex, ey *= ey, ez *= ez;
        const bool vx0 = cx >= 0 && cx < ncell, vx1 = nx >= 0 && nx < ncell;
        const bool vy0 = cy >= 0 && cy < ncell, vy1 = ny >= 0 && ny < ncell;
        const bool vz0 = cz >= 0 && cz < ncell, vz1 = nz >= 0 && nz < ncell;
        const float bound0 = fminf(best_d2, r2_prune) * 1.0001f + 1e-12f;
        // eight independent probes (dead cells predicated off), then their resolution
        uint4 e[8];
        uint32_t slot[8];
        uint32_t live = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool i = k & 1, j = (k >> 1) & 1, m = k >> 2;
            const float d2c = (i ? ex : 0.0f) + (j ? ey : 0.0f) + (m ? ez : 0.0f);
            const bool ok = (i ? vx1 : vx0) && (j ? vy1 : vy0) && (m ? vz1 : vz0) && d2c <= bound0;
            e[k] = make_uint4(0u, 0u, 0u, 0u);
            slot[k] = 0;
            if (ok) {
                const uint32_t x = (uint32_t)(i ? nx : cx), y = (uint32_t)(j ? ny : cy), z = (uint32_t)(m ? nz : cz);
                const uint32_t klo = cell_key_lo(x, y, z), khi = cell_key_hi(z, l);
                slot[k] = cell_hash(klo, khi) & g.mask;
                e[k] = ld_entry(&g.table[slot[k]]);
                live |= 1u << k;
                st.probe(0);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (!((live >> k) & 1u)) continue;
            const bool i = k & 1, j = (k >> 1) & 1, m = k >> 2;
            const uint32_t x = (uint32_t)(i ? nx : cx), y = (uint32_t)(j ? ny : cy), z = (uint32_t)(m ? nz : cz);
            uint32_t start, count, cmask;
            if (!probe_finish(g, slot[k], e[k], cell_key_lo(x, y, z), cell_key_hi(z, l), start, count, cmask)) continue;
            if (count <= (uint32_t)g.leaf_count || l == 0 || sp == kStack) S.range(nr++) = make_uint2(start, count);
            else S.stack(sp++) = pack_cell(x, y, z, l, cmask);
        }
        scan_ranges(g, px, py, pz, S, nr, best_d2, best_j, st);
        // descent through the dense cells of the block
        while (sp > 0) {
            const uint2 ce = S.stack(--sp);
            const int lv = (int)((ce.y >> 4) & 0xfu);
            const uint32_t cmask = (ce.y >> 8) & 0xffu;
            const int x = (int)(ce.x & 0xfffu), y = (int)((ce.x >> 12) & 0xfffu), z = (int)((ce.x >> 24) | ((ce.y & 0xfu) << 8));
            const float hc = 0.5f * g.h0 * (float)(1 << lv); // child size
            // per-axis squared distances to the two child slabs
            float ax[2], ay[2], az[2];
            {
                const float lox = g.ox + (float)(2 * x) * hc, mdx = g.ox + (float)(2 * x + 1) * hc, hix = g.ox + (float)(2 * x + 2) * hc;
                const float loy = g.oy + (float)(2 * y) * hc, mdy = g.oy + (float)(2 * y + 1) * hc, hiy = g.oy + (float)(2 * y + 2) * hc;
                const float loz = g.oz + (float)(2 * z) * hc, mdz = g.oz + (float)(2 * z + 1) * hc, hiz = g.oz + (float)(2 * z + 2) * hc;
                ax[0] = slab_dist(lox, mdx, px, margin), ax[1] = slab_dist(mdx, hix, px, margin);
                ay[0] = slab_dist(loy, mdy, py, margin), ay[1] = slab_dist(mdy, hiy, py, margin);
                az[0] = slab_dist(loz, mdz, pz, margin), az[1] = slab_dist(mdz, hiz, pz, margin);
                ax[0] *= ax[0], ax[1] *= ax[1], ay[0] *= ay[0], ay[1] *= ay[1], az[0] *= az[0], az[1] *= az[1];
            }
            const float bound = fminf(best_d2, r2_prune) * 1.0001f + 1e-12f;
            // the cell itself may have fallen behind the bound since it was pushed
            if (fminf(ax[0], ax[1]) + fminf(ay[0], ay[1]) + fminf(az[0], az[1]) > bound) continue;
            st.expand();
            if (nr + 8 > kRanges) scan_ranges(g, px, py, pz, S, nr, best_d2, best_j, st);
            const int near_child = (ax[1] < ax[0] ? 1 : 0) | (ay[1] < ay[0] ? 2 : 0) | (az[1] < az[0] ? 4 : 0);
            uint32_t pass = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int ch = (7 - k) ^ near_child; // the nearest octant comes last: pushed last, popped first
                const bool ok = ((cmask >> ch) & 1u) && (((ch & 1) ? ax[1] : ax[0]) + ((ch & 2) ? ay[1] : ay[0]) + ((ch & 4) ? az[1] : az[0]) <= bound);
                e[k] = make_uint4(0u, 0u, 0u, 0u);
                slot[k] = 0;
                if (ok) {
                    const uint32_t x2 = (uint32_t)(2 * x + (ch & 1)), y2 = (uint32_t)(2 * y + ((ch >> 1) & 1)),
                                   z2 = (uint32_t)(2 * z + (ch >> 2));
                    const uint32_t klo = cell_key_lo(x2, y2, z2), khi = cell_key_hi(z2, lv - 1);
                    slot[k] = cell_hash(klo, khi) & g.mask;
                    e[k] = ld_entry(&g.table[slot[k]]);
                    pass |= 1u << k;
                    st.probe(1);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (!((pass >> k) & 1u)) continue;
                const int ch = (7 - k) ^ near_child;
                const uint32_t x2 = (uint32_t)(2 * x + (ch & 1)), y2 = (uint32_t)(2 * y + ((ch >> 1) & 1)),
                               z2 = (uint32_t)(2 * z + (ch >> 2));
                uint32_t start, count, cm2;
                if (!probe_finish(g, slot[k], e[k], cell_key_lo(x2, y2, z2), cell_key_hi(z2, lv - 1), start, count, cm2)) continue;
                if (count <= (uint32_t)g.leaf_count || lv - 1 == 0 || sp == kStack) S.range(nr++) = make_uint2(start, count);
                else S.stack(sp++) = pack_cell(x2, y2, z2, lv - 1, cm2);
            }
            // small cells found so far tighten the bound before the next dense cell is opened
            scan_ranges(g, px, py, pz, S, nr, best_d2, best_j, st);
        }
        const float cover = (l == 0) ? 0.998f * 0.5f * g.h0 : 0.999f * 0.5f * H; // every closer target has been examined
        const float cover2 = cover * cover;
        if (best_d2 <= cover2) break;  // the best found is the global nearest
        if (cover2 >= r2_prune) break; // whole search radius examined
        if (l == L - 1) break;         // (n_levels is chosen so that the line above fires first)
    }
}

} // namespace mulls
