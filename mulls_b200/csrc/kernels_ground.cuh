// Kernels of mulls_fast_ground_filter (CFilter::fast_ground_filter, cfilter.hpp:1658-2036): thin launch wrappers around
// the per-point / per-cell functions of ground_core.cuh, plus the statistics that need a fixed order.
#pragma once
#include "ground_core.cuh"
#include "kernels_classify.cuh" // block_sample_append (random_downsample_pcl)

namespace mulls {

constexpr int kGfBlock = 256;

// bounding box (get_cloud_bbx, utility.hpp:817-847): min / max of floats are order independent
__global__ void __launch_bounds__(kGfBlock) k_gf_bbox(GfArgs A) {
    const uint32_t j = blockIdx.x * kGfBlock + threadIdx.x;
    float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
    if (j < A.n) {
        const float4 a = A.rows[3 * (size_t)j];
        if (a.x == a.x) mnx = mxx = a.x;
        if (a.y == a.y) mny = mxy = a.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&A.st->bb[0], gf_ord(mnx));
        atomicMin(&A.st->bb[1], gf_ord(mny));
        atomicMax(&A.st->bb[2], gf_ord(mxx));
        atomicMax(&A.st->bb[3], gf_ord(mxy));
    }
}

// :1686-1700 the approximate mean height: a FLOAT sum over every 100th point in index order (order dependent), so one
// warp fetches 32 samples at a time and adds them in order; then the grid geometry (:1707-1713).
__global__ void k_gf_setup(GfArgs A) {
    const int lane = threadIdx.x;
    float sum_height = 0.001f;
    int count = 0;
    const uint32_t n_s = (A.n + 99) / 100; // j = 0, 100, 200, ...
    for (uint32_t base = 0; base < n_s; base += 32) {
        const uint32_t s = base + lane;
        const float z = (s < n_s) ? A.rows[3 * (size_t)(s * 100)].z : 0.0f;
        const int m = (int)min(32u, n_s - base);
        for (int k = 0; k < m; ++k) sum_height += __shfl_sync(0xffffffffu, z, k);
        count += m;
    }
    if (lane == 0) {
        GfState &S = *A.st;
        S.sum_height = sum_height;
        S.count_checkpoint = count;
        S.appro_mean_height = sum_height / count;
        S.non_ground_height_thre = S.appro_mean_height + A.P.max_ground_height;
        S.min_x = (double)gf_unord(S.bb[0]), S.min_y = (double)gf_unord(S.bb[1]);
        S.max_x = (double)gf_unord(S.bb[2]), S.max_y = (double)gf_unord(S.bb[3]);
        const double r = ceil((S.max_y - S.min_y) / (double)A.P.grid_resolution);
        const double c = ceil((S.max_x - S.min_x) / (double)A.P.grid_resolution);
        // (int) of an out-of-range double is undefined in the reference; anything beyond 2^15 per side is refused by the host
        S.row = (r >= 0.0 && r < 1e9) ? (int)r : -1;
        S.col = (c >= 0.0 && c < 1e9) ? (int)c : -1;
        S.num_grid = (S.row >= 0 && S.col >= 0 && (long long)S.row * S.col < (1ll << 31)) ? S.row * S.col : -1;
    }
}

__global__ void __launch_bounds__(kGfBlock) k_gf_assign(GfArgs A) {
    const uint32_t j = blockIdx.x * kGfBlock + threadIdx.x;
    if (j < A.n) gf_assign_point(A, j);
}
__global__ void __launch_bounds__(kGfBlock) k_gf_bounds(GfArgs A) {
    const uint32_t i = blockIdx.x * kGfBlock + threadIdx.x;
    if (i < A.n) gf_mark_bounds(A, i);
}
// one warp per cell
__global__ void __launch_bounds__(kGfBlock) k_gf_cell_min(GfArgs A, int num_grid) {
    const int c = (int)((blockIdx.x * (unsigned)kGfBlock + threadIdx.x) >> 5);
    if (c < num_grid) gf_cell_min(A, c);
}
__global__ void __launch_bounds__(kGfBlock) k_gf_neighbors(GfArgs A, int num_grid) {
    const int m = (int)(blockIdx.x * (unsigned)kGfBlock + threadIdx.x);
    if (m < num_grid) gf_cell_neighbors(A, m);
}
__global__ void __launch_bounds__(kGfBlock) k_gf_high(GfArgs A) {
    const uint32_t j = blockIdx.x * kGfBlock + threadIdx.x;
    if (j < A.n) gf_high_point(A, j);
}
__global__ void __launch_bounds__(kGfBlock) k_gf_high_emit(GfArgs A) {
    const uint32_t j = blockIdx.x * kGfBlock + threadIdx.x;
    if (j < A.n) gf_high_emit(A, j);
    if (j + 1 == A.n) A.st->n_high = A.high_pos[j] + A.high_flag[j];
}
__global__ void __launch_bounds__(kGfBlock) k_gf_cell_decide(GfArgs A, int num_grid) {
    const int c = (int)((blockIdx.x * (unsigned)kGfBlock + threadIdx.x) >> 5);
    if (c < num_grid) gf_cell_decide(A, c);
}
// after the two exclusive scans: totals
__global__ void k_gf_totals(GfArgs A, int num_grid) {
    GfState &S = *A.st;
    S.n_ground = A.cell_og[num_grid - 1] + A.cell_ng[num_grid - 1];
    S.n_unground = S.n_high + A.cell_ou[num_grid - 1] + A.cell_nu[num_grid - 1];
}
__global__ void __launch_bounds__(kGfBlock) k_gf_cell_emit(GfArgs A, int num_grid) {
    const int c = (int)((blockIdx.x * (unsigned)kGfBlock + threadIdx.x) >> 5);
    if (c < num_grid) gf_cell_emit(A, c);
}

// cloud_ground_down (:1955-1964): every ground_random_down_down_rate-th ground point, or random_downsample_pcl
__global__ void __launch_bounds__(kClsBlock) k_gf_down(GfArgs A) {
    __shared__ SampleShared S;
    const uint32_t n = A.st->n_ground;
    if (!A.P.fixed_num_downsampling) {
        const uint32_t r = (uint32_t)A.P.ground_random_down_down_rate;
        const uint32_t m = (n + r - 1) / r;
        for (uint32_t k = threadIdx.x; k < m; k += kClsBlock) {
            const float4 *s = A.out_ground + 3 * (size_t)k * r;
            float4 *o = A.out_ground_down + 3 * (size_t)k;
            o[0] = s[0], o[1] = s[1], o[2] = s[2];
        }
        if (threadIdx.x == 0) A.st->n_ground_down = m;
        return;
    }
    if (threadIdx.x == 0) S.total = 0;
    __syncthreads();
    block_sample_append(A.out_ground, n, A.P.down_ground_fixed_num, A.P.random_seed, 30u, A.out_ground_down, S);
    if (threadIdx.x == 0) A.st->n_ground_down = S.total;
}

// ---- voxel_downsample (cfilter.hpp:83-165) -----------------------------------------------------------------------------
struct VxArgs {
    uint32_t n;
    float voxel_size;
    const float4 *rows;
    VxState *st;
    unsigned long long *key, *key_s;
    uint32_t *idx, *idx_s, *head, *pos;
    float4 *out;
};
__global__ void __launch_bounds__(kGfBlock) k_vx_bbox(VxArgs V) {
    const uint32_t j = blockIdx.x * kGfBlock + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (j < V.n) {
        const float4 a = V.rows[3 * (size_t)j];
        mn[0] = mx[0] = a.x, mn[1] = mx[1] = a.y, mn[2] = mx[2] = a.z;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&V.st->bb[d], gf_ord(mn[d]));
            atomicMax(&V.st->bb[3 + d], gf_ord(mx[d]));
        }
}
__global__ void k_vx_setup(VxArgs V) { vx_setup(*V.st, V.voxel_size); }
__global__ void __launch_bounds__(kGfBlock) k_vx_keys(VxArgs V) {
    const uint32_t j = blockIdx.x * kGfBlock + threadIdx.x;
    if (j < V.n) {
        V.key[j] = vx_key(*V.st, V.rows[3 * (size_t)j]);
        V.idx[j] = j;
    }
}
__global__ void __launch_bounds__(kGfBlock) k_vx_heads(VxArgs V) {
    const uint32_t i = blockIdx.x * kGfBlock + threadIdx.x;
    if (i < V.n) V.head[i] = (i == 0 || V.key_s[i] != V.key_s[i - 1]) ? 1u : 0u;
}
__global__ void __launch_bounds__(kGfBlock) k_vx_gather(VxArgs V) {
    const uint32_t i = blockIdx.x * kGfBlock + threadIdx.x;
    if (i >= V.n) return;
    if (V.head[i]) {
        const float4 *r = V.rows + 3 * (size_t)V.idx_s[i];
        float4 *o = V.out + 3 * (size_t)V.pos[i];
        o[0] = r[0], o[1] = r[1], o[2] = r[2];
    }
    if (i + 1 == V.n) V.st->n_out = V.pos[i] + V.head[i];
}

} // namespace mulls
