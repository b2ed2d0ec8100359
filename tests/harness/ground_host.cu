// TEST INFRASTRUCTURE. Runs the `__host__ __device__` work functions of mulls_b200/csrc/ground_core.cuh on the CPU with
// a one-lane "warp", in the order the kernels of kernels_ground.cuh launch them, so that the CPU test-suite can compare
// the product's per-point / per-cell logic with the oracle without a GPU. The device-only pieces (bounding-box
// reduction, ordered height sum, library sort / scans, the final down-sampling) are restated in plain loops here and
// are covered by the GPU tests. Built by tests/test_ground.py with nvcc (host code only; no CUDA call is made).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../mulls_b200/csrc/ground_core.cuh"

using namespace mulls;

static std::vector<uint32_t> draw_table() {
    std::vector<uint32_t> t(kSacDraws);
    uint32_t mt[624];
    mt[0] = 12345u;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    int idx = 624;
    for (int k = 0; k < kSacDraws; ++k) {
        if (idx >= 624) {
            for (int i = 0; i < 624; ++i) {
                const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        t[k] = y;
    }
    return t;
}

extern "C" int gfh_run(const float *rows48, size_t n, const mulls_ground_params *params, mulls_ground_out *out) {
    out->n_ground = out->n_ground_down = out->n_unground = 0;
    if (n == 0) return 0;
    GfArgs A;
    std::memset(&A, 0, sizeof(A));
    A.P = *params;
    A.n = (uint32_t)n;
    A.rows = reinterpret_cast<const float4 *>(rows48);
    GfState S;
    std::memset(&S, 0, sizeof(S));
    A.st = &S;
    // k_gf_bbox + k_gf_setup
    float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
    float sum_height = 0.001f;
    int count = 0;
    for (size_t j = 0; j < n; ++j) {
        const float4 a = A.rows[3 * j];
        if (a.x == a.x) mnx = std::min(mnx, a.x), mxx = std::max(mxx, a.x);
        if (a.y == a.y) mny = std::min(mny, a.y), mxy = std::max(mxy, a.y);
        if (j % 100 == 0) sum_height += a.z, ++count;
    }
    S.appro_mean_height = sum_height / count;
    S.non_ground_height_thre = S.appro_mean_height + A.P.max_ground_height;
    S.min_x = mnx, S.min_y = mny, S.max_x = mxx, S.max_y = mxy;
    S.row = (int)std::ceil((S.max_y - S.min_y) / (double)A.P.grid_resolution);
    S.col = (int)std::ceil((S.max_x - S.min_x) / (double)A.P.grid_resolution);
    S.num_grid = S.row * S.col;
    const int G = S.num_grid;
    if (G <= 0) return 0;
    std::vector<uint32_t> key(n), idx(n), key_s(n), idx_s(n), high_flag(n), high_pos(n);
    std::vector<int> cell_all(n), shuf(n);
    std::vector<uint8_t> decision(n), inl(n);
    std::vector<float4> cand(n), og(3 * n), ogd(3 * n), ou(3 * n), cell_normal(G);
    std::vector<uint32_t> cs(G, 0), ce(G, 0), cng(G), cnu(G), cog(G), cou(G);
    std::vector<float> minz(G), nbz(G), oth(G);
    std::vector<int> rel(G);
    std::vector<uint32_t> draws = draw_table();
    A.key = key.data(), A.idx = idx.data(), A.key_s = key_s.data(), A.idx_s = idx_s.data(), A.cell_all = cell_all.data();
    A.high_flag = high_flag.data(), A.high_pos = high_pos.data(), A.decision = decision.data(), A.cand = cand.data();
    A.shuf = shuf.data(), A.inl = inl.data(), A.cell_start = cs.data(), A.cell_end = ce.data(), A.min_z = minz.data();
    A.neighbor_min_z = nbz.data(), A.outlier_thre = oth.data(), A.reliable = rel.data(), A.cell_normal = cell_normal.data();
    A.cell_ng = cng.data(), A.cell_nu = cnu.data(), A.cell_og = cog.data(), A.cell_ou = cou.data(), A.draws = draws.data();
    A.out_ground = og.data(), A.out_ground_down = ogd.data(), A.out_unground = ou.data();
    for (uint32_t j = 0; j < n; ++j) gf_assign_point(A, j);
    {   // stable sort by cell (the radix sort of the device path)
        std::vector<uint32_t> perm(n);
        std::iota(perm.begin(), perm.end(), 0u);
        std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        for (size_t i = 0; i < n; ++i) key_s[i] = key[perm[i]], idx_s[i] = idx[perm[i]];
    }
    for (uint32_t i = 0; i < n; ++i) gf_mark_bounds(A, i);
    for (int c = 0; c < G; ++c) gf_cell_min(A, c);
    for (int c = 0; c < G; ++c) gf_cell_neighbors(A, c);
    for (uint32_t j = 0; j < n; ++j) gf_high_point(A, j);
    uint32_t run = 0;
    for (size_t j = 0; j < n; ++j) high_pos[j] = run, run += high_flag[j];
    S.n_high = run;
    for (uint32_t j = 0; j < n; ++j) gf_high_emit(A, j);
    for (int c = 0; c < G; ++c) gf_cell_decide(A, c);
    uint32_t rg = 0, ru = 0;
    for (int c = 0; c < G; ++c) cog[c] = rg, rg += cng[c], cou[c] = ru, ru += cnu[c];
    S.n_ground = rg, S.n_unground = S.n_high + ru;
    for (int c = 0; c < G; ++c) gf_cell_emit(A, c);
    // cloud_ground_down without fixed_num_downsampling (the sampled variant is device-only: block_sample_append)
    uint32_t nd = 0;
    for (uint32_t i = 0; i < S.n_ground; i += (uint32_t)A.P.ground_random_down_down_rate, ++nd)
        for (int k = 0; k < 3; ++k) ogd[3 * (size_t)nd + k] = og[3 * (size_t)i + k];
    S.n_ground_down = nd;
    out->n_ground = S.n_ground, out->n_ground_down = S.n_ground_down, out->n_unground = S.n_unground;
    if (out->ground) std::memcpy(out->ground, og.data(), 48 * (size_t)S.n_ground);
    if (out->ground_down) std::memcpy(out->ground_down, ogd.data(), 48 * (size_t)S.n_ground_down);
    if (out->unground) std::memcpy(out->unground, ou.data(), 48 * (size_t)S.n_unground);
    return 0;
}

// CFilter::voxel_downsample through the product's vx_setup / vx_key (the sort / scan / gather of the device path are
// plain loops here)
extern "C" int gfh_voxel(const float *rows48, size_t n, float voxel_size, float *out, size_t *n_out) {
    *n_out = 0;
    if (n == 0) return 0;
    const float4 *rows = reinterpret_cast<const float4 *>(rows48);
    VxState S;
    std::memset(&S, 0, sizeof(S));
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t j = 0; j < n; ++j) {
        const float4 a = rows[3 * j];
        mn[0] = std::min(mn[0], a.x), mn[1] = std::min(mn[1], a.y), mn[2] = std::min(mn[2], a.z);
        mx[0] = std::max(mx[0], a.x), mx[1] = std::max(mx[1], a.y), mx[2] = std::max(mx[2], a.z);
    }
    for (int d = 0; d < 3; ++d) S.bb[d] = gf_ord(mn[d]), S.bb[3 + d] = gf_ord(mx[d]);
    vx_setup(S, voxel_size);
    std::vector<unsigned long long> key(n);
    std::vector<uint32_t> perm(n);
    for (size_t j = 0; j < n; ++j) key[j] = vx_key(S, rows[3 * j]);
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    size_t m = 0;
    for (size_t i = 0; i < n; ++i)
        if (i == 0 || key[perm[i]] != key[perm[i - 1]]) std::memcpy(out + 12 * (m++), rows48 + 12 * (size_t)perm[i], 48);
    *n_out = m;
    return 0;
}
