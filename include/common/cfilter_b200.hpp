// Drop-in shims: lo::CFilter<PointT>::classify_nground_pts (include/common/cfilter.hpp:2058-2290), fast_ground_filter
// (:1658-2036) and voxel_downsample (:83-165) over the mulls_b200 C-ABI. A MULLS maintainer replaces the BODY of classify_nground_pts by
//
//     return lo::b200::classify_nground_pts<PointT>(cloud_in, cloud_pillar, ... );      // all arguments forwarded
//
// and CFilter::extract_semantic_pts (:2295-2413) keeps calling it unchanged. Same names, order, types and defaults as
// :2060-2081. The output clouds are expected empty at the call (as extract_semantic_pts passes them): results are
// appended. Differences: every random_downsample_pcl inside is a reproducible uniform sample;
// std::sort's unspecified order of equal NMS scores is fixed to "first pushed first".
#ifndef MULLS_B200_CFILTER_SHIM_HPP
#define MULLS_B200_CFILTER_SHIM_HPP

#include <cfloat>
#include <vector>

#include "common/cregistration_b200.hpp" // thread_context, view_of
#include "mulls_b200/abi.h"

namespace lo {
namespace b200 {

template <typename PointT>
bool classify_nground_pts(typename pcl::PointCloud<PointT>::Ptr &cloud_in, typename pcl::PointCloud<PointT>::Ptr &cloud_pillar,
                          typename pcl::PointCloud<PointT>::Ptr &cloud_beam, typename pcl::PointCloud<PointT>::Ptr &cloud_facade,
                          typename pcl::PointCloud<PointT>::Ptr &cloud_roof, typename pcl::PointCloud<PointT>::Ptr &cloud_pillar_down,
                          typename pcl::PointCloud<PointT>::Ptr &cloud_beam_down, typename pcl::PointCloud<PointT>::Ptr &cloud_facade_down,
                          typename pcl::PointCloud<PointT>::Ptr &cloud_roof_down, typename pcl::PointCloud<PointT>::Ptr &cloud_vertex,
                          float neighbor_searching_radius, int neighbor_k, int neigh_k_min, int pca_down_rate, float edge_thre,
                          float planar_thre, float edge_thre_down, float planar_thre_down, int extract_vertex_points_method,
                          float curvature_thre, float vertex_curvature_non_max_radius, float linear_vertical_sin_high_thre,
                          float linear_vertical_sin_low_thre, float planar_vertical_sin_high_thre,
                          float planar_vertical_sin_low_thre, bool fixed_num_downsampling = false, int pillar_down_fixed_num = 200,
                          int facade_down_fixed_num = 800, int beam_down_fixed_num = 200, int roof_down_fixed_num = 100,
                          int unground_down_fixed_num = 20000, float beam_height_max = FLT_MAX, float roof_height_min = -FLT_MAX,
                          float feature_pts_ratio_guess = 0.3, bool sharpen_with_nms = true,
                          bool use_distance_adaptive_pca = false) {
    static_assert(sizeof(PointT) == 48, "the C-ABI consumes pcl::PointXYZINormal rows (48 bytes)");
    static thread_local uint32_t call_seed = 0;
    mulls_classify_params p;
    mulls_classify_default_params(&p);
    p.neighbor_searching_radius = neighbor_searching_radius;
    p.neighbor_k = neighbor_k;
    p.neigh_k_min = neigh_k_min;
    p.pca_down_rate = pca_down_rate;
    p.edge_thre = edge_thre;
    p.planar_thre = planar_thre;
    p.edge_thre_down = edge_thre_down;
    p.planar_thre_down = planar_thre_down;
    p.extract_vertex_points_method = extract_vertex_points_method;
    p.curvature_thre = curvature_thre;
    p.vertex_curvature_non_max_radius = vertex_curvature_non_max_radius;
    p.linear_vertical_sin_high_thre = linear_vertical_sin_high_thre;
    p.linear_vertical_sin_low_thre = linear_vertical_sin_low_thre;
    p.planar_vertical_sin_high_thre = planar_vertical_sin_high_thre;
    p.planar_vertical_sin_low_thre = planar_vertical_sin_low_thre;
    p.fixed_num_downsampling = fixed_num_downsampling;
    p.pillar_down_fixed_num = pillar_down_fixed_num;
    p.facade_down_fixed_num = facade_down_fixed_num;
    p.beam_down_fixed_num = beam_down_fixed_num;
    p.roof_down_fixed_num = roof_down_fixed_num;
    p.unground_down_fixed_num = unground_down_fixed_num;
    p.beam_height_max = beam_height_max;
    p.roof_height_min = roof_height_min;
    p.feature_pts_ratio_guess = feature_pts_ratio_guess;
    p.sharpen_with_nms = sharpen_with_nms;
    p.use_distance_adaptive_pca = use_distance_adaptive_pca;
    p.random_seed = call_seed++;

    const size_t n = cloud_in->points.size();
    mulls_ctx *ctx = thread_context(1, n);
    std::vector<std::vector<PointT>> rows(MULLS_OUT_COUNT, std::vector<PointT>(n ? n : 1));
    mulls_classify_out out;
    for (int k = 0; k < MULLS_OUT_COUNT; ++k) out.rows[k] = reinterpret_cast<float *>(rows[k].data()), out.n[k] = 0;
    out.cap = n ? n : 1;
    if (!ctx || mulls_classify_nground(ctx, view_of<PointT>(cloud_in), &p, &out) != MULLS_OK) {
        LOG(ERROR) << "mulls_b200: " << mulls_last_error(ctx);
        return false;
    }
    typename pcl::PointCloud<PointT>::Ptr *dst[MULLS_OUT_COUNT] = {&cloud_pillar,      &cloud_beam,        &cloud_facade,
                                                                   &cloud_roof,        &cloud_pillar_down, &cloud_beam_down,
                                                                   &cloud_facade_down, &cloud_roof_down,   &cloud_vertex,
                                                                   &cloud_in};
    for (int k = 0; k < MULLS_OUT_COUNT; ++k) {
        if (k == MULLS_OUT_UNGROUND) (*dst[k])->points.clear(); // cloud_in is rewritten in place by the reference
        (*dst[k])->points.insert((*dst[k])->points.end(), rows[k].begin(), rows[k].begin() + out.n[k]);
    }
    return true;
}

// lo::CFilter<PointT>::fast_ground_filter (include/common/cfilter.hpp:1658-2036), same names, order, types and defaults
// as :1658-1672. cloud_curb / detect_curb_or_not are accepted and unused (the reference's curb code is `#if 0`, :1987).
// The output clouds are appended to, as the reference does. estimate_ground_normal_method 1 / 2 return false.
template <typename PointT>
bool fast_ground_filter(const typename pcl::PointCloud<PointT>::Ptr &cloud_in, typename pcl::PointCloud<PointT>::Ptr &cloud_ground,
                        typename pcl::PointCloud<PointT>::Ptr &cloud_ground_down, typename pcl::PointCloud<PointT>::Ptr &cloud_unground,
                        typename pcl::PointCloud<PointT>::Ptr &cloud_curb, int min_grid_pt_num, float grid_resolution,
                        float max_height_difference, float neighbor_height_diff, float max_ground_height,
                        int ground_random_down_rate, int ground_random_down_down_rate, int nonground_random_down_rate,
                        int reliable_neighbor_grid_num_thre, int estimate_ground_normal_method, float normal_estimation_radius,
                        int distance_weight_downsampling_method, float standard_distance, bool fixed_num_downsampling = false,
                        int down_ground_fixed_num = 1000, bool detect_curb_or_not = false, float intensity_thre = FLT_MAX,
                        bool apply_grid_wise_outlier_filter = false, float outlier_std_scale = 3.0) {
    static_assert(sizeof(PointT) == 48, "the C-ABI consumes pcl::PointXYZINormal rows (48 bytes)");
    static thread_local uint32_t call_seed = 0;
    (void)cloud_curb;
    (void)detect_curb_or_not;
    mulls_ground_params p;
    mulls_ground_default_params(&p);
    p.min_grid_pt_num = min_grid_pt_num;
    p.grid_resolution = grid_resolution;
    p.max_height_difference = max_height_difference;
    p.neighbor_height_diff = neighbor_height_diff;
    p.max_ground_height = max_ground_height;
    p.ground_random_down_rate = ground_random_down_rate;
    p.ground_random_down_down_rate = ground_random_down_down_rate;
    p.nonground_random_down_rate = nonground_random_down_rate;
    p.reliable_neighbor_grid_num_thre = reliable_neighbor_grid_num_thre;
    p.estimate_ground_normal_method = estimate_ground_normal_method;
    p.normal_estimation_radius = normal_estimation_radius;
    p.distance_weight_downsampling_method = distance_weight_downsampling_method;
    p.standard_distance = standard_distance;
    p.fixed_num_downsampling = fixed_num_downsampling;
    p.down_ground_fixed_num = down_ground_fixed_num;
    p.intensity_thre = intensity_thre;
    p.apply_grid_wise_outlier_filter = apply_grid_wise_outlier_filter;
    p.outlier_std_scale = outlier_std_scale;
    p.random_seed = call_seed++;
    const size_t n = cloud_in->points.size();
    mulls_ctx *ctx = thread_context(1, n);
    std::vector<PointT> g(n ? n : 1), gd(n ? n : 1), u(n ? n : 1);
    mulls_ground_out out;
    out.ground = reinterpret_cast<float *>(g.data()), out.ground_down = reinterpret_cast<float *>(gd.data());
    out.unground = reinterpret_cast<float *>(u.data());
    out.cap = n ? n : 1;
    out.n_ground = out.n_ground_down = out.n_unground = 0;
    if (!ctx || mulls_fast_ground_filter(ctx, view_of<PointT>(cloud_in), &p, &out) != MULLS_OK) {
        LOG(ERROR) << "mulls_b200: " << mulls_last_error(ctx);
        return false;
    }
    cloud_ground->points.insert(cloud_ground->points.end(), g.begin(), g.begin() + out.n_ground);
    cloud_ground_down->points.insert(cloud_ground_down->points.end(), gd.begin(), gd.begin() + out.n_ground_down);
    cloud_unground->points.insert(cloud_unground->points.end(), u.begin(), u.begin() + out.n_unground);
    return true;
}

// lo::CFilter<PointT>::voxel_downsample (include/common/cfilter.hpp:83-165)
template <typename PointT>
bool voxel_downsample(const typename pcl::PointCloud<PointT>::Ptr &cloud_in, typename pcl::PointCloud<PointT>::Ptr &cloud_out,
                      float voxel_size) {
    static_assert(sizeof(PointT) == 48, "the C-ABI consumes pcl::PointXYZINormal rows (48 bytes)");
    if (voxel_size < 0.001) { // :89-97: the reference shares the input cloud and reports "disabled"
        cloud_out = cloud_in;
        return false;
    }
    const size_t n = cloud_in->points.size();
    mulls_ctx *ctx = thread_context(1, n);
    std::vector<PointT> rows(n ? n : 1);
    size_t n_out = 0;
    if (!ctx || mulls_voxel_downsample(ctx, view_of<PointT>(cloud_in), voxel_size, reinterpret_cast<float *>(rows.data()),
                                       rows.size(), &n_out) != MULLS_OK) {
        LOG(ERROR) << "mulls_b200: " << mulls_last_error(ctx);
        return false;
    }
    cloud_out->points.insert(cloud_out->points.end(), rows.begin(), rows.begin() + n_out);
    return true;
}

// lo::CFilter<PointT>::extract_semantic_pts (include/common/cfilter.hpp:2295-2413): same names, order, types and
// defaults as :2295-2318. Covers :2346-2399 — voxel_downsample, pc_sketch, fast_ground_filter, classify_nground_pts and
// the feature count — with the three heavy stages chained in HBM (mulls_extract_semantic_pts): replace those lines of
// the member by a call forwarding every argument. The pre-filters on pc_raw (:2328-2342) and
// update_parameters_self_adaptive (:2406-2410) are CFilter members and stay where they are, before / after the call.
template <typename PointT, typename BlockPtr>
bool extract_semantic_pts(BlockPtr in_block, float vf_downsample_resolution, float gf_grid_resolution, float gf_max_grid_height_diff,
                          float gf_neighbor_height_diff, float gf_max_ground_height, int &gf_down_rate_ground,
                          int &gf_downsample_rate_nonground, float pca_neighbor_radius, int pca_neighbor_k, float edge_thre,
                          float planar_thre, float curvature_thre, float edge_thre_down, float planar_thre_down,
                          bool use_distance_adaptive_pca = false, int distance_inverse_sampling_method = 0,
                          float standard_distance = 15.0, int estimate_ground_normal_method = 3,
                          float normal_estimation_radius = 2.0, bool use_adpative_parameters = false,
                          bool apply_scanner_filter = false, bool extract_curb_or_not = false,
                          int extract_vertex_points_method = 2, int gf_grid_pt_num_thre = 8,
                          int gf_reliable_neighbor_grid_thre = 0, int gf_down_down_rate_ground = 2, int pca_neighbor_k_min = 8,
                          int pca_down_rate = 1, float intensity_thre = FLT_MAX, float linear_vertical_sin_high_thre = 0.94,
                          float linear_vertical_sin_low_thre = 0.17, float planar_vertical_sin_high_thre = 0.98,
                          float planar_vertical_sin_low_thre = 0.34, bool sharpen_with_nms_on = true,
                          bool fixed_num_downsampling = false, int ground_down_fixed_num = 500, int pillar_down_fixed_num = 200,
                          int facade_down_fixed_num = 800, int beam_down_fixed_num = 200, int roof_down_fixed_num = 200,
                          int unground_down_fixed_num = 20000, float beam_height_max = FLT_MAX, float roof_height_min = 0.0,
                          float approx_scanner_height = 2.0, float underground_thre = -7.0, float feature_pts_ratio_guess = 0.3,
                          bool semantic_assisted = false, bool apply_roi_filtering = false, float roi_min_y = 0.0,
                          float roi_max_y = 0.0) {
    static_assert(sizeof(PointT) == 48, "the C-ABI consumes pcl::PointXYZINormal rows (48 bytes)");
    static thread_local uint32_t call_seed = 0;
    (void)use_adpative_parameters, (void)extract_curb_or_not, (void)approx_scanner_height, (void)underground_thre;
    (void)semantic_assisted, (void)apply_roi_filtering, (void)roi_min_y, (void)roi_max_y;
    mulls_extract_params P;
    P.vf_downsample_resolution = vf_downsample_resolution;
    mulls_ground_default_params(&P.ground);
    P.ground.min_grid_pt_num = gf_grid_pt_num_thre;
    P.ground.grid_resolution = gf_grid_resolution;
    P.ground.max_height_difference = gf_max_grid_height_diff;
    P.ground.neighbor_height_diff = gf_neighbor_height_diff;
    P.ground.max_ground_height = gf_max_ground_height;
    P.ground.ground_random_down_rate = gf_down_rate_ground;
    P.ground.ground_random_down_down_rate = gf_down_down_rate_ground;
    P.ground.nonground_random_down_rate = gf_downsample_rate_nonground;
    P.ground.reliable_neighbor_grid_num_thre = gf_reliable_neighbor_grid_thre;
    P.ground.estimate_ground_normal_method = estimate_ground_normal_method;
    P.ground.normal_estimation_radius = normal_estimation_radius;
    P.ground.distance_weight_downsampling_method = distance_inverse_sampling_method;
    P.ground.standard_distance = standard_distance;
    P.ground.fixed_num_downsampling = fixed_num_downsampling;
    P.ground.down_ground_fixed_num = ground_down_fixed_num;
    P.ground.intensity_thre = intensity_thre;
    P.ground.apply_grid_wise_outlier_filter = apply_scanner_filter; // the argument extract_semantic_pts passes there (:2361)
    P.ground.random_seed = call_seed;
    mulls_classify_default_params(&P.classify);
    P.classify.neighbor_searching_radius = pca_neighbor_radius;
    P.classify.neighbor_k = pca_neighbor_k;
    P.classify.neigh_k_min = pca_neighbor_k_min;
    P.classify.pca_down_rate = pca_down_rate;
    P.classify.edge_thre = edge_thre;
    P.classify.planar_thre = planar_thre;
    P.classify.edge_thre_down = edge_thre_down;
    P.classify.planar_thre_down = planar_thre_down;
    P.classify.extract_vertex_points_method = extract_vertex_points_method;
    P.classify.curvature_thre = curvature_thre;
    P.classify.vertex_curvature_non_max_radius = 1.5 * pca_neighbor_radius; // :2363
    P.classify.linear_vertical_sin_high_thre = linear_vertical_sin_high_thre;
    P.classify.linear_vertical_sin_low_thre = linear_vertical_sin_low_thre;
    P.classify.planar_vertical_sin_high_thre = planar_vertical_sin_high_thre;
    P.classify.planar_vertical_sin_low_thre = planar_vertical_sin_low_thre;
    P.classify.fixed_num_downsampling = fixed_num_downsampling;
    P.classify.pillar_down_fixed_num = pillar_down_fixed_num;
    P.classify.facade_down_fixed_num = facade_down_fixed_num;
    P.classify.beam_down_fixed_num = beam_down_fixed_num;
    P.classify.roof_down_fixed_num = roof_down_fixed_num;
    P.classify.unground_down_fixed_num = unground_down_fixed_num;
    P.classify.beam_height_max = beam_height_max;
    P.classify.roof_height_min = roof_height_min;
    P.classify.feature_pts_ratio_guess = feature_pts_ratio_guess;
    P.classify.sharpen_with_nms = sharpen_with_nms_on;
    P.classify.use_distance_adaptive_pca = use_distance_adaptive_pca;
    P.classify.random_seed = call_seed++;

    const size_t n = in_block->pc_raw->points.size();
    mulls_ctx *ctx = thread_context(1, n);
    const size_t cap = n ? n : 1;
    std::vector<PointT> down(cap), ground(cap), ground_down(cap);
    std::vector<std::vector<PointT>> rows(MULLS_OUT_COUNT, std::vector<PointT>(cap));
    mulls_extract_out out;
    out.pc_down = reinterpret_cast<float *>(down.data());
    out.pc_ground = reinterpret_cast<float *>(ground.data());
    out.pc_ground_down = reinterpret_cast<float *>(ground_down.data());
    out.cap = cap;
    out.n_down = out.n_ground = out.n_ground_down = 0;
    for (int k = 0; k < MULLS_OUT_COUNT; ++k) out.cls.rows[k] = reinterpret_cast<float *>(rows[k].data()), out.cls.n[k] = 0;
    out.cls.cap = cap;
    if (!ctx || mulls_extract_semantic_pts(ctx, view_of<PointT>(in_block->pc_raw), &P, &out) != MULLS_OK) {
        LOG(ERROR) << "mulls_b200: " << mulls_last_error(ctx);
        return false;
    }
    auto append = [](typename pcl::PointCloud<PointT>::Ptr &dst, const std::vector<PointT> &src, size_t cnt) {
        dst->points.insert(dst->points.end(), src.begin(), src.begin() + cnt);
    };
    if (vf_downsample_resolution < 0.001) in_block->pc_down = in_block->pc_raw; // :92 the reference shares the cloud
    else append(in_block->pc_down, down, out.n_down);
    {   // :2348 random_downsample(pc_down, pc_sketch, size / 1024 + 1): every k-th point (:713-728)
        const int ratio = (int)(in_block->pc_down->points.size() / 1024 + 1);
        if (ratio > 1) {
            in_block->pc_sketch->points.clear();
            for (size_t i = 0; i < in_block->pc_down->points.size(); i += (size_t)ratio)
                in_block->pc_sketch->points.push_back(in_block->pc_down->points[i]);
        }
    }
    append(in_block->pc_ground, ground, out.n_ground);
    append(in_block->pc_ground_down, ground_down, out.n_ground_down);
    typename pcl::PointCloud<PointT>::Ptr *dst[MULLS_OUT_COUNT] = {
        &in_block->pc_pillar,      &in_block->pc_beam,      &in_block->pc_facade,      &in_block->pc_roof,   &in_block->pc_pillar_down,
        &in_block->pc_beam_down,   &in_block->pc_facade_down, &in_block->pc_roof_down, &in_block->pc_vertex, &in_block->pc_unground};
    for (int k = 0; k < MULLS_OUT_COUNT; ++k) append(*dst[k], rows[k], out.cls.n[k]);
    in_block->down_feature_point_num = in_block->pc_ground_down->points.size() + in_block->pc_pillar_down->points.size() +
                                       in_block->pc_beam_down->points.size() + in_block->pc_facade_down->points.size() +
                                       in_block->pc_roof_down->points.size() + in_block->pc_vertex->points.size(); // :2398-2399
    return true;
}

} // namespace b200
} // namespace lo
#endif
