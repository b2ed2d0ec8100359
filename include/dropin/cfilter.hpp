// DROP-IN replacement of lo::CFilter<PointT> (reference: include/common/cfilter.hpp). Same mechanism as
// dropin/cregistration.hpp: this directory goes BEFORE the reference's include/common on the include path; the
// reference's own cfilter.hpp is pulled in with its class renamed to CFilter_reference, and lo::CFilter<PointT> is
// defined here as a class derived from it whose extract_semantic_pts (cfilter.hpp:2295-2318: same name, argument
// order, types and defaults — all fifty of them), voxel_downsample (:83), fast_ground_filter (:1658-1672) and
// classify_nground_pts (:2058-2081) run on the GPU through the C-ABI. Every other member (dist_filter,
// non_max_suppress, random_downsample, apply_motion_compensation, get_cloud_bbx, ... SURVEY.md section 8b) is
// inherited from the reference. test/mulls_slam.cpp:360-377 and test/mulls_reg.cpp:134-145 compile unchanged.
#ifndef MULLS_B200_DROPIN_CFILTER_HPP
#define MULLS_B200_DROPIN_CFILTER_HPP

#define CFilter CFilter_reference
#include_next "cfilter.hpp"
#undef CFilter

#include "common/cfilter_b200.hpp"

namespace lo {

template <typename PointT>
class CFilter : public CFilter_reference<PointT> {
    typedef typename pcl::PointCloud<PointT>::Ptr CloudPtr;

  public:
    // cfilter.hpp:2295-2318
    bool extract_semantic_pts(cloudblock_Ptr in_block, float vf_downsample_resolution, float gf_grid_resolution,
                              float gf_max_grid_height_diff, float gf_neighbor_height_diff, float gf_max_ground_height,
                              int &gf_down_rate_ground, int &gf_downsample_rate_nonground, float pca_neighbor_radius,
                              int pca_neighbor_k, float edge_thre, float planar_thre, float curvature_thre, float edge_thre_down,
                              float planar_thre_down, bool use_distance_adaptive_pca = false,
                              int distance_inverse_sampling_method = 0, float standard_distance = 15.0,
                              int estimate_ground_normal_method = 3, float normal_estimation_radius = 2.0,
                              bool use_adpative_parameters = false, bool apply_scanner_filter = false,
                              bool extract_curb_or_not = false, int extract_vertex_points_method = 2,
                              int gf_grid_pt_num_thre = 8, int gf_reliable_neighbor_grid_thre = 0,
                              int gf_down_down_rate_ground = 2, int pca_neighbor_k_min = 8, int pca_down_rate = 1,
                              float intensity_thre = FLT_MAX, float linear_vertical_sin_high_thre = 0.94,
                              float linear_vertical_sin_low_thre = 0.17, float planar_vertical_sin_high_thre = 0.98,
                              float planar_vertical_sin_low_thre = 0.34, bool sharpen_with_nms_on = true,
                              bool fixed_num_downsampling = false, int ground_down_fixed_num = 500,
                              int pillar_down_fixed_num = 200, int facade_down_fixed_num = 800, int beam_down_fixed_num = 200,
                              int roof_down_fixed_num = 200, int unground_down_fixed_num = 20000, float beam_height_max = FLT_MAX,
                              float roof_height_min = 0.0, float approx_scanner_height = 2.0, float underground_thre = -7.0,
                              float feature_pts_ratio_guess = 0.3, bool semantic_assisted = false,
                              bool apply_roi_filtering = false, float roi_min_y = 0.0, float roi_max_y = 0.0) {
        return b200::extract_semantic_pts<PointT>(
            in_block, vf_downsample_resolution, gf_grid_resolution, gf_max_grid_height_diff, gf_neighbor_height_diff,
            gf_max_ground_height, gf_down_rate_ground, gf_downsample_rate_nonground, pca_neighbor_radius, pca_neighbor_k, edge_thre,
            planar_thre, curvature_thre, edge_thre_down, planar_thre_down, use_distance_adaptive_pca,
            distance_inverse_sampling_method, standard_distance, estimate_ground_normal_method, normal_estimation_radius,
            use_adpative_parameters, apply_scanner_filter, extract_curb_or_not, extract_vertex_points_method, gf_grid_pt_num_thre,
            gf_reliable_neighbor_grid_thre, gf_down_down_rate_ground, pca_neighbor_k_min, pca_down_rate, intensity_thre,
            linear_vertical_sin_high_thre, linear_vertical_sin_low_thre, planar_vertical_sin_high_thre,
            planar_vertical_sin_low_thre, sharpen_with_nms_on, fixed_num_downsampling, ground_down_fixed_num,
            pillar_down_fixed_num, facade_down_fixed_num, beam_down_fixed_num, roof_down_fixed_num, unground_down_fixed_num,
            beam_height_max, roof_height_min, approx_scanner_height, underground_thre, feature_pts_ratio_guess, semantic_assisted,
            apply_roi_filtering, roi_min_y, roi_max_y);
    }
    // cfilter.hpp:83
    bool voxel_downsample(const CloudPtr &cloud_in, CloudPtr &cloud_out, float voxel_size) {
        return b200::voxel_downsample<PointT>(cloud_in, cloud_out, voxel_size);
    }
};

} // namespace lo
#endif
