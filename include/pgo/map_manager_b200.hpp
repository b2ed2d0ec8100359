// Drop-in shim: lo::MapManager::update_local_map (include/pgo/map_manager.h:22-32, src/map_manager.cpp:17-145) and the
// scan-to-map registration over the mulls_b200 C-ABI, with the local map resident in HBM.
//
// A MULLS maintainer replaces in test/mulls_slam.cpp
//     mmanager.update_local_map(cblock_local_map, cblock_target, ...)            (:438-442)
//     creg.mm_lls_icp(current_registration_edge, ...)   // block1 = cblock_local_map  (:477-482)
// by the same calls on one lo::b200::MapManagerB200 object (identical argument lists). Per frame only the new scan's
// down-sampled feature clouds cross PCIe; the map's host clouds (local_map->pc_*) are refreshed from the device after
// every update so that the other readers of cblock_local_map (viewer, submap bookkeeping) keep working.
//
// Contract differences: last_target_cblock is left untouched (the reference leaves its *_down
// clouds in the old map frame); the budgeted down-sampling is a reproducible uniform sample.
#ifndef MULLS_B200_MAP_MANAGER_SHIM_HPP
#define MULLS_B200_MAP_MANAGER_SHIM_HPP

#include <cfloat>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "common/cregistration_b200.hpp"
#include "mulls_b200/abi.h"

namespace lo {
namespace b200 {

class MapManagerB200 {
  public:
    explicit MapManagerB200(size_t max_pts_per_class = 1 << 18, size_t max_src_pts = 1 << 20, uint32_t random_seed = 0)
        : seed_(random_seed) {
        ctx_ = mulls_create(0, 1, max_src_pts, 6 * max_pts_per_class);
        map_ = ctx_ ? mulls_map_create(ctx_, max_pts_per_class) : nullptr;
        if (!map_) LOG(ERROR) << "mulls_b200: " << mulls_last_error(ctx_);
    }
    ~MapManagerB200() {
        if (map_) mulls_map_destroy(map_);
        if (ctx_) mulls_destroy(ctx_);
    }
    MapManagerB200(const MapManagerB200 &) = delete;
    MapManagerB200 &operator=(const MapManagerB200 &) = delete;

    // MapManager::update_local_map — same names, order, types and defaults as include/pgo/map_manager.h:22-32
    bool update_local_map(cloudblock_Ptr local_map, cloudblock_Ptr last_target_cblock, float local_map_radius = 80,
                          int max_num_pts = 20000, int kept_vertex_num = 800, float last_frame_reliable_radius = 60,
                          bool map_based_dynamic_removal_on = false, std::string used_feature_type = "111110",
                          float dynamic_removal_center_radius = 30.0, float dynamic_dist_thre_min = 0.3,
                          float dynamic_dist_thre_max = 3.0, float near_dist_thre = 0.03,
                          bool recalculate_feature_on = false) {
        if (!map_) return false;
        typedef Point_T P;
        cloudblock_t &s = *last_target_cblock;
        const mulls_cloud_view scan[MULLS_NUM_CLASSES] = {view_of<P>(s.pc_ground_down), view_of<P>(s.pc_pillar_down),
                                                          view_of<P>(s.pc_facade_down), view_of<P>(s.pc_beam_down),
                                                          view_of<P>(s.pc_roof_down),   view_of<P>(s.pc_vertex)};
        mulls_map_params p;
        mulls_map_default_params(&p);
        p.local_map_radius = local_map_radius;
        p.max_num_pts = max_num_pts;
        p.kept_vertex_num = kept_vertex_num;
        p.last_frame_reliable_radius = last_frame_reliable_radius;
        p.map_based_dynamic_removal_on = map_based_dynamic_removal_on;
        std::strncpy(p.used_feature_type, used_feature_type.c_str(), 7);
        p.dynamic_removal_center_radius = dynamic_removal_center_radius;
        p.dynamic_dist_thre_min = dynamic_dist_thre_min;
        p.dynamic_dist_thre_max = dynamic_dist_thre_max;
        p.near_dist_thre = near_dist_thre;
        p.recalculate_feature_on = recalculate_feature_on;
        p.random_seed = seed_++;
        double pose[16];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) pose[4 * r + c] = s.pose_lo(r, c);
        mulls_map_info info;
        if (mulls_map_update(map_, scan, pose, &p, &info) != MULLS_OK) {
            LOG(ERROR) << "mulls_b200: " << mulls_last_error(ctx_);
            return false;
        }
        // mirror what the reference leaves in local_map (:58-59, :69-71, :88-93, :129-131)
        cloudblock_t &m = *local_map;
        cloudblock_t::pcTPtr *cls[MULLS_NUM_CLASSES] = {&m.pc_ground, &m.pc_pillar, &m.pc_facade, &m.pc_beam, &m.pc_roof, &m.pc_vertex};
        for (int c = 0; c < MULLS_NUM_CLASSES; ++c) {
            (*cls[c])->points.resize(info.n[c]);
            size_t n = 0;
            mulls_map_download(map_, c, info.n[c] ? reinterpret_cast<float *>((*cls[c])->points.data()) : nullptr, info.n[c], &n);
        }
        m.pose_lo = s.pose_lo;
        m.pose_gt = s.pose_gt;
        m.feature_point_num = info.feature_point_num;
        m.local_bound.min_x = info.local_bound[0], m.local_bound.min_y = info.local_bound[1], m.local_bound.min_z = info.local_bound[2];
        m.local_bound.max_x = info.local_bound[3], m.local_bound.max_y = info.local_bound[4], m.local_bound.max_z = info.local_bound[5];
        m.bound.min_x = info.bound[0], m.bound.min_y = info.bound[1], m.bound.min_z = info.bound[2];
        m.bound.max_x = info.bound[3], m.bound.max_y = info.bound[4], m.bound.max_z = info.bound[5];
        for (int c = 0; c < MULLS_NUM_CLASSES; ++c) n_[c] = info.n[c];
        synced_ = true;
        return true;
    }

    // CRegistration::mm_lls_icp with registration_cons.block1 = the local map kept by this object: same argument list
    // as cregistration.hpp:1114-1123; block1's host clouds are not read, the target comes from HBM.
    int mm_lls_icp(constraint_t &registration_cons, int max_iter_num = 20, float dis_thre_unit = 1.5,
                   float converge_translation = 0.002, float converge_rotation_d = 0.01, float dis_thre_min = 0.4,
                   float dis_thre_update_rate = 1.1, std::string used_feature_type = "111110",
                   std::string weight_strategy = "1101", float z_xy_balanced_ratio = 1.0, float pt2pt_residual_window = 0.1,
                   float pt2pl_residual_window = 0.1, float pt2li_residual_window = 0.1,
                   Eigen::Matrix4d initial_guess = Eigen::Matrix4d::Identity(), bool apply_intersection_filter = true,
                   bool apply_motion_undistortion_while_registration = false, bool normal_shooting_on = false,
                   float normal_bearing = 45.0, bool use_more_points = false, bool keep_less_source_points = false,
                   float sigma_thre = 0.5, float min_neccessary_corr_ratio = 0.03, float max_bearable_rotation_d = 45.0) {
        if (!map_) return -4;
        typedef Point_T P;
        cloudblock_t &b2 = *registration_cons.block2;
        const bool down = !use_more_points || apply_motion_undistortion_while_registration;
        const mulls_cloud_view src[MULLS_NUM_CLASSES] = {
            view_of<P>(down ? b2.pc_ground_down : b2.pc_ground), view_of<P>(down ? b2.pc_pillar_down : b2.pc_pillar),
            view_of<P>(down ? b2.pc_facade_down : b2.pc_facade), view_of<P>(down ? b2.pc_beam_down : b2.pc_beam),
            view_of<P>(down ? b2.pc_roof_down : b2.pc_roof),     view_of<P>(b2.pc_vertex)};
        mulls_icp_params p;
        mulls_icp_default_params(&p);
        p.max_iter_num = max_iter_num;
        p.dis_thre_unit = dis_thre_unit;
        p.converge_translation = converge_translation;
        p.converge_rotation_d = converge_rotation_d;
        p.dis_thre_min = dis_thre_min;
        p.dis_thre_update_rate = dis_thre_update_rate;
        std::strncpy(p.used_feature_type, used_feature_type.c_str(), 7);
        std::strncpy(p.weight_strategy, weight_strategy.c_str(), 7);
        p.z_xy_balanced_ratio = z_xy_balanced_ratio;
        p.pt2pt_residual_window = pt2pt_residual_window;
        p.pt2pl_residual_window = pt2pl_residual_window;
        p.pt2li_residual_window = pt2li_residual_window;
        p.apply_intersection_filter = apply_intersection_filter;
        p.apply_motion_undistortion_while_registration = apply_motion_undistortion_while_registration;
        p.normal_shooting_on = normal_shooting_on;
        p.normal_bearing = normal_bearing;
        p.use_more_points = use_more_points;
        p.keep_less_source_points = keep_less_source_points;
        p.sigma_thre = sigma_thre;
        p.min_neccessary_corr_ratio = min_neccessary_corr_ratio;
        p.max_bearable_rotation_d = max_bearable_rotation_d;
        p.random_seed = seed_;
        double init[16];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) init[4 * r + c] = initial_guess(r, c);
        mulls_icp_result out;
        if (mulls_icp_run_to_map(ctx_, map_, src, &p, init, &out, nullptr) != MULLS_OK) {
            LOG(ERROR) << "mulls_b200: " << mulls_last_error(ctx_); // same failure convention as b200::mm_lls_icp
            registration_cons.Trans1_2 = initial_guess;
            registration_cons.information_matrix.setIdentity();
            registration_cons.sigma = FLT_MAX;
            registration_cons.confidence = 0.0f;
            return -4;
        }
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) registration_cons.Trans1_2(r, c) = out.T[4 * r + c];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) registration_cons.information_matrix(r, c) = out.info[6 * r + c];
        registration_cons.sigma = out.sigma;
        registration_cons.confidence = out.confidence;
        return out.code;
    }

    // does `block` still hold exactly the clouds this object downloaded after its last update? (a caller that edits the
    // map's host clouds itself falls back to the host path of mm_lls_icp)
    bool mirrors(const cloudblock_t &block) const {
        if (!map_ || !synced_) return false;
        const cloudblock_t::pcTPtr *cls[MULLS_NUM_CLASSES] = {&block.pc_ground, &block.pc_pillar, &block.pc_facade,
                                                              &block.pc_beam,   &block.pc_roof,   &block.pc_vertex};
        for (int c = 0; c < MULLS_NUM_CLASSES; ++c)
            if ((*cls[c])->points.size() != n_[c]) return false;
        return true;
    }

  private:
    mulls_ctx *ctx_ = nullptr;
    mulls_map *map_ = nullptr;
    uint32_t seed_ = 0;
    bool synced_ = false;
    size_t n_[MULLS_NUM_CLASSES] = {0, 0, 0, 0, 0, 0};
};

// The cloudblocks whose clouds live in HBM (lo::MapManager::update_local_map of dropin/map_manager.h registers its
// local_map argument here; lo::CRegistration::mm_lls_icp of dropin/cregistration.hpp looks registration_cons.block1 up).
// One entry per cloudblock address and host thread.
inline MapManagerB200 *&resident_slot(const cloudblock_t *block) {
    static thread_local std::vector<std::pair<const cloudblock_t *, MapManagerB200 *>> table;
    for (auto &e : table)
        if (e.first == block) return e.second;
    table.push_back(std::make_pair(block, (MapManagerB200 *)nullptr));
    return table.back().second;
}
inline MapManagerB200 &resident_map_for(const cloudblock_t *block) {
    MapManagerB200 *&slot = resident_slot(block);
    if (!slot) slot = new MapManagerB200(); // lives as long as the thread's map (the SLAM loop's only local map)
    return *slot;
}
inline MapManagerB200 *resident_map_if_current(const cloudblock_t *block) {
    MapManagerB200 *m = resident_slot(block);
    return (m && m->mirrors(*block)) ? m : nullptr;
}

} // namespace b200
} // namespace lo
#endif
