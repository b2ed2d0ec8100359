// DROP-IN replacement of lo::MapManager (reference: include/pgo/map_manager.h:19-55, src/map_manager.cpp:17-314).
//
// Put this directory BEFORE the reference's include/pgo on the include path and leave src/map_manager.cpp out of the
// build (this header is the whole class). test/mulls_slam.cpp:270, :438-446 compile unchanged: same class name, same
// public members with the same names, argument orders, types and defaults.
//
// update_local_map keeps the local map RESIDENT IN HBM (mulls_map_*, include/mulls_b200/abi.h): per frame only the new
// scan's down-sampled feature clouds cross PCIe; the map's host clouds (local_map->pc_*) are refreshed from the device
// after every update, so every other reader of cblock_local_map (viewer, submap copies at test/mulls_slam.cpp:454) keeps
// working. The cloudblock is also REGISTERED as resident: the drop-in lo::CRegistration::mm_lls_icp recognises
// registration_cons.block1 == that cloudblock (test/mulls_slam.cpp:669-685) and registers against the copy in HBM
// instead of uploading the map again. The kd-trees the reference's mm_lls_icp leaves in local_map->tree_* are not
// needed: the dynamic-object removal (src/map_manager.cpp:149-258) runs on the device on the sorted target slices the
// preceding registration left there.
#ifndef _INCLUDE_MAP_MANAGER_H /* the reference's guard: its own header becomes a no-op after this one */
#define _INCLUDE_MAP_MANAGER_H

#include <cfloat>
#include <memory>
#include <string>

#include "pgo/map_manager_b200.hpp"

namespace lo {

class MapManager {
    typedef pcl::PointCloud<Point_T>::Ptr CloudPtr;

  public:
    // include/pgo/map_manager.h:22-32
    bool update_local_map(cloudblock_Ptr local_map, cloudblock_Ptr last_target_cblock, float local_map_radius = 80,
                          int max_num_pts = 20000, int kept_vertex_num = 800, float last_frame_reliable_radius = 60,
                          bool map_based_dynamic_removal_on = false, std::string used_feature_type = "111110",
                          float dynamic_removal_center_radius = 30.0, float dynamic_dist_thre_min = 0.3,
                          float dynamic_dist_thre_max = 3.0, float near_dist_thre = 0.03, bool recalculate_feature_on = false) {
        b200::MapManagerB200 &m = b200::resident_map_for(local_map.get());
        return m.update_local_map(local_map, last_target_cblock, local_map_radius, max_num_pts, kept_vertex_num,
                                  last_frame_reliable_radius, map_based_dynamic_removal_on, used_feature_type,
                                  dynamic_removal_center_radius, dynamic_dist_thre_min, dynamic_dist_thre_max, near_dist_thre,
                                  recalculate_feature_on);
    }

    // :34-35 — part of update_local_map here (it needs the device copy of the map and of the last registration's
    // target slices); a stand-alone call has nothing to work on
    bool map_based_dynamic_close_removal(cloudblock_Ptr, cloudblock_Ptr, std::string, float, float, float, float) {
        LOG(WARNING) << "mulls_b200: map_based_dynamic_close_removal runs inside update_local_map (map_based_dynamic_removal_on)";
        return false;
    }
    // :39-40, :42-44 — helpers of the two members above in the reference; they take PCL kd-trees, which this path
    // never builds (mulls_nn_query answers the same nearestKSearch(p, 1) on the device)
    template <typename TreePtr>
    bool map_scan_feature_pts_distance_removal(CloudPtr, const TreePtr, float, float = FLT_MAX, float = FLT_MAX, float = 0.0) {
        LOG(WARNING) << "mulls_b200: map_scan_feature_pts_distance_removal is part of update_local_map on the device";
        return false;
    }
    template <typename TreePtr>
    bool update_cloud_vectors(CloudPtr, const TreePtr, float = 1.5, int = 20, int = 8, float = 0.5, float = 0.5, float = 0.0) {
        LOG(WARNING) << "mulls_b200: update_cloud_vectors is part of update_local_map on the device (recalculate_feature_on)";
        return false;
    }

    // :46-47, src/map_manager.cpp:296-314: a new submap starts when the accumulated motion or frame count passes a bound
    bool judge_new_submap(float &accu_tran, float &accu_rot, int &accu_frame, float max_accu_tran = 30.0,
                          float max_accu_rot = 90.0, int max_accu_frame = 150) {
        if (!(accu_tran > max_accu_tran || accu_rot > max_accu_rot || accu_frame > max_accu_frame)) return false;
        accu_tran = 0.0;
        accu_rot = 0.0;
        accu_frame = 0;
        return true;
    }
};

} // namespace lo
#endif //_INCLUDE_MAP_MANAGER_H
