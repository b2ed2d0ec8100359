// The reference's call (test/mulls_reg.cpp:194-195) against the drop-in shim, with stand-in PCL/Eigen types.
#include <cstdio>
#include <string>

#include "utility.hpp"
#include "common/cregistration_b200.hpp"
#include "pgo/map_manager_b200.hpp"
#include "common/cfilter_b200.hpp"

using namespace lo;

int main() {
    constraint_t reg_con;
    int reg_max_iter_num = 10;
    float reg_corr_dis_thre = 3.0f, converge_tran = 0.001f, converge_rot_d = 0.01f;
    Eigen::Matrix4d init_mat = Eigen::Matrix4d::Identity();
    // argument list of test/mulls_reg.cpp:194-195
    int code = lo::b200::mm_lls_icp<Point_T>(reg_con, reg_max_iter_num, reg_corr_dis_thre, converge_tran, converge_rot_d,
                                             0.25 * reg_corr_dis_thre, 1.1, "111110", "1101", 1.0, 0.1, 0.1, 0.1, init_mat);
    // defaults-only call
    int code2 = lo::b200::mm_lls_icp<Point_T>(reg_con);
    bool ok4 = lo::b200::mm_lls_icp_4dof_global<Point_T>(reg_con, 45.0f); // the commented-out call of test/mulls_reg.cpp:197
    // the local-map calls of test/mulls_slam.cpp:438-442 and :477-482 on the device-resident map
    lo::b200::MapManagerB200 mmanager(1 << 12, 1 << 12);
    cloudblock_Ptr cblock_local_map(new cloudblock_t), cblock_target(new cloudblock_t);
    bool up1 = mmanager.update_local_map(cblock_local_map, cblock_target, 50.0f, 8000, 1000, 60.0f, false, "111110");
    bool up2 = mmanager.update_local_map(cblock_local_map, cblock_target, 50.0f, 8000, 1000, 60.0f, true, "111110", 15.0f,
                                         0.15f, 1.5f, 0.03f, true);
    reg_con.block1 = cblock_local_map;
    int code3 = mmanager.mm_lls_icp(reg_con, reg_max_iter_num, reg_corr_dis_thre, converge_tran, converge_rot_d);
    // CFilter::classify_nground_pts with the argument list of cfilter.hpp:2367-2378 (extract_semantic_pts)
    cloudblock_Ptr in_block(new cloudblock_t);
    bool cls = lo::b200::classify_nground_pts<Point_T>(in_block->pc_unground, in_block->pc_pillar, in_block->pc_beam, in_block->pc_facade,
                                                       in_block->pc_roof, in_block->pc_pillar_down, in_block->pc_beam_down,
                                                       in_block->pc_facade_down, in_block->pc_roof_down, in_block->pc_vertex, 1.0f, 50, 8, 1,
                                                       0.65f, 0.65f, 0.75f, 0.75f, 2, 0.12f, 1.5f, 0.94f, 0.17f, 0.98f, 0.34f, true, 200, 800,
                                                       200, 100, 20000, FLT_MAX, 0.0f, 0.3f, true, false);
    (void)cls;
    // the first two stages of extract_semantic_pts: voxel_downsample (cfilter.hpp:2346) and fast_ground_filter with the
    // argument list of :2355-2361
    in_block->pc_raw.reset(new pcl::PointCloud<Point_T>());
    in_block->pc_down.reset(new pcl::PointCloud<Point_T>());
    bool vox = lo::b200::voxel_downsample<Point_T>(in_block->pc_raw, in_block->pc_down, 0.05f);
    bool gf = lo::b200::fast_ground_filter<Point_T>(in_block->pc_down, in_block->pc_ground, in_block->pc_ground_down,
                                                    in_block->pc_unground, in_block->pc_vertex, 10, 3.0f, 0.3f, 1.5f, 5.0f, 15, 2, 3,
                                                    0, 3, 2.0f, 2, 15.0f, false, 300, false, FLT_MAX, false);
    (void)vox;
    (void)gf;
    // CFilter::extract_semantic_pts with the argument list of test/mulls_slam.cpp:363-377 (all 50 arguments)
    cloudblock_Ptr blk(new cloudblock_t);
    int ground_down_rate = 15, nonground_down_rate = 3;
    bool ext = lo::b200::extract_semantic_pts<Point_T>(blk, 0.05f, 3.0f, 0.3f, 1.5f, 5.0f, ground_down_rate, nonground_down_rate, 1.0f, 50,
                                                       0.65f, 0.65f, 0.12f, 0.75f, 0.75f, false, 2, 15.0f, 3, 2.0f, false, false, false, 2, 10,
                                                       0, 2, 8, 1, FLT_MAX, 0.94f, 0.17f, 0.98f, 0.34f, true, false, 300, 200, 800, 200, 100,
                                                       10000, FLT_MAX, 0.0f, 2.0f, -7.0f, 0.3f, false, false, 0.0f, 0.0f);
    (void)ext;
    std::printf("shim compiled and linked; codes %d %d %d | map %d %d %d\n", code, code2, (int)ok4, (int)up1, (int)up2, code3);
    return 0;
}
