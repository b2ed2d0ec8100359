// Minimal stand-ins for the PCL / Eigen / boost / glog declarations and for the members of
// lo::cloudblock_t / lo::constraint_t (include/common/utility.hpp:233-553, :561-590 of the reference) that
// the drop-in shim touches. Used ONLY by tests/test_shim_compile.py to prove that the shim has the
// reference's signature and builds against the C-ABI without PCL installed.
#pragma once
#include <iostream>
#include <memory>
#include <vector>

namespace boost {
template <typename T>
using shared_ptr = std::shared_ptr<T>;
}
namespace Eigen {
template <int R, int C>
struct MatrixStub {
    double v[R * C];
    MatrixStub() {
        for (int i = 0; i < R * C; ++i) v[i] = 0.0;
    }
    double &operator()(int r, int c) { return v[r + R * c]; } // column-major like Eigen's default
    double operator()(int r, int c) const { return v[r + R * c]; }
    static MatrixStub Identity() {
        MatrixStub m;
        for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0;
        return m;
    }
    void setIdentity() { *this = Identity(); }
};
typedef MatrixStub<4, 4> Matrix4d;
} // namespace Eigen
typedef Eigen::MatrixStub<6, 6> Matrix6d;

namespace pcl {
struct PointXYZINormal { // 48 bytes, same field order as PCL's
    float x, y, z, pad0;
    float normal_x, normal_y, normal_z, pad1;
    float intensity, curvature, pad2, pad3;
};
template <typename P>
struct PointCloud {
    std::vector<P> points;
    typedef boost::shared_ptr<PointCloud<P>> Ptr;
};
} // namespace pcl
typedef pcl::PointXYZINormal Point_T;

struct LogSink {
    template <typename T>
    LogSink &operator<<(const T &t) {
        std::cerr << t;
        return *this;
    }
    ~LogSink() { std::cerr << std::endl; }
};
#define LOG(level) LogSink()

namespace lo {
struct bounds_t {
    double min_x = 0, min_y = 0, min_z = 0, max_x = 0, max_y = 0, max_z = 0;
};
struct centerpoint_t {
    double x = 0, y = 0, z = 0;
};
struct cloudblock_t {
    typedef pcl::PointCloud<Point_T>::Ptr pcTPtr;
    bounds_t local_bound, bound;
    centerpoint_t local_station;
    Eigen::Matrix4d pose_lo = Eigen::Matrix4d::Identity(), pose_gt = Eigen::Matrix4d::Identity();
    int feature_point_num = 0;
    pcTPtr pc_ground, pc_facade, pc_roof, pc_pillar, pc_beam, pc_vertex;
    pcTPtr pc_ground_down, pc_facade_down, pc_roof_down, pc_pillar_down, pc_beam_down, pc_unground, pc_raw, pc_down, pc_sketch;
    int down_feature_point_num = 0;
    cloudblock_t() {
        pcTPtr *all[] = {&pc_ground, &pc_facade, &pc_roof, &pc_pillar, &pc_beam, &pc_vertex, &pc_unground, &pc_raw, &pc_down, &pc_sketch,
                         &pc_ground_down, &pc_facade_down, &pc_roof_down, &pc_pillar_down, &pc_beam_down};
        for (auto p : all) *p = pcTPtr(new pcl::PointCloud<Point_T>());
    }
};
typedef boost::shared_ptr<cloudblock_t> cloudblock_Ptr;
struct constraint_t {
    cloudblock_Ptr block1, block2;
    Eigen::Matrix4d Trans1_2;
    Matrix6d information_matrix;
    float confidence = 0, sigma = 0;
    constraint_t() : block1(new cloudblock_t), block2(new cloudblock_t) {
        Trans1_2.setIdentity();
        information_matrix.setIdentity();
    }
};
} // namespace lo
