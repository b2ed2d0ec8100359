// TEST STAND-IN for the reference's include/common/cfilter.hpp (see ref/cregistration.hpp).
#ifndef STUB_REFERENCE_CFILTER_HPP
#define STUB_REFERENCE_CFILTER_HPP
#include <cfloat>

#include "utility.hpp"

namespace lo {
template <typename PointT>
class CFilter {
  public:
    bool extract_semantic_pts(cloudblock_Ptr, float, float, float, float, float, int &, int &, float, int, float, float, float, float,
                              float, bool = false, int = 0, float = 15.0, int = 3, float = 2.0, bool = false, bool = false,
                              bool = false, int = 2, int = 8, int = 0, int = 2, int = 8, int = 1, float = FLT_MAX, float = 0.94,
                              float = 0.17, float = 0.98, float = 0.34, bool = true, bool = false, int = 500, int = 200, int = 800,
                              int = 200, int = 200, int = 20000, float = FLT_MAX, float = 0.0, float = 2.0, float = -7.0,
                              float = 0.3, bool = false, bool = false, float = 0.0, float = 0.0) {
        reference_body_ran = true;
        return false;
    }
    bool voxel_downsample(const typename pcl::PointCloud<PointT>::Ptr &, typename pcl::PointCloud<PointT>::Ptr &, float) {
        reference_body_ran = true;
        return false;
    }
    bool dist_filter(const typename pcl::PointCloud<PointT>::Ptr &, double, double) { return true; } // inherited, untouched
    bool reference_body_ran = false;
};
} // namespace lo
#endif
