// oracle/pcl_shim/pcl/point_types.h -- TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-in for <pcl/point_types.h> so that the reference ikd-Tree
// (/root/reference/eskf_lio/include/ikd-Tree/ikd_Tree.{h,cpp}, whose only
// non-std include is <pcl/point_types.h>, ikd_Tree.h:11) compiles in this
// container, where PCL and Eigen are not installed.  It provides the three
// point types the reference instantiates (ikd_Tree.cpp:1725-1727) with PCL's
// memory layout, and the Eigen::aligned_allocator name ikd_Tree.h:56 uses.
#pragma once
#include <memory>
#include <vector>

namespace pcl {
struct alignas(16) PointXYZ {
    float x, y, z, data3;
    PointXYZ() : x(0), y(0), z(0), data3(1.0f) {}
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), data3(1.0f) {}
};
struct alignas(16) PointXYZI {
    float x, y, z, data3;
    float intensity, pad0, pad1, pad2;
    PointXYZI() : x(0), y(0), z(0), data3(1.0f), intensity(0), pad0(0), pad1(0), pad2(0) {}
};
// 48 bytes: x y z 1 | normal_x normal_y normal_z 0 | intensity curvature pad pad
struct alignas(16) PointXYZINormal {
    float x, y, z, data3;
    float normal_x, normal_y, normal_z, data_n3;
    float intensity, curvature, pad0, pad1;
    PointXYZINormal()
        : x(0), y(0), z(0), data3(1.0f), normal_x(0), normal_y(0), normal_z(0), data_n3(0),
          intensity(0), curvature(0), pad0(0), pad1(0) {}
};
}  // namespace pcl

namespace Eigen {
template <class T>
using aligned_allocator = std::allocator<T>;
}
