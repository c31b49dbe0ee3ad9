// oracle/shim/pcl/point_types.h — TEST INFRASTRUCTURE ONLY.
// pcl::PointXYZ / pcl::PointXYZI with PCL 1.8's memory layout (16 / 32 bytes, 16-byte aligned, data[3] = 1.0f) so that the
// reference's node sources compile unmodified (PCL is absent from this image).
#pragma once
#include <Eigen/Core>
#include <cmath>
namespace pcl {
struct alignas(16) PointXYZ {
  union { float data[4]; struct { float x, y, z; }; };
  PointXYZ() { x = y = z = 0.0f; data[3] = 1.0f; }
  PointXYZ(float a, float b, float c) { x = a; y = b; z = c; data[3] = 1.0f; }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};
struct alignas(16) PointXYZI {
  union { float data[4]; struct { float x, y, z; }; };
  union { struct { float intensity; }; float data_c[4]; };
  PointXYZI() { x = y = z = 0.0f; data[3] = 1.0f; intensity = 0.0f; data_c[1] = data_c[2] = data_c[3] = 0.0f; }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZI) == 32, "PCL point layout");
}
#define pcl_isfinite(x) std::isfinite(x)
