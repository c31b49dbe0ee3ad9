// shim (test infrastructure): sensor_msgs/PointCloud2 reduced to what pcl::fromROSMsg / toROSMsg move for PointXYZ / PointXYZI:
// n records of (x, y, z, intensity) floats + is_dense.
#pragma once
#include <std_msgs/Header.h>
#include <memory>
#include <vector>
namespace sensor_msgs {
struct PointCloud2 {
  std_msgs::Header header;
  std::vector<float> xyzi;   // 4 floats per point
  unsigned width = 0, height = 1;
  bool is_dense = true;
  typedef std::shared_ptr<PointCloud2> Ptr;
  typedef std::shared_ptr<const PointCloud2> ConstPtr;
};
typedef std::shared_ptr<PointCloud2> PointCloud2Ptr;
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
}
