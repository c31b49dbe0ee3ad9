// oracle/shim/pcl/filters/filter.h — TEST INFRASTRUCTURE ONLY: pcl::removeNaNFromPointCloud (PCL 1.8.0 filters/impl/filter.hpp):
// a plain copy when the input is flagged dense, otherwise drops points with a non-finite coordinate.  Call site: scanRegistration.cpp:136.
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT>
void removeNaNFromPointCloud(const PointCloud<PointT>& cloud_in, PointCloud<PointT>& cloud_out, std::vector<int>& index) {
  if (&cloud_in != &cloud_out) { cloud_out.header = cloud_in.header; cloud_out.points.resize(cloud_in.points.size()); }
  index.resize(cloud_in.points.size());
  size_t j = 0;
  if (cloud_in.is_dense) {
    cloud_out = cloud_in;
    for (j = 0; j < cloud_out.points.size(); ++j) index[j] = static_cast<int>(j);
  } else {
    for (size_t i = 0; i < cloud_in.points.size(); ++i) {
      if (!pcl_isfinite(cloud_in.points[i].x) || !pcl_isfinite(cloud_in.points[i].y) || !pcl_isfinite(cloud_in.points[i].z)) continue;
      cloud_out.points[j] = cloud_in.points[i];
      index[j] = static_cast<int>(i);
      j++;
    }
    if (j != cloud_in.points.size()) { cloud_out.points.resize(j); index.resize(j); }
    cloud_out.height = 1;
    cloud_out.width = static_cast<uint32_t>(j);
    cloud_out.is_dense = true;
  }
}
}
