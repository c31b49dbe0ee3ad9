// oracle/shim/pcl/filters/voxel_grid.h — TEST INFRASTRUCTURE ONLY: pcl::VoxelGrid<PointXYZI> with the interface the reference uses
// (setInputCloud / setLeafSize / filter; scanRegistration.cpp:401-405, laserMapping.cpp:129-130,543-549,793-799,905-906).  PCL is
// absent from this image: the arithmetic is the oracle's restatement of PCL 1.8.0 VoxelGrid::applyFilter (oracle/voxel_grid.hpp,
// [3P-mem], parity unpinned for that part); everything around the call is the reference's own code.
#pragma once
#include <pcl/filters/filter.h>
#include "voxel_grid.hpp"   // oracle/voxel_grid.hpp via -I<oracle>
namespace pcl {
template <typename PointT>
class VoxelGrid {
 public:
  typedef typename PointCloud<PointT>::Ptr PointCloudPtr;
  typedef typename PointCloud<PointT>::ConstPtr PointCloudConstPtr;
  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void filter(PointCloud<PointT>& output) {
    if (!input_) { output.width = output.height = 0; output.points.clear(); return; }
    output.header = input_->header;
    lvo_oracle::Cloud in(input_->points.size()), out;
    for (size_t i = 0; i < in.size(); ++i) in[i] = lvo_oracle::Pt{input_->points[i].x, input_->points[i].y, input_->points[i].z, input_->points[i].intensity};
    lvo_oracle::voxel_grid(in, leaf_[0], out);   // the reference only ever sets cubic leaves
    output.points.resize(out.size());
    for (size_t i = 0; i < out.size(); ++i) { PointT p; p.x = out[i].x; p.y = out[i].y; p.z = out[i].z; p.intensity = out[i].i; output.points[i] = p; }
    output.height = 1; output.is_dense = true; output.width = static_cast<uint32_t>(output.points.size());
  }
 private:
  PointCloudConstPtr input_;
  float leaf_[3] = {0, 0, 0};
};
}
