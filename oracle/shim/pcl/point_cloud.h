// oracle/shim/pcl/point_cloud.h — TEST INFRASTRUCTURE ONLY: pcl::PointCloud<PointT> as far as the reference's lidar nodes use it
// (points / width / height / is_dense / header, push_back, clear, +=, Ptr), following PCL 1.8.0 common/include/pcl/point_cloud.h.
#pragma once
#include <pcl/point_types.h>
#include <Eigen/StdVector>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
namespace pcl {
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; double stamp_sec = 0; std::string frame_id; };
template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  typedef std::vector<PointT, Eigen::aligned_allocator<PointT>> VectorType;
  typedef typename VectorType::iterator iterator;
  typedef typename VectorType::const_iterator const_iterator;
  PCLHeader header;
  VectorType points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  PointCloud() {}
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
  const PointT& operator[](size_t i) const { return points[i]; }
  PointT& operator[](size_t i) { return points[i]; }
  void push_back(const PointT& p) { points.push_back(p); width = static_cast<uint32_t>(points.size()); height = 1; }
  void clear() { points.clear(); width = 0; height = 0; }
  void resize(size_t n) { points.resize(n); if (width * height != n) { width = static_cast<uint32_t>(n); height = 1; } }
  PointCloud& operator+=(const PointCloud& rhs) {
    if (rhs.header.stamp_sec > header.stamp_sec) { header.stamp = rhs.header.stamp; header.stamp_sec = rhs.header.stamp_sec; }
    size_t nr = points.size();
    points.resize(nr + rhs.points.size());
    for (size_t i = nr; i < points.size(); ++i) points[i] = rhs.points[i - nr];
    width = static_cast<uint32_t>(points.size()); height = 1;
    is_dense = is_dense && rhs.is_dense;
    return *this;
  }
  PointCloud operator+(const PointCloud& rhs) const { PointCloud r = *this; r += rhs; return r; }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};
}
