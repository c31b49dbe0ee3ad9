// oracle/shim/pcl_conversions/pcl_conversions.h — TEST INFRASTRUCTURE ONLY: pcl::fromROSMsg / pcl::toROSMsg between the shim
// sensor_msgs::PointCloud2 (x, y, z, intensity floats) and pcl::PointCloud<PointXYZ | PointXYZI>, as pcl_conversions does for those
// fields: a PointXYZ target drops the intensity (scanRegistration.cpp:132-133), header stamp / is_dense / width / height carried over.
#pragma once
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
namespace detail {
inline void put_intensity(PointXYZ&, float) {}
inline void put_intensity(PointXYZI& p, float v) { p.intensity = v; }
inline float get_intensity(const PointXYZ&) { return 0.f; }
inline float get_intensity(const PointXYZI& p) { return p.intensity; }
}
template <typename T>
void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<T>& cloud) {
  cloud.header.stamp_sec = msg.header.stamp.toSec();
  cloud.header.stamp = static_cast<uint64_t>(msg.header.stamp.toSec() * 1e6);
  cloud.header.frame_id = msg.header.frame_id;
  const size_t n = msg.xyzi.size() / 4;
  cloud.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    T p;
    p.x = msg.xyzi[4 * i]; p.y = msg.xyzi[4 * i + 1]; p.z = msg.xyzi[4 * i + 2];
    detail::put_intensity(p, msg.xyzi[4 * i + 3]);
    cloud.points[i] = p;
  }
  cloud.width = msg.width; cloud.height = msg.height; cloud.is_dense = msg.is_dense;
}
template <typename T>
void toROSMsg(const PointCloud<T>& cloud, sensor_msgs::PointCloud2& msg) {
  const size_t n = cloud.points.size();
  msg.xyzi.resize(4 * n);
  for (size_t i = 0; i < n; ++i) {
    msg.xyzi[4 * i] = cloud.points[i].x; msg.xyzi[4 * i + 1] = cloud.points[i].y; msg.xyzi[4 * i + 2] = cloud.points[i].z;
    msg.xyzi[4 * i + 3] = detail::get_intensity(cloud.points[i]);
  }
  if (cloud.width == 0 && cloud.height == 0) { msg.width = static_cast<unsigned>(n); msg.height = 1; }
  else { msg.width = cloud.width; msg.height = cloud.height; }
  msg.is_dense = cloud.is_dense;
  msg.header.stamp.fromSec(cloud.header.stamp_sec);
  msg.header.frame_id = cloud.header.frame_id;
}
}
