// oracle/refnode/refnode.cpp — TEST INFRASTRUCTURE ONLY.
// Wraps ONE of the reference's node sources, compiled UNMODIFIED from where it lies under /root/reference
//     -DLVO_REF_SOURCE='"/root/reference/src/scanRegistration.cpp"'   -> oracle/_ref/libref_scan_registration.so
//     -DLVO_REF_SOURCE='"/root/reference/src/laserOdometry.cpp"'      -> oracle/_ref/libref_laser_odometry.so
//     -DLVO_REF_SOURCE='"/root/reference/src/laserMapping.cpp"'       -> oracle/_ref/libref_laser_mapping.so
// against the shim ROS / PCL / Ceres / glog headers of oracle/shim (their back ends: oracle/voxel_grid.hpp, knn.hpp, lm_core.hpp),
// and exports a small C harness API: start the node's main() on a thread, push input messages, wait for / take what it publishes.
// The only preprocessor intervention is `#define main` (the node's main becomes a callable) and a quiet `printf`.
// Everything the node computes — control flow, constants, float/double conversions, queue handling, the Eigen calls — is the
// reference's own text; this is what pins the hand restatement in oracle/*.hpp (tests/test_refnode_cpu.py).
#include <cstdio>
#include <cmath>
#include <cstring>
#include <iostream>
#include <mutex>
#include <queue>
#include <set>
#include <string>
#include <thread>
#include <vector>
#include <Eigen/Dense>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#include <sensor_msgs/Imu.h>
#include <nav_msgs/Odometry.h>
#include <nav_msgs/Path.h>
#include <geometry_msgs/PoseStamped.h>
#include <tf/transform_datatypes.h>
#include <tf/transform_broadcaster.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/kdtree/kdtree_flann.h>
#include <pcl_conversions/pcl_conversions.h>
#include <ceres/ceres.h>
#include <gflags/gflags.h>
#include <glog/logging.h>

namespace {
std::set<std::string>& kept_topics() { static std::set<std::string>* s = new std::set<std::string>(); return *s; }
// one record per ceres::Solve call: row 0 = (number of residual blocks, 0...), then the LM trace rows
void lm_trace_to_bus(int n_blocks, const std::vector<lvo_oracle::LmTraceRow>& tr) {
  auto v = std::make_shared<std::vector<lvo_oracle::LmTraceRow>>();
  lvo_oracle::LmTraceRow head; memset(&head, 0, sizeof(head)); head.x[0] = n_blocks;
  v->push_back(head);
  v->insert(v->end(), tr.begin(), tr.end());
  lvo_shim::publish("/lvo_shim/lm_trace", std::shared_ptr<const std::vector<lvo_oracle::LmTraceRow>>(v));
}
}

#ifdef LVO_REFNODE_LVO_ATAN
// Variant build of scanRegistration (libref_scan_registration_lvoatan.so): the node's atan / atan2 calls (scanRegistration.cpp:141-143,
// 166,208) go to the deterministic lvo_math.h routines that the restatement and the CUDA kernels share (DESIGN.md deviation 3: glibc and
// CUDA libm differ in the last ulp), so that the restatement can be compared with the reference's own text BIT FOR BIT.  Both float
// and double overloads are offered, as <cmath> / <math.h> do: which one a call takes is decided by the reference's expression types.
#include "../../lidar-visual-odometry_b200/csrc/lvo_math.h"
namespace lvo_refatan {
inline float atan_(float v) { return lvo_atanf(v); }
inline double atan_(double v) { return ::atan(v); }
inline float atan2_(float y, float x) { return lvo_atan2f(y, x); }
inline double atan2_(double y, double x) { return ::atan2(y, x); }
}
namespace std { using lvo_refatan::atan_; using lvo_refatan::atan2_; }
using lvo_refatan::atan_;
using lvo_refatan::atan2_;
#define atan atan_
#define atan2 atan2_
#endif
#define printf(...) lvo_shim::quiet_printf(__VA_ARGS__)
#define main lvo_refnode_main
#include LVO_REF_SOURCE
#undef main
#undef printf
#ifdef LVO_REFNODE_LVO_ATAN
#undef atan
#undef atan2
#endif

#define LVO_API __attribute__((visibility("default")))
extern "C" {

// params: "name=value;name=value" as the launch file would set them; keep: "topic;topic" = the outputs the harness will take
// (everything else the node publishes is counted and dropped).
LVO_API int refnode_start(const char* params, const char* keep) {
  lvo_shim::Bus& b = lvo_shim::bus();
  {
    std::lock_guard<std::mutex> lk(b.mu);
    b.verbose = getenv("LVO_REFNODE_VERBOSE") != nullptr;
    std::string s = params ? params : "";
    size_t p = 0;
    while (p < s.size()) {
      size_t e = s.find(';', p); if (e == std::string::npos) e = s.size();
      const std::string kv = s.substr(p, e - p);
      const size_t q = kv.find('=');
      if (q != std::string::npos) b.params[kv.substr(0, q)] = kv.substr(q + 1);
      p = e + 1;
    }
    s = keep ? keep : ""; p = 0;
    while (p < s.size()) {
      size_t e = s.find(';', p); if (e == std::string::npos) e = s.size();
      if (e > p) kept_topics().insert(s.substr(p, e - p));
      p = e + 1;
    }
    b.keep = &kept_topics();
  }
  lvo_oracle::lm_trace_hook() = lm_trace_to_bus;
  std::thread([] { static char arg0[] = "refnode"; char* argv[] = {arg0, nullptr}; lvo_refnode_main(1, argv); }).detach();
  return 0;
}
LVO_API void refnode_push_cloud(const char* topic, double stamp, const float* xyzi, long n, int is_dense) {
  auto m = std::make_shared<sensor_msgs::PointCloud2>();
  m->header.stamp.fromSec(stamp);
  m->xyzi.assign(xyzi, xyzi + 4 * n);
  m->width = (unsigned)n; m->height = 1; m->is_dense = is_dense != 0;
  lvo_shim::post(topic, m);
}
LVO_API void refnode_push_odom(const char* topic, double stamp, const double* qt7) {
  auto m = std::make_shared<nav_msgs::Odometry>();
  m->header.stamp.fromSec(stamp);
  m->pose.pose.orientation.x = qt7[0]; m->pose.pose.orientation.y = qt7[1]; m->pose.pose.orientation.z = qt7[2]; m->pose.pose.orientation.w = qt7[3];
  m->pose.pose.position.x = qt7[4]; m->pose.pose.position.y = qt7[5]; m->pose.pose.position.z = qt7[6];
  lvo_shim::post(topic, m);
}
// blocks until `count` messages have been published on topic (or timeout); returns the number published so far
LVO_API long refnode_wait(const char* topic, long count, int timeout_ms) {
  lvo_shim::Bus& b = lvo_shim::bus();
  std::unique_lock<std::mutex> lk(b.mu);
  b.cv_out.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return b.published[topic] >= count; });
  return b.published[topic];
}
static std::shared_ptr<const void> take(const char* topic, bool pop) {
  lvo_shim::Bus& b = lvo_shim::bus();
  std::lock_guard<std::mutex> lk(b.mu);
  auto it = b.outbound.find(topic);
  if (it == b.outbound.end() || it->second.empty()) return nullptr;
  std::shared_ptr<const void> m = it->second.front();
  if (pop) it->second.pop_front();
  return m;
}
// out == NULL: size of the oldest kept message (or -1); otherwise copies it (4 floats per point) and pops it
LVO_API long refnode_take_cloud(const char* topic, float* out, long cap, double* stamp) {
  auto m = std::static_pointer_cast<const sensor_msgs::PointCloud2>(take(topic, out != nullptr));
  if (!m) return -1;
  const long n = (long)(m->xyzi.size() / 4);
  if (stamp) *stamp = m->header.stamp.toSec();
  if (out) { if (n > cap) return -2; if (n) memcpy(out, m->xyzi.data(), sizeof(float) * 4 * n); }
  return n;
}
LVO_API int refnode_take_odom(const char* topic, double* qt7, double* stamp) {
  auto m = std::static_pointer_cast<const nav_msgs::Odometry>(take(topic, true));
  if (!m) return -1;
  qt7[0] = m->pose.pose.orientation.x; qt7[1] = m->pose.pose.orientation.y; qt7[2] = m->pose.pose.orientation.z; qt7[3] = m->pose.pose.orientation.w;
  qt7[4] = m->pose.pose.position.x; qt7[5] = m->pose.pose.position.y; qt7[6] = m->pose.pose.position.z;
  if (stamp) *stamp = m->header.stamp.toSec();
  return 0;
}
// LM traces of the node's ceres::Solve calls, oldest first: rows x 10 doubles (x7, cost, radius, flags); returns rows or -1
LVO_API long refnode_take_lm_trace(double* out, long cap_rows) {
  auto m = std::static_pointer_cast<const std::vector<lvo_oracle::LmTraceRow>>(take("/lvo_shim/lm_trace", out != nullptr));
  if (!m) return -1;
  const long n = (long)m->size();
  if (out) {
    if (n > cap_rows) return -2;
    for (long i = 0; i < n; ++i) { for (int k = 0; k < 7; ++k) out[i * 10 + k] = (*m)[i].x[k]; out[i * 10 + 7] = (*m)[i].cost; out[i * 10 + 8] = (*m)[i].radius; out[i * 10 + 9] = (*m)[i].flags; }
  }
  return n;
}
LVO_API // Read-only views of the node's own globals (the harness is part of the same translation unit), for parity probes:
//   scanRegistration (LVO_REFNODE_KIND 1): what = 0 cloudCurvature (float), 1 cloudSortInd, 2 cloudNeighborPicked, 3 cloudLabel (int), n values
//   laserMapping     (LVO_REFNODE_KIND 3): what = 0 / 1 corner / surf map, cube-major (out = points xyzi, aux = cube index of each point);
//                                          2 / 3 laserCloudCornerFromMap / SurfFromMap; 4: aux[0..2] = laserCloudCen{Width,Height,Depth}
// Call only while the node is idle (after refnode_wait returned for the frame's last publication).  Returns the count, -1 if unknown.
LVO_API long refnode_peek(int what, void* out, int* aux, long cap) {
#if LVO_REFNODE_KIND == 1
  if (what < 0 || what > 3 || cap > 400000) return -1;
  if (what == 0) memcpy(out, cloudCurvature, sizeof(float) * cap);
  else memcpy(out, what == 1 ? cloudSortInd : (what == 2 ? cloudNeighborPicked : cloudLabel), sizeof(int) * cap);
  return cap;
#elif LVO_REFNODE_KIND == 3
  float* o = static_cast<float*>(out);
  long n = 0;
  auto put = [&](const pcl::PointCloud<PointType>& c, int cube) {
    for (const PointType& p : c.points) {
      if (o && n < cap) { o[4 * n] = p.x; o[4 * n + 1] = p.y; o[4 * n + 2] = p.z; o[4 * n + 3] = p.intensity; if (aux) aux[n] = cube; }
      ++n;
    }
  };
  if (what == 0 || what == 1) { for (int i = 0; i < laserCloudNum; ++i) put(what == 0 ? *laserCloudCornerArray[i] : *laserCloudSurfArray[i], i); return n; }
  if (what == 2) { put(*laserCloudCornerFromMap, 0); return n; }
  if (what == 3) { put(*laserCloudSurfFromMap, 0); return n; }
  if (what == 4 && aux) { aux[0] = laserCloudCenWidth; aux[1] = laserCloudCenHeight; aux[2] = laserCloudCenDepth; return 3; }
  return -1;
#else
  (void)what; (void)out; (void)aux; (void)cap;
  return -1;
#endif
}
LVO_API void refnode_stop() {
  lvo_shim::Bus& b = lvo_shim::bus();
  { std::lock_guard<std::mutex> lk(b.mu); b.shutdown = true; }
  b.cv_in.notify_all();
}

}  // extern "C"
