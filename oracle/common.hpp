// oracle/common.hpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement ("oracle") of the reference hot path; it is the parity checker for the CUDA library and the
// CPU baseline that bench.py times.  Nothing under oracle/ is linked into, imported by or executed from the
// product (liblvo.so); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may use it.
//
// PARITY STATUS: the reference ships no tests, fixtures or golden vectors for this path, and its
// PCL/FLANN/Ceres dependencies are absent, so the oracle cannot be pinned against reference outputs: *parity
// unpinned* for the PCL VoxelGrid / FLANN kNN / Ceres LM restatements.  What IS pinned: the residual functors —
// oracle/_ref compiles the reference's own src/lidarFactor.hpp (through oracle/shim) plus the vendored Eigen
// 3.3.7 and tests/test_oracle_ref.py checks this restatement against it.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>

#include "../lidar-visual-odometry_b200/csrc/lvo_math.h"  // shared deterministic atan/atan2 (SURVEY §7 hard part 2)

namespace lvo_oracle {

// pcl::PointXYZI payload (include/aloam_velodyne/common.h:43), packed.
struct Pt { float x, y, z, i; };
typedef std::vector<Pt> Cloud;

struct Quat { double x, y, z, w; };  // storage order of Eigen::Quaterniond::coeffs() and of para_q
struct Vec3 { double x, y, z; };

inline Vec3 cross(const Vec3& a, const Vec3& b) {
  return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen 3.3.7 QuaternionBase::_transformVector (Eigen/src/Geometry/Quaternion.h:466-477):
//   uv = vec x v; uv += uv; return v + w*uv + vec x uv
inline Vec3 rotate(const Quat& q, const Vec3& v) {
  Vec3 qv{q.x, q.y, q.z};
  Vec3 uv = cross(qv, v);
  uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
  Vec3 c2 = cross(qv, uv);
  return Vec3{(v.x + q.w * uv.x) + c2.x, (v.y + q.w * uv.y) + c2.y, (v.z + q.w * uv.z) + c2.z};
}
// Eigen generic quat_product (Quaternion.h:436-444)
inline Quat qmul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
// Eigen QuaternionBase::inverse (Quaternion.h:679-693): conjugate / squaredNorm
inline Quat qinv(const Quat& q) {
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  return Quat{-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
}

}  // namespace lvo_oracle
