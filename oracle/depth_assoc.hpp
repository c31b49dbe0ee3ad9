// oracle/depth_assoc.hpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).
// Restates the camera-lidar depth association of the reference (BASELINE config 5):
//   depth cloud      src/vloam/CamLidarProcess.cpp:249-253 (pcl::transformPointCloud by the 3x4 extrinsic, float) and
//                    src/vloam/Frame.cpp:328-336: every point with z > 0 becomes (10 x / z, 10 y / z, 10, I = z)
//   association      src/vloam/Frontend.cpp:233-301: 3-NN of (10 u, 10 v, 10) in the depth cloud (pcl::KdTreeFLANN, restated
//                    as exact search with (distance, index) order), gate d0^2 < 0.5, planar interpolation through the three
//                    points in double, the non-finite / max-min > 2 / +-0.2 clamps.
// pcl::transformPointCloud's float summation order is not visible here (PCL absent): fixed to ((m0 x + m1 y) + m2 z) + m3.
#pragma once
#include "common.hpp"
#include "knn.hpp"

namespace lvo_oracle {

// extr: 3x4 row-major lidar -> camera.  Returns the depth cloud; src_index (optional) gets the sweep index of each point.
inline void depth_cloud(const Pt* sweep, size_t n, const float* extr, Cloud& out, std::vector<int>* src_index) {
  out.clear();
  if (src_index) src_index->clear();
  for (size_t i = 0; i < n; ++i) {
    const Pt& p = sweep[i];
    const float x = ((extr[0] * p.x + extr[1] * p.y) + extr[2] * p.z) + extr[3];
    const float y = ((extr[4] * p.x + extr[5] * p.y) + extr[6] * p.z) + extr[7];
    const float z = ((extr[8] * p.x + extr[9] * p.y) + extr[10] * p.z) + extr[11];
    if (z > 0.0) {  // Frame.cpp:328
      Pt q;
      q.i = z;
      q.x = x * 10.f / z;
      q.y = y * 10.f / z;
      q.z = 10.f;
      out.push_back(q);
      if (src_index) src_index->push_back((int)i);
    }
  }
}

// uv: normalised keypoint coordinates [n][2].  depth[n] (0 when invalid), valid[n], nn[n][3] (-1 when the gate fails).
inline void depth_associate(const Cloud& dc, const float* uv, size_t n, float* depth, int* valid, int* nn) {
  KdTree kd;
  kd.build(dc);
  for (size_t k = 0; k < n; ++k) {
    const float u = uv[2 * k], v = uv[2 * k + 1];
    Neighbor nb[3];
    const int cnt = kd.knn(10 * u, 10 * v, 10.f, 3, nb);   // Frontend.cpp:233-241
    float s = 0.f; int ok = 0;
    nn[3 * k] = nn[3 * k + 1] = nn[3 * k + 2] = -1;
    if (cnt == 3 && nb[0].d < 0.5) {                        // :245
      for (int j = 0; j < 3; ++j) nn[3 * k + j] = nb[j].i;
      Pt dp = dc[nb[0].i];
      double x1 = dp.x * dp.i / 10, y1 = dp.y * dp.i / 10, z1 = dp.i;
      double minDepth = z1, maxDepth = z1;
      dp = dc[nb[1].i];
      double x2 = dp.x * dp.i / 10, y2 = dp.y * dp.i / 10, z2 = dp.i;
      minDepth = (z2 < minDepth) ? z2 : minDepth; maxDepth = (z2 > maxDepth) ? z2 : maxDepth;
      dp = dc[nb[2].i];
      double x3 = dp.x * dp.i / 10, y3 = dp.y * dp.i / 10, z3 = dp.i;
      minDepth = (z3 < minDepth) ? z3 : minDepth; maxDepth = (z3 > maxDepth) ? z3 : maxDepth;
      double uu = u, vv = v;
      s = (float)((x1 * y2 * z3 - x1 * y3 * z2 - x2 * y1 * z3 + x2 * y3 * z1 + x3 * y1 * z2 - x3 * y2 * z1) /
                  (x1 * y2 - x2 * y1 - x1 * y3 + x3 * y1 + x2 * y3 - x3 * y2 + uu * y1 * z2 - uu * y2 * z1 - vv * x1 * z2 + vv * x2 * z1 - uu * y1 * z3 +
                   uu * y3 * z1 + vv * x1 * z3 - vv * x3 * z1 + uu * y2 * z3 - uu * y3 * z2 - vv * x2 * z3 + vv * x3 * z2));   // :270
      ok = 1;
      if (!std::isfinite(s)) { s = (float)z1; ok = 1; }     // :274-279
      if (maxDepth - minDepth > 2) { s = 0; ok = 0; }        // :280-285
      else if (s - maxDepth > 0.2) s = (float)maxDepth;      // :286-289
      else if (s - minDepth < -0.2) s = (float)minDepth;     // :290-293
    }
    depth[k] = s; valid[k] = ok;
  }
}

}  // namespace lvo_oracle
