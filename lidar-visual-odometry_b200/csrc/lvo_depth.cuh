// lvo_depth.cuh — camera-lidar depth association (BASELINE config 5).
//   k_depth_flag / k_depth_write : the sweep through the 3x4 lidar->camera extrinsic (pcl::transformPointCloud, reference
//                                  src/vloam/CamLidarProcess.cpp:249-253) and the depth cloud of src/vloam/Frame.cpp:328-336:
//                                  every point with z > 0 becomes (10 x / z, 10 y / z, 10, I = z), order preserved
//                                  (flag -> exclusive scan -> scatter)
//   lvo_grid_build               : 2-D uniform grid over the clamped image plane (replaces kdTree_->setInputCloud, Frontend.cpp:466)
//   k_depth_assoc                : one warp per keypoint: exact 3-NN of (10 u, 10 v, 10) (Frontend.cpp:233-241), gate d0^2 < 0.5
//                                  (:245), planar interpolation through the three points in double (:247-270) and the clamps of
//                                  :274-293
#pragma once
#include "lvo_internal.h"
#include "lvo_knn.cuh"

struct DepthArgs {
  const float4* sweep; int n;        // packed sweep (device)
  float extr[12];                    // 3x4 row-major lidar -> camera
  unsigned* flag;                    // [n] -> exclusive scan
  int* d_n;                          // device copy of n (scan length)
  int* d_ndc;                        // number of depth-cloud points (device)
  float4* dc;                        // depth cloud out
  const float* uv; int nkp;          // keypoints, normalised coordinates [nkp][2]
  float* depth; int* valid; int* nn; // outputs [nkp], [nkp], [nkp][3]
  GridSet grid;                      // one problem over dc
};

__device__ __forceinline__ float3 cam_point(const DepthArgs& a, float4 p) {
  return make_float3(((a.extr[0] * p.x + a.extr[1] * p.y) + a.extr[2] * p.z) + a.extr[3], ((a.extr[4] * p.x + a.extr[5] * p.y) + a.extr[6] * p.z) + a.extr[7],
                     ((a.extr[8] * p.x + a.extr[9] * p.y) + a.extr[10] * p.z) + a.extr[11]);
}
__global__ void k_depth_flag(DepthArgs a) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.d_n = a.n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) a.flag[i] = cam_point(a, a.sweep[i]).z > 0.0 ? 1u : 0u;
}
__global__ void k_depth_write(DepthArgs a, const unsigned* d_total) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.d_ndc = (int)*d_total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
    const float3 c = cam_point(a, a.sweep[i]);
    if (c.z > 0.0) a.dc[a.flag[i]] = make_float4(c.x * 10.f / c.z, c.y * 10.f / c.z, 10.f, c.z);   // Frame.cpp:330-336
  }
}

__global__ void __launch_bounds__(256) k_depth_assoc(DepthArgs a) {
  const GridView g = grid_view(a.grid, 0);
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const unsigned ln = threadIdx.x & 31;
  const int ndc = *a.d_ndc;
  for (int k = wid; k < a.nkp; k += nw) {
    const float u = a.uv[2 * k], v = a.uv[2 * k + 1];
    const float qx = 10 * u, qy = 10 * v, qz = 10.f;
    TopK<3> tk;
    tk.init();
    bool found = false;
    if (g.dim[0] > 0 && ndc >= 3) {
      // the query's cell is clamped into the table like the points': a query outside the clamp box starts in the border
      // cell that holds the outside points, and all ring bounds stay valid (true distances only grow)
      const int cx = min(max(cell_coord(qx, g.inv_cell) - g.org[0], 0), g.dim[0] - 1), cy = min(max(cell_coord(qy, g.inv_cell) - g.org[1], 0), g.dim[1] - 1);
      for (int dy = -1; dy <= 1; ++dy) scan_row<3>(g, 0, cy + dy, cx - 1, cx + 1, qx, qy, qz, tk, ln);
      warp_merge<3>(tk);
      // grow the search square ring by ring (all depth points share z = 10: the grid has one layer)
      const int rmax = max(g.dim[0], g.dim[1]) + 1;
      for (int r = 2; r <= rmax; ++r) {
        const float bound = (float)(r - 1) * g.cell, b2 = bound * bound;
        if (tk.id[2] != INT_MAX && tk.d[2] < b2) break;   // the three nearest are final
        if (!(tk.d[0] < 0.5f) && b2 >= 0.5f) break;        // the gate on the nearest point can no longer pass
        scan_row<3>(g, 0, cy - r, cx - r, cx + r, qx, qy, qz, tk, ln);
        scan_row<3>(g, 0, cy + r, cx - r, cx + r, qx, qy, qz, tk, ln);
        for (int dy = -r + 1; dy <= r - 1; ++dy) {
          scan_row<3>(g, 0, cy + dy, cx - r, cx - r, qx, qy, qz, tk, ln);
          scan_row<3>(g, 0, cy + dy, cx + r, cx + r, qx, qy, qz, tk, ln);
        }
        warp_merge<3>(tk);
      }
      found = tk.id[2] != INT_MAX && tk.d[0] < 0.5f;      // Frontend.cpp:245
    }
    if (ln == 0) {
      float s = 0.f; int ok = 0;
      int* nn = a.nn + 3 * k;
      nn[0] = nn[1] = nn[2] = -1;
      if (found) {
        nn[0] = tk.id[0]; nn[1] = tk.id[1]; nn[2] = tk.id[2];
        float4 dp = a.dc[tk.id[0]];
        const double x1 = dp.x * dp.w / 10, y1 = dp.y * dp.w / 10, z1 = dp.w;
        double minDepth = z1, maxDepth = z1;
        dp = a.dc[tk.id[1]];
        const double x2 = dp.x * dp.w / 10, y2 = dp.y * dp.w / 10, z2 = dp.w;
        minDepth = (z2 < minDepth) ? z2 : minDepth; maxDepth = (z2 > maxDepth) ? z2 : maxDepth;
        dp = a.dc[tk.id[2]];
        const double x3 = dp.x * dp.w / 10, y3 = dp.y * dp.w / 10, z3 = dp.w;
        minDepth = (z3 < minDepth) ? z3 : minDepth; maxDepth = (z3 > maxDepth) ? z3 : maxDepth;
        const double uu = u, vv = v;
        s = (float)((x1 * y2 * z3 - x1 * y3 * z2 - x2 * y1 * z3 + x2 * y3 * z1 + x3 * y1 * z2 - x3 * y2 * z1) /
                    (x1 * y2 - x2 * y1 - x1 * y3 + x3 * y1 + x2 * y3 - x3 * y2 + uu * y1 * z2 - uu * y2 * z1 - vv * x1 * z2 + vv * x2 * z1 - uu * y1 * z3 +
                     uu * y3 * z1 + vv * x1 * z3 - vv * x3 * z1 + uu * y2 * z3 - uu * y3 * z2 - vv * x2 * z3 + vv * x3 * z2));   // :270
        ok = 1;
        if (!isfinite(s)) { s = (float)z1; ok = 1; }
        if (maxDepth - minDepth > 2) { s = 0; ok = 0; }
        else if (s - maxDepth > 0.2) s = (float)maxDepth;
        else if (s - minDepth < -0.2) s = (float)minDepth;
      }
      a.depth[k] = s; a.valid[k] = ok;
    }
  }
}
