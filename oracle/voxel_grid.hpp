// oracle/voxel_grid.hpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).
// Restates pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.8.0, filters/include/pcl/filters/impl/voxel_grid.hpp —
// NOT in /root/reference; restated from its published algorithm, parity unpinned).  Reference call sites:
// src/scanRegistration.cpp:401-405, src/laserMapping.cpp:543-549, :793-799.
// Deviation register: PCL sorts (voxel idx, point) pairs with std::sort (unstable, centroid summation order
// unspecified for equal idx); the contract here is stable order by (idx, input index).
#pragma once
#include "common.hpp"
#include <climits>
#include <cfloat>

namespace lvo_oracle {

// idx_out (optional): voxel index of every input point; order_out (optional): the sorted permutation.
inline void voxel_grid(const Cloud& in, float leaf, Cloud& out, std::vector<int>* idx_out = nullptr,
                       std::vector<int>* order_out = nullptr) {
  out.clear();
  if (idx_out) idx_out->clear();
  if (order_out) order_out->clear();
  const size_t n = in.size();
  if (n == 0) return;
  const float inv = 1.0f / leaf;  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
  // getMinMax3D
  float mnx = FLT_MAX, mny = FLT_MAX, mnz = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX, mxz = -FLT_MAX;
  for (size_t i = 0; i < n; ++i) {
    mnx = std::min(mnx, in[i].x); mny = std::min(mny, in[i].y); mnz = std::min(mnz, in[i].z);
    mxx = std::max(mxx, in[i].x); mxy = std::max(mxy, in[i].y); mxz = std::max(mxz, in[i].z);
  }
  // "Leaf size is too small for the input dataset. Integer indices would overflow." -> output = input
  int64_t dx = static_cast<int64_t>((mxx - mnx) * inv) + 1;
  int64_t dy = static_cast<int64_t>((mxy - mny) * inv) + 1;
  int64_t dz = static_cast<int64_t>((mxz - mnz) * inv) + 1;
  if ((dx * dy * dz) > static_cast<int64_t>(INT_MAX)) {
    out = in;
    return;
  }
  const int min_b0 = static_cast<int>(std::floor(mnx * inv)), max_b0 = static_cast<int>(std::floor(mxx * inv));
  const int min_b1 = static_cast<int>(std::floor(mny * inv)), max_b1 = static_cast<int>(std::floor(mxy * inv));
  const int min_b2 = static_cast<int>(std::floor(mnz * inv)), max_b2 = static_cast<int>(std::floor(mxz * inv));
  const int div0 = max_b0 - min_b0 + 1, div1 = max_b1 - min_b1 + 1;
  (void)max_b2;
  const int mul1 = div0, mul2 = div0 * div1;

  std::vector<std::pair<unsigned int, int>> iv(n);
  for (size_t i = 0; i < n; ++i) {
    int ijk0 = static_cast<int>(std::floor(in[i].x * inv) - static_cast<float>(min_b0));
    int ijk1 = static_cast<int>(std::floor(in[i].y * inv) - static_cast<float>(min_b1));
    int ijk2 = static_cast<int>(std::floor(in[i].z * inv) - static_cast<float>(min_b2));
    int idx = ijk0 + ijk1 * mul1 + ijk2 * mul2;
    iv[i] = std::make_pair(static_cast<unsigned int>(idx), (int)i);
  }
  if (idx_out) { idx_out->resize(n); for (size_t i = 0; i < n; ++i) (*idx_out)[i] = (int)iv[i].first; }
  std::stable_sort(iv.begin(), iv.end(),
                   [](const std::pair<unsigned, int>& a, const std::pair<unsigned, int>& b) { return a.first < b.first; });
  if (order_out) { order_out->resize(n); for (size_t i = 0; i < n; ++i) (*order_out)[i] = iv[i].second; }

  // one centroid per run (min_points_per_voxel_ = 0, downsample_all_data_ = true -> intensity averaged too)
  size_t first = 0;
  while (first < n) {
    size_t last = first + 1;
    while (last < n && iv[last].first == iv[first].first) ++last;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (size_t li = first; li < last; ++li) {
      const Pt& p = in[iv[li].second];
      sx += p.x; sy += p.y; sz += p.z; si += p.i;
    }
    float cnt = static_cast<float>(last - first);
    out.push_back(Pt{sx / cnt, sy / cnt, sz / cnt, si / cnt});
    first = last;
  }
}

}  // namespace lvo_oracle
