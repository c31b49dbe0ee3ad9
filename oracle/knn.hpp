// oracle/knn.hpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).
// Stands in for pcl::KdTreeFLANN<PointXYZI>::nearestKSearch (PCL 1.8.0 -> FLANN KDTreeSingleIndex, L2_Simple<float>,
// eps 0, sorted results; NOT in /root/reference, parity unpinned).  Reference call sites:
// src/laserOdometry.cpp:386,470,640-641; src/laserMapping.cpp:558-559,582,648.
// Exact search; squared distance is FLANN's L2_Simple accumulation ((dx*dx + dy*dy) + dz*dz) in float.
// Deviation register: result order on exact float ties is (distance, index) ascending.
#pragma once
#include "common.hpp"
#include <cfloat>

namespace lvo_oracle {

inline float sqdist(const Pt& a, float qx, float qy, float qz) {
  float dx = a.x - qx, dy = a.y - qy, dz = a.z - qz;
  return dx * dx + dy * dy + dz * dz;
}

struct Neighbor { float d; int i; };
inline bool nb_less(const Neighbor& a, const Neighbor& b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }

// Brute force; fills up to K results sorted by (d, i); returns the number found.
inline int knn_brute(const Cloud& c, float qx, float qy, float qz, int K, Neighbor* out) {
  int cnt = 0;
  for (int i = 0; i < (int)c.size(); ++i) {
    Neighbor nb{sqdist(c[i], qx, qy, qz), i};
    if (cnt < K) { out[cnt++] = nb; }
    else if (nb_less(nb, out[K - 1])) out[K - 1] = nb;
    else continue;
    for (int k = cnt - 1; k > 0 && nb_less(out[k], out[k - 1]); --k) std::swap(out[k], out[k - 1]);
  }
  return cnt;
}

// Exact kd-tree (median split on the widest axis, leaf size 16).  Pruning uses the single-axis plane distance,
// which is a safe lower bound of the float L2_Simple value (monotone rounding), and is strict, so ties are kept.
class KdTree {
 public:
  void build(const Cloud& c) {
    cloud_ = &c;
    idx_.resize(c.size());
    for (size_t i = 0; i < c.size(); ++i) idx_[i] = (int)i;
    nodes_.clear();
    if (!c.empty()) { nodes_.reserve(c.size() / 4 + 4); build_rec(0, (int)c.size()); }
  }
  int knn(float qx, float qy, float qz, int K, Neighbor* out) const {
    int cnt = 0;
    if (nodes_.empty()) return 0;
    float q[3] = {qx, qy, qz};
    search(0, q, K, out, cnt);
    return cnt;
  }
  size_t size() const { return idx_.size(); }

 private:
  struct Node { int lo, hi, axis, left, right; float split; };
  const Cloud* cloud_ = nullptr;
  std::vector<int> idx_;
  std::vector<Node> nodes_;

  static float coord(const Pt& p, int a) { return a == 0 ? p.x : (a == 1 ? p.y : p.z); }

  int build_rec(int lo, int hi) {
    int id = (int)nodes_.size();
    nodes_.push_back(Node{lo, hi, -1, -1, -1, 0.f});
    if (hi - lo <= 16) return id;
    const Cloud& c = *cloud_;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = lo; k < hi; ++k)
      for (int a = 0; a < 3; ++a) { float v = coord(c[idx_[k]], a); mn[a] = std::min(mn[a], v); mx[a] = std::max(mx[a], v); }
    int axis = 0;
    if (mx[1] - mn[1] > mx[axis] - mn[axis]) axis = 1;
    if (mx[2] - mn[2] > mx[axis] - mn[axis]) axis = 2;
    if (!(mx[axis] > mn[axis])) return id;  // all points identical: keep as a leaf
    int mid = (lo + hi) / 2;
    std::nth_element(idx_.begin() + lo, idx_.begin() + mid, idx_.begin() + hi,
                     [&](int a, int b) { return coord(c[a], axis) < coord(c[b], axis); });
    float split = coord(c[idx_[mid]], axis);
    nodes_[id].axis = axis;
    nodes_[id].split = split;
    int l = build_rec(lo, mid);
    int r = build_rec(mid, hi);
    nodes_[id].left = l; nodes_[id].right = r;
    return id;
  }

  void search(int id, const float* q, int K, Neighbor* out, int& cnt) const {
    const Node& nd = nodes_[id];
    const Cloud& c = *cloud_;
    if (nd.axis < 0) {
      for (int k = nd.lo; k < nd.hi; ++k) {
        int i = idx_[k];
        Neighbor nb{sqdist(c[i], q[0], q[1], q[2]), i};
        if (cnt < K) out[cnt++] = nb;
        else if (nb_less(nb, out[K - 1])) out[K - 1] = nb;
        else continue;
        for (int j = cnt - 1; j > 0 && nb_less(out[j], out[j - 1]); --j) std::swap(out[j], out[j - 1]);
      }
      return;
    }
    float diff = q[nd.axis] - nd.split;
    int near = diff < 0 ? nd.left : nd.right, far = diff < 0 ? nd.right : nd.left;
    search(near, q, K, out, cnt);
    if (cnt < K || !(diff * diff > out[K - 1].d)) search(far, q, K, out, cnt);
  }
};

}  // namespace lvo_oracle
