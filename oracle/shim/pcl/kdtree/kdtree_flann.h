// oracle/shim/pcl/kdtree/kdtree_flann.h — TEST INFRASTRUCTURE ONLY: pcl::KdTreeFLANN<PointXYZI> with the interface the reference uses
// (setInputCloud / nearestKSearch; laserOdometry.cpp:95-97,386,470,640-641, laserMapping.cpp:107-108,558-559,582,648).  FLANN is
// absent: the search is the oracle's exact kd-tree (oracle/knn.hpp: FLANN L2_Simple float accumulation, results sorted by
// (distance, index); [3P-mem], parity unpinned for tie order only — the search itself is exact like FLANN's with eps = 0).
#pragma once
#include <pcl/point_cloud.h>
#include "knn.hpp"   // oracle/knn.hpp via -I<oracle>
namespace pcl {
template <typename PointT>
class KdTreeFLANN {
 public:
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  typedef typename PointCloud<PointT>::ConstPtr PointCloudConstPtr;
  void setInputCloud(const PointCloudConstPtr& cloud) {
    cloud_.resize(cloud->points.size());
    for (size_t i = 0; i < cloud_.size(); ++i) cloud_[i] = lvo_oracle::Pt{cloud->points[i].x, cloud->points[i].y, cloud->points[i].z, cloud->points[i].intensity};
    tree_.build(cloud_);
  }
  int nearestKSearch(const PointT& point, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    if (k > (int)cloud_.size()) k = (int)cloud_.size();   // kdtree_flann.hpp: "if (k > total_nr_points_) k = total_nr_points_"
    k_indices.resize(k); k_sqr_distances.resize(k);
    if (k == 0) return 0;
    std::vector<lvo_oracle::Neighbor> nb(k);
    const int got = tree_.knn(point.x, point.y, point.z, k, nb.data());
    for (int i = 0; i < got; ++i) { k_indices[i] = nb[i].i; k_sqr_distances[i] = nb[i].d; }
    return got;
  }
 private:
  lvo_oracle::Cloud cloud_;
  lvo_oracle::KdTree tree_;
};
}
