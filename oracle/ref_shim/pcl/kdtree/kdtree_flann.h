// Stand-in for pcl::KdTreeFLANN<PointXYZI> (test infrastructure, our code): forwards to oracle/kdtree.cc, which
// tests/test_oracle_vs_flann.py pins to a real FLANN build.
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include "oracle.h"
namespace pcl {
template <typename PointT>
class KdTreeFLANN {
 public:
  typedef std::shared_ptr<KdTreeFLANN<PointT>> Ptr;
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) {
    orc::Cloud a(c->points.size());
    for (size_t i = 0; i < a.size(); ++i) { a[i].x = c->points[i].x; a[i].y = c->points[i].y; a[i].z = c->points[i].z; a[i].intensity = c->points[i].intensity; }
    tree_.build(a);
  }
  int nearestKSearch(const PointT& p, int k, std::vector<int>& idx, std::vector<float>& sqd) const {
    const float q[3] = {p.x, p.y, p.z};
    idx.resize(k); sqd.resize(k);
    const int found = tree_.knn(q, k, idx.data(), sqd.data());
    idx.resize(found); sqd.resize(found);
    return found;
  }
 private:
  orc::KdTree tree_;
};
}  // namespace pcl
