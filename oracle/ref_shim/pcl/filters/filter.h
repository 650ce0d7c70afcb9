// Stand-in for pcl::removeNaNFromPointCloud (test infrastructure, our code): order-preserving, in place allowed.
#pragma once
#include <cmath>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT>
void removeNaNFromPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, std::vector<int>& index) {
  if (&in != &out) { out.header = in.header; out.points.resize(in.points.size()); }
  index.resize(in.points.size());
  size_t j = 0;
  for (size_t i = 0; i < in.points.size(); ++i) {
    if (!std::isfinite(in.points[i].x) || !std::isfinite(in.points[i].y) || !std::isfinite(in.points[i].z)) continue;
    out.points[j] = in.points[i];
    index[j] = (int)i;
    ++j;
  }
  if (j != in.points.size()) { out.points.resize(j); index.resize(j); }
  out.height = 1; out.width = (uint32_t)j; out.is_dense = true;
}
}  // namespace pcl
