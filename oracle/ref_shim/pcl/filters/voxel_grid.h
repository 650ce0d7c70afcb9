// Stand-in for pcl::VoxelGrid<PointXYZI> (test infrastructure, our code): forwards to the restatement in oracle/voxelgrid.cc
// (the filter's semantics are third-party: DESIGN.md section 2, row 12).
#pragma once
#include <pcl/filters/filter.h>
#include <pcl/point_types.h>
#include "oracle.h"
namespace pcl {
// how ties of PCL's unstable std::sort are resolved (oracle.h SortMode); set by the driver
inline int& ref_voxel_sort_mode() { static int mode = orc::SORT_CANONICAL; return mode; }
template <typename PointT>
class VoxelGrid {
 public:
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { in_ = c; }
  void setLeafSize(float lx, float, float) { leaf_ = lx; }
  void filter(PointCloud<PointT>& out) {
    orc::Cloud a(in_->points.size()), b;
    for (size_t i = 0; i < a.size(); ++i) { a[i].x = in_->points[i].x; a[i].y = in_->points[i].y; a[i].z = in_->points[i].z; a[i].intensity = in_->points[i].intensity; }
    orc::voxel_grid(a, leaf_, (orc::SortMode)ref_voxel_sort_mode(), b);
    out.header = in_->header;
    out.points.resize(b.size());
    for (size_t i = 0; i < b.size(); ++i) { out.points[i].x = b[i].x; out.points[i].y = b[i].y; out.points[i].z = b[i].z; out.points[i].intensity = b[i].intensity; }
    out.width = (uint32_t)b.size(); out.height = 1; out.is_dense = true;
  }
 private:
  typename PointCloud<PointT>::ConstPtr in_;
  float leaf_ = 0.f;
};
}  // namespace pcl
