// Stand-in for pcl::fromROSMsg / pcl::toROSMsg (test infrastructure, our code).
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
inline void fromROSMsg(const sensor_msgs::PointCloud2& m, PointCloud<PointXYZ>& c) {
  const size_t n = m.xyzi.size() / 4;
  c.points.resize(n);
  for (size_t i = 0; i < n; ++i) { c.points[i].x = m.xyzi[4 * i]; c.points[i].y = m.xyzi[4 * i + 1]; c.points[i].z = m.xyzi[4 * i + 2]; }
  c.width = (uint32_t)n; c.height = 1; c.is_dense = false;
}
inline void fromROSMsg(const sensor_msgs::PointCloud2& m, PointCloud<PointXYZI>& c) {
  const size_t n = m.xyzi.size() / 4;
  c.points.resize(n);
  for (size_t i = 0; i < n; ++i) { c.points[i].x = m.xyzi[4 * i]; c.points[i].y = m.xyzi[4 * i + 1]; c.points[i].z = m.xyzi[4 * i + 2]; c.points[i].intensity = m.xyzi[4 * i + 3]; }
  c.width = (uint32_t)n; c.height = 1; c.is_dense = true;
}
inline void toROSMsg(const PointCloud<PointXYZI>& c, sensor_msgs::PointCloud2& m) {
  m.xyzi.resize(c.points.size() * 4);
  for (size_t i = 0; i < c.points.size(); ++i) { m.xyzi[4 * i] = c.points[i].x; m.xyzi[4 * i + 1] = c.points[i].y; m.xyzi[4 * i + 2] = c.points[i].z; m.xyzi[4 * i + 3] = c.points[i].intensity; }
}
inline void toROSMsg(const PointCloud<PointXYZ>& c, sensor_msgs::PointCloud2& m) {
  m.xyzi.assign(c.points.size() * 4, 0.f);
  m.has_intensity = false;
  for (size_t i = 0; i < c.points.size(); ++i) { m.xyzi[4 * i] = c.points[i].x; m.xyzi[4 * i + 1] = c.points[i].y; m.xyzi[4 * i + 2] = c.points[i].z; }
}
}  // namespace pcl
