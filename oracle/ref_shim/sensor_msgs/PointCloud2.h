// Stand-in for <sensor_msgs/PointCloud2.h> (test infrastructure, our code): the payload is kept as x, y, z, intensity floats.
#pragma once
#include <ros/ros.h>
namespace sensor_msgs {
struct PointCloud2 {
  std_msgs::Header header;
  std::vector<float> xyzi;    // 4 floats per point
  bool has_intensity = true;
  typedef std::shared_ptr<PointCloud2> Ptr;
  typedef std::shared_ptr<PointCloud2 const> ConstPtr;
};
typedef std::shared_ptr<PointCloud2> PointCloud2Ptr;
typedef std::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
}  // namespace sensor_msgs
