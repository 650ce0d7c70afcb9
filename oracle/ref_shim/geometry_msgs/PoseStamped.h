// Stand-in for the geometry_msgs types the reference fills in (test infrastructure, our code).
#pragma once
#include <ros/ros.h>
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
}  // namespace geometry_msgs
