// Stand-in for the tf types the mapping node fills in before broadcasting (test infrastructure, our code): nothing is computed with them.
#pragma once
#include <ros/ros.h>
namespace tf {
struct Vector3 { double x, y, z; Vector3(double a = 0, double b = 0, double c = 0) : x(a), y(b), z(c) {} };
struct Quaternion {
  double x_ = 0, y_ = 0, z_ = 0, w_ = 1;
  void setW(double v) { w_ = v; } void setX(double v) { x_ = v; } void setY(double v) { y_ = v; } void setZ(double v) { z_ = v; }
};
struct Transform { Vector3 origin; Quaternion rotation; void setOrigin(const Vector3& o) { origin = o; } void setRotation(const Quaternion& q) { rotation = q; } };
struct StampedTransform { StampedTransform(const Transform&, const ros::Time&, const std::string&, const std::string&) {} };
}  // namespace tf
