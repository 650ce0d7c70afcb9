// Stand-in for <pcl/point_types.h> (test infrastructure, our code): the two point types the reference uses, by their fields.
#pragma once
namespace pcl {
struct PointXYZ { float x = 0.f, y = 0.f, z = 0.f; };
struct PointXYZI { float x = 0.f, y = 0.f, z = 0.f, intensity = 0.f; };
}  // namespace pcl
