// TEST INFRASTRUCTURE.  Compiles the reference's own scanRegistration.cpp (from /root/reference, never copied) against the
// stand-in headers of oracle/ref_shim and drives it the way ROS would: main() runs once (parameters, subscriber, publishers),
// then every raw scan is published on /velodyne_points and delivered to the reference's laserCloudHandler; the clouds it
// publishes and its file-scope work arrays (cloudCurvature, cloudLabel, cloudNeighborPicked) are read back.
// Built into oracle/_ref/ by `make -C oracle ref`.
#define main ref_scan_registration_main
#include REF_SCAN_REGISTRATION_CPP
#undef main
#include <cstring>
#include <type_traits>

extern "C" {

int ref_reg_init(int n_scans, double minimum_range) {
  ros::shim::Bus& b = ros::shim::Bus::get();
  b.params_i["scan_line"] = n_scans;
  b.params_d["minimum_range"] = minimum_range;
  int argc = 0;
  return ref_scan_registration_main(argc, nullptr);   // returns after ros::spin() (a no-op in the stand-in)
}

// one sensor_msgs::PointCloud2 with n points (x, y, z at `stride` floats) through the reference's handler
int ref_reg_process(const float* xyz, int n, int stride, double stamp) {
  sensor_msgs::PointCloud2 msg;
  msg.header.stamp = ros::Time(stamp);
  msg.xyzi.resize((size_t)n * 4);
  for (int i = 0; i < n; ++i) { msg.xyzi[4 * i] = xyz[(size_t)i * stride]; msg.xyzi[4 * i + 1] = xyz[(size_t)i * stride + 1]; msg.xyzi[4 * i + 2] = xyz[(size_t)i * stride + 2]; msg.xyzi[4 * i + 3] = 0.f; }
  ros::Publisher("/velodyne_points").publish(msg);
  ros::spinOnce();
  return 0;
}

// last cloud published on `topic`: number of points, and up to cap points (x, y, z, intensity) into out
int ref_reg_cloud(const char* topic, float* out, int cap) {
  ros::shim::Bus& b = ros::shim::Bus::get();
  auto it = b.last.find(topic);
  if (it == b.last.end()) return -1;
  const sensor_msgs::PointCloud2& m = *std::static_pointer_cast<const sensor_msgs::PointCloud2>(it->second);
  const int n = (int)(m.xyzi.size() / 4);
  if (out) std::memcpy(out, m.xyzi.data(), sizeof(float) * 4 * (size_t)(n < cap ? n : cap));
  return n;
}
void ref_reg_voxel_sort_mode(int mode) { pcl::ref_voxel_sort_mode() = mode; }   // 0 literal std::sort, 1 canonical (ties by index)
long ref_reg_published(const char* topic) { return ros::shim::Bus::get().count[topic]; }

// the reference's file-scope arrays after the last scan
void ref_reg_arrays(float* curvature, int* label, int* neighbor_picked, int n) {
  for (int i = 0; i < n; ++i) { curvature[i] = cloudCurvature[i]; label[i] = cloudLabel[i]; neighbor_picked[i] = cloudNeighborPicked[i]; }
}

// which overloads the unqualified calls at scanRegistration.cpp:166 resolve to in THIS translation unit: 8 = the C double
// functions (a build whose headers never pull <math.h>'s std overloads into the global namespace, e.g. GCC 5 of the reference's
// docker image), 4 = the float overloads
int ref_reg_atan_result_bytes() { return (int)sizeof(decltype(atan(1.0f))); }
int ref_reg_sqrt_result_bytes() { return (int)sizeof(decltype(sqrt(1.0f))); }

}  // extern "C"
