// TEST INFRASTRUCTURE.  Compiles the reference's own laserOdometry.cpp (from /root/reference, never copied) against the stand-in
// headers of oracle/ref_shim and drives it the way ROS would: its main() is entered once to subscribe and advertise, then, per
// scan, the five clouds of scanRegistration are published on its input topics and main() is entered again with ros::ok() true
// for exactly one turn of its loop (file-scope state -- the "last" clouds, kd-trees, para_q / para_t, q_w_curr / t_w_curr --
// carries over; the loop-local frame counter restarts, which is the mapping_skip_frame = 1 behaviour of the launch files).
// ceres::Solve of the stand-in hands the residual blocks to oracle/lm.cc.  Built into oracle/_ref/ by `make -C oracle ref`.
#include "oracle.h"
struct LidarEdgeFactor;
struct LidarPlaneFactor;
struct LidarPlaneNormFactor;
bool ref_block_of(const LidarEdgeFactor& f, orc::ResidualBlock* b);
bool ref_block_of(const LidarPlaneFactor& f, orc::ResidualBlock* b);
bool ref_block_of(const LidarPlaneNormFactor& f, orc::ResidualBlock* b);

#define main ref_laser_odometry_main
#include REF_LASER_ODOMETRY_CPP
#undef main
#include <cstring>

static void put3(double* d, const Eigen::Vector3d& v) { d[0] = v.x(); d[1] = v.y(); d[2] = v.z(); }
bool ref_block_of(const LidarEdgeFactor& f, orc::ResidualBlock* b) {
  b->type = orc::FACTOR_EDGE; put3(b->cp, f.curr_point); put3(b->a, f.last_point_a); put3(b->b, f.last_point_b); b->s = f.s; return true;
}
bool ref_block_of(const LidarPlaneFactor& f, orc::ResidualBlock* b) {   // ljm_norm is the one the reference's constructor computed
  b->type = orc::FACTOR_PLANE; put3(b->cp, f.curr_point); put3(b->a, f.last_point_j); put3(b->b, f.ljm_norm); b->s = f.s; return true;
}
bool ref_block_of(const LidarPlaneNormFactor& f, orc::ResidualBlock* b) {
  b->type = orc::FACTOR_PLANE_NORM; put3(b->cp, f.curr_point); put3(b->a, f.plane_unit_norm); b->b[0] = b->b[1] = b->b[2] = 0; b->s = f.negative_OA_dot_norm; return true;
}

namespace {
void publish_cloud(const char* topic, const float* xyzi, int n, double stamp) {
  sensor_msgs::PointCloud2 m;
  m.header.stamp = ros::Time(stamp);
  m.xyzi.assign(xyzi, xyzi + (size_t)n * 4);
  ros::Publisher(topic).publish(m);
}
}  // namespace

extern "C" {

int ref_odom_init(int mapping_skip_frame) {
  ros::shim::Bus& b = ros::shim::Bus::get();
  b.params_i["mapping_skip_frame"] = mapping_skip_frame;
  b.ok_budget = 0;
  int argc = 0;
  return ref_laser_odometry_main(argc, nullptr);   // subscribes, advertises, leaves its loop at once
}

// one scan: the five clouds scanRegistration publishes (x, y, z, intensity packed), then one turn of the reference's loop
int ref_odom_process(const float* sharp, int n_sharp, const float* less_sharp, int n_less_sharp, const float* flat, int n_flat,
                     const float* less_flat, int n_less_flat, const float* full, int n_full, double stamp) {
  publish_cloud("/laser_cloud_sharp", sharp, n_sharp, stamp);
  publish_cloud("/laser_cloud_less_sharp", less_sharp, n_less_sharp, stamp);
  publish_cloud("/laser_cloud_flat", flat, n_flat, stamp);
  publish_cloud("/laser_cloud_less_flat", less_flat, n_less_flat, stamp);
  publish_cloud("/velodyne_cloud_2", full, n_full, stamp);
  ros::shim::Bus::get().ok_budget = 1;
  int argc = 0;
  return ref_laser_odometry_main(argc, nullptr);
}

// state after the last scan: para_q (x, y, z, w), para_t, q_w_curr (x, y, z, w), t_w_curr, correspondence counts of the last pass
void ref_odom_state(double* q_last_curr, double* t_last_curr, double* q_w, double* t_w, int* counts2) {
  for (int k = 0; k < 4; ++k) q_last_curr[k] = para_q[k];
  for (int k = 0; k < 3; ++k) t_last_curr[k] = para_t[k];
  q_w[0] = q_w_curr.x(); q_w[1] = q_w_curr.y(); q_w[2] = q_w_curr.z(); q_w[3] = q_w_curr.w();
  t_w[0] = t_w_curr.x(); t_w[1] = t_w_curr.y(); t_w[2] = t_w_curr.z();
  counts2[0] = corner_correspondence; counts2[1] = plane_correspondence;
}

// the pose published on /laser_odom_to_init (q x, y, z, w ; t) ; returns the number of messages published so far
long ref_odom_published_pose(double* q, double* t) {
  ros::shim::Bus& b = ros::shim::Bus::get();
  auto it = b.last.find("/laser_odom_to_init");
  if (it == b.last.end()) return 0;
  const nav_msgs::Odometry& m = *std::static_pointer_cast<const nav_msgs::Odometry>(it->second);
  q[0] = m.pose.pose.orientation.x; q[1] = m.pose.pose.orientation.y; q[2] = m.pose.pose.orientation.z; q[3] = m.pose.pose.orientation.w;
  t[0] = m.pose.pose.position.x; t[1] = m.pose.pose.position.y; t[2] = m.pose.pose.position.z;
  return b.count["/laser_odom_to_init"];
}

int ref_odom_cloud(const char* topic, float* out, int cap) {
  ros::shim::Bus& b = ros::shim::Bus::get();
  auto it = b.last.find(topic);
  if (it == b.last.end()) return -1;
  const sensor_msgs::PointCloud2& m = *std::static_pointer_cast<const sensor_msgs::PointCloud2>(it->second);
  const int n = (int)(m.xyzi.size() / 4);
  if (out) std::memcpy(out, m.xyzi.data(), sizeof(float) * 4 * (size_t)(n < cap ? n : cap));
  return n;
}

// TransformToStart (laserOdometry.cpp:111-129) on one point with the current para_q / para_t
void ref_odom_transform_to_start(const float* in_xyzi, float* out_xyzi) {
  PointType pi, po;
  pi.x = in_xyzi[0]; pi.y = in_xyzi[1]; pi.z = in_xyzi[2]; pi.intensity = in_xyzi[3];
  TransformToStart(&pi, &po);
  out_xyzi[0] = po.x; out_xyzi[1] = po.y; out_xyzi[2] = po.z; out_xyzi[3] = po.intensity;
}

}  // extern "C"
