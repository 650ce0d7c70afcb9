// TEST INFRASTRUCTURE.  Compiles the reference's own laserMapping.cpp (from /root/reference, never copied) against the stand-in
// headers of oracle/ref_shim and drives it frame by frame WITHOUT its worker thread: the node's main() spawns
// `std::thread mapping_process{process}` and process() is an endless loop that sleeps 2 ms between polls of its input queues.
// Two macro renames (no edit of the source) make that synchronous:
//   thread     -> an inert stand-in type, so main() initialises the node (parameters, filters, subscribers, publishers, the 4851
//                 cube clouds) and returns;
//   sleep_for  -> a hook that throws, so a call of the reference's process() handles the frames that are queued and then leaves
//                 its endless loop at the first idle poll.
// Per frame the driver publishes what laserOdometry publishes (corner_last, surf_last, full cloud, odometry pose), lets the
// node's handlers queue them, and calls process().  ceres::Solve of the stand-in is oracle/lm.cc; VoxelGrid / KdTreeFLANN /
// SelfAdjointEigenSolver / colPivHouseholderQr are the oracle's restatements (third-party semantics).
// Built into oracle/_ref/ by `make -C oracle ref`.
#include <chrono>
#include <mutex>
#include <queue>
#include <thread>
#include "oracle.h"
struct LidarEdgeFactor;
struct LidarPlaneFactor;
struct LidarPlaneNormFactor;
bool ref_block_of(const LidarEdgeFactor& f, orc::ResidualBlock* b);
bool ref_block_of(const LidarPlaneFactor& f, orc::ResidualBlock* b);
bool ref_block_of(const LidarPlaneNormFactor& f, orc::ResidualBlock* b);

namespace ref_sync {
struct FrameDone {};
struct inert_thread { template <typename F> explicit inert_thread(F&&) {} template <typename F> inert_thread(std::initializer_list<F>) {} };
}  // namespace ref_sync
namespace std {
using ref_inert_thread = ::ref_sync::inert_thread;
namespace this_thread { template <typename D> inline void ref_leave_loop(const D&) { throw ::ref_sync::FrameDone(); } }
}  // namespace std

#define thread ref_inert_thread
#define sleep_for ref_leave_loop
#define main ref_laser_mapping_main
#include REF_LASER_MAPPING_CPP
#undef main
#undef sleep_for
#undef thread
#include <cstring>

static void put3(double* d, const Eigen::Vector3d& v) { d[0] = v.x(); d[1] = v.y(); d[2] = v.z(); }
bool ref_block_of(const LidarEdgeFactor& f, orc::ResidualBlock* b) {
  b->type = orc::FACTOR_EDGE; put3(b->cp, f.curr_point); put3(b->a, f.last_point_a); put3(b->b, f.last_point_b); b->s = f.s; return true;
}
bool ref_block_of(const LidarPlaneFactor& f, orc::ResidualBlock* b) {
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

int ref_map_init(double line_res, double plane_res, int voxel_sort_mode) {
  ros::shim::Bus& b = ros::shim::Bus::get();
  b.params_d["mapping_line_resolution"] = line_res;
  b.params_d["mapping_plane_resolution"] = plane_res;
  pcl::ref_voxel_sort_mode() = voxel_sort_mode;
  int argc = 0;
  return ref_laser_mapping_main(argc, nullptr);
}

// one frame: laserOdometry's four messages, then the reference's process() until its queues are empty
int ref_map_process(const float* corner_last, int n_corner, const float* surf_last, int n_surf, const float* full, int n_full,
                    const double* q_wodom_curr, const double* t_wodom_curr, double stamp) {
  publish_cloud("/laser_cloud_corner_last", corner_last, n_corner, stamp);
  publish_cloud("/laser_cloud_surf_last", surf_last, n_surf, stamp);
  publish_cloud("/velodyne_cloud_3", full, n_full, stamp);
  nav_msgs::Odometry od;
  od.header.stamp = ros::Time(stamp);
  od.pose.pose.orientation.x = q_wodom_curr[0]; od.pose.pose.orientation.y = q_wodom_curr[1];
  od.pose.pose.orientation.z = q_wodom_curr[2]; od.pose.pose.orientation.w = q_wodom_curr[3];
  od.pose.pose.position.x = t_wodom_curr[0]; od.pose.pose.position.y = t_wodom_curr[1]; od.pose.pose.position.z = t_wodom_curr[2];
  ros::Publisher("/laser_odom_to_init").publish(od);
  ros::spinOnce();                          // the node's four handlers queue the messages
  try { process(); } catch (const ref_sync::FrameDone&) {}
  return 0;
}

// parameters[7] (q_w_curr x, y, z, w ; t_w_curr), q_wmap_wodom (x, y, z, w), t_wmap_wodom, centre indices, frame count
void ref_map_state(double* pose7, double* q_wmap_wodom4, double* t_wmap_wodom3, int* centre3, int* frames, int* n_valid) {
  for (int k = 0; k < 7; ++k) pose7[k] = parameters[k];
  q_wmap_wodom4[0] = q_wmap_wodom.x(); q_wmap_wodom4[1] = q_wmap_wodom.y(); q_wmap_wodom4[2] = q_wmap_wodom.z(); q_wmap_wodom4[3] = q_wmap_wodom.w();
  put3(t_wmap_wodom3, t_wmap_wodom);
  centre3[0] = laserCloudCenWidth; centre3[1] = laserCloudCenHeight; centre3[2] = laserCloudCenDepth;
  *frames = frameCount;
  *n_valid = 0;
}

// cube `index` of the corner (which = 0) or surf (1) array: number of points, and up to cap points into out
int ref_map_cube(int which, int index, float* out, int cap) {
  if (index < 0 || index >= laserCloudNum) return -1;
  const pcl::PointCloud<PointType>& c = which ? *laserCloudSurfArray[index] : *laserCloudCornerArray[index];
  const int n = (int)c.points.size();
  for (int i = 0; i < n && i < cap; ++i) { out[4 * i] = c.points[i].x; out[4 * i + 1] = c.points[i].y; out[4 * i + 2] = c.points[i].z; out[4 * i + 3] = c.points[i].intensity; }
  return n;
}
// sizes of all cubes of one array (laserCloudNum entries)
int ref_map_cube_sizes(int which, int* sizes) {
  for (int i = 0; i < laserCloudNum; ++i) sizes[i] = (int)(which ? laserCloudSurfArray[i] : laserCloudCornerArray[i])->points.size();
  return laserCloudNum;
}

// the pose published on /aft_mapped_to_init ; returns the number of messages published so far
long ref_map_published_pose(double* q, double* t) {
  ros::shim::Bus& b = ros::shim::Bus::get();
  auto it = b.last.find("/aft_mapped_to_init");
  if (it == b.last.end()) return 0;
  const nav_msgs::Odometry& m = *std::static_pointer_cast<const nav_msgs::Odometry>(it->second);
  q[0] = m.pose.pose.orientation.x; q[1] = m.pose.pose.orientation.y; q[2] = m.pose.pose.orientation.z; q[3] = m.pose.pose.orientation.w;
  t[0] = m.pose.pose.position.x; t[1] = m.pose.pose.position.y; t[2] = m.pose.pose.position.z;
  return b.count["/aft_mapped_to_init"];
}

}  // extern "C"
