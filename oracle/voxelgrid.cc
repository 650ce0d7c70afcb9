// ORACLE (test infrastructure) -- restatement of pcl::VoxelGrid<pcl::PointXYZI>::applyFilter (PCL 1.8.0,
// filters/include/pcl/filters/impl/voxel_grid.hpp) as the reference uses it:
//   scanRegistration.cpp:401-407 (leaf 0.2 per ring), laserMapping.cpp:543-549,787-801 (lineRes / planeRes).
// PCL is not in /root/reference (un-vendored, pinned by docker/Dockerfile:4); semantics per SURVEY.md 8a "V".
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <limits>
#include "oracle.h"

namespace orc {

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

namespace {
struct cloud_point_index_idx {
  unsigned int idx;
  unsigned int cloud_point_index;
  bool operator<(const cloud_point_index_idx& p) const { return idx < p.idx; }  // PCL compares idx only
};
}  // namespace

void voxel_grid(const Cloud& in, float leaf, SortMode mode, Cloud& out) {
  out.clear();
  if (in.empty()) return;
  // leaf_size_ = (leaf,leaf,leaf,1) ; inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
  const float inv = 1.0f / leaf;
  // getMinMax3D (dense cloud path; NaNs never reach here -- removeNaN upstream)
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (const PointXYZI& p : in) {
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv) + 1;
  if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
    out = in;  // "Leaf size is too small for the input dataset. Integer indices would overflow."
    return;
  }
  int min_b[3], max_b[3], div_b[3], divb_mul[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = static_cast<int>(std::floor(mn[a] * inv));
    max_b[a] = static_cast<int>(std::floor(mx[a] * inv));
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  divb_mul[0] = 1; divb_mul[1] = div_b[0]; divb_mul[2] = div_b[0] * div_b[1];

  std::vector<cloud_point_index_idx> index_vector;
  index_vector.reserve(in.size());
  for (unsigned int i = 0; i < in.size(); ++i) {
    const PointXYZI& p = in[i];
    int ijk0 = static_cast<int>(std::floor(p.x * inv) - static_cast<float>(min_b[0]));
    int ijk1 = static_cast<int>(std::floor(p.y * inv) - static_cast<float>(min_b[1]));
    int ijk2 = static_cast<int>(std::floor(p.z * inv) - static_cast<float>(min_b[2]));
    int idx = ijk0 * divb_mul[0] + ijk1 * divb_mul[1] + ijk2 * divb_mul[2];
    index_vector.push_back({static_cast<unsigned int>(idx), i});
  }
  if (mode == SORT_LITERAL) {
    std::sort(index_vector.begin(), index_vector.end(), std::less<cloud_point_index_idx>());
  } else {
    std::sort(index_vector.begin(), index_vector.end(),
              [](const cloud_point_index_idx& a, const cloud_point_index_idx& b) {
                return a.idx != b.idx ? a.idx < b.idx : a.cloud_point_index < b.cloud_point_index;
              });
  }
  // one output per occupied voxel, ascending idx; centroid of all 4 fields (downsample_all_data_ = true)
  // accumulated in float in sorted order (pcl::CentroidPoint: AccumulatorXYZ Vector3f, AccumulatorIntensity float)
  size_t i = 0;
  while (i < index_vector.size()) {
    size_t j = i + 1;
    while (j < index_vector.size() && index_vector[j].idx == index_vector[i].idx) ++j;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (size_t k = i; k < j; ++k) {
      const PointXYZI& p = in[index_vector[k].cloud_point_index];
      sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
    }
    const float n = static_cast<float>(j - i);
    out.push_back({sx / n, sy / n, sz / n, si / n});
    i = j;
  }
}

}  // namespace orc
