// ORACLE (test infrastructure) -- CPU restatement of the feature-extraction stage of A-LOAM:
//   scanRegistration.cpp:85-112   min-range filter                        -> stage A below
//   scanRegistration.cpp:140-252  ring id + relative time + ring-major concat -> stage B
//   scanRegistration.cpp:256-266  11-tap curvature                        -> stage C
//   scanRegistration.cpp:277-408  per-ring, per-sixth: sort, greedy sharp / flat picks with neighbour
//                                 suppression, less-flat gather, 0.2 m voxel filter -> stage D
// Arithmetic types and evaluation order follow the reference expression by expression (float vs double
// promotion matters for the discrete decisions); only the code shape is ours.
//
// Toolchain-dependent detail made explicit (SURVEY.md 8a note 5): at :166 `atan`/`sqrt` are the unqualified
// C names; with the pinned toolchain (gcc 5) they bind to the double overloads, so the elevation angle is
// computed in double and rounded to float on assignment.  `atan2` is `using std::atan2` (:56) on floats.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "oracle.h"

namespace orc {

namespace {
const double kScanPeriod = 0.1;   // scanRegistration.cpp:60
const int kStaticCapacity = 400000;  // :66-69 file-scope arrays

struct Xyz { float x, y, z; };

// ring id of one return, or -1 when the reference drops it (:169-205)
inline int ring_of(float angle, int n_scans) {
  int id;
  if (n_scans == 16) {
    id = int((angle + 15) / 2 + 0.5);                  // float expr, then + double 0.5
    if (id > n_scans - 1 || id < 0) return -1;
  } else if (n_scans == 32) {
    id = int((angle + 92.0 / 3.0) * 3.0 / 4.0);        // truncation, no +0.5
    if (id > n_scans - 1 || id < 0) return -1;
  } else {
    if (angle >= -8.83) id = int((2 - angle) * 3.0 + 0.5);
    else id = n_scans / 2 + int((-8.83 - angle) * 2.0 + 0.5);
    if (angle > 2 || angle < -24.33 || id > 50 || id < 0) return -1;
  }
  return id;
}

// squared gap between consecutive points i and i-1... evaluated exactly like :321-324 (float), compared to double 0.05
inline bool gap_exceeds(const Cloud& c, int hi, int lo) {
  float dx = c[hi].x - c[lo].x, dy = c[hi].y - c[lo].y, dz = c[hi].z - c[lo].z;
  return dx * dx + dy * dy + dz * dz > 0.05;
}

// :317-342 / :364-388 : mark the pick and up to 5 neighbours each side until a gap
inline void suppress_around(const Cloud& c, std::vector<int>& picked, int ind) {
  picked[ind] = 1;
  for (int l = 1; l <= 5; ++l) {
    if (gap_exceeds(c, ind + l, ind + l - 1)) break;
    picked[ind + l] = 1;
  }
  for (int l = -1; l >= -5; --l) {
    if (gap_exceeds(c, ind + l, ind + l + 1)) break;
    picked[ind + l] = 1;
  }
}
}  // namespace

int extract_features(const float* xyz, int n, int stride, int n_scans, double minimum_range, SortMode mode,
                     Features& out) {
  const double t_begin = now_ms();
  if (n_scans != 16 && n_scans != 32 && n_scans != 64) return -1;  // :472-476
  out = Features();
  double t0 = now_ms();

  // ---- stage A: NaN removal (:136) + removeClosedPointCloud (:85-112); `double MINIMUM_RANGE` -> `float thres`
  std::vector<Xyz> in;
  in.reserve(n);
  const float thres = static_cast<float>(minimum_range);
  for (int i = 0; i < n; ++i) {
    const float* p = xyz + (size_t)i * stride;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    if (p[0] * p[0] + p[1] * p[1] + p[2] * p[2] < thres * thres) continue;
    in.push_back({p[0], p[1], p[2]});
  }
  const int n_in = (int)in.size();
  if (n_in == 0) return -2;  // the reference dereferences points[0] here
  if (n_in > kStaticCapacity) return -3;

  // ---- stage B: sweep start/end azimuth (:141-153), per-return ring + relTime (:160-241)
  float start_ori = -std::atan2(in[0].y, in[0].x);
  float end_ori = -std::atan2(in[n_in - 1].y, in[n_in - 1].x) + 2 * M_PI;
  if (end_ori - start_ori > 3 * M_PI) end_ori -= 2 * M_PI;
  else if (end_ori - start_ori < M_PI) end_ori += 2 * M_PI;

  std::vector<Cloud> rings(n_scans);
  bool half_passed = false;
  for (int i = 0; i < n_in; ++i) {
    const Xyz& p = in[i];
    // :166 : z / sqrt(x*x+y*y) with float products, double sqrt/atan, *180/M_PI in double, stored to float
    float angle = (float)(::atan((double)p.z / ::sqrt((double)(p.x * p.x + p.y * p.y))) * 180 / M_PI);
    const int id = ring_of(angle, n_scans);
    if (id < 0) continue;  // note: `continue` also skips the half_passed update (:174-175)

    float ori = -std::atan2(p.y, p.x);
    if (!half_passed) {
      if (ori < start_ori - M_PI / 2) ori += 2 * M_PI;
      else if (ori > start_ori + M_PI * 3 / 2) ori -= 2 * M_PI;
      if (ori - start_ori > M_PI) half_passed = true;
    } else {
      ori += 2 * M_PI;
      if (ori < end_ori - M_PI * 3 / 2) ori += 2 * M_PI;
      else if (ori > end_ori + M_PI / 2) ori -= 2 * M_PI;
    }
    float rel_time = (ori - start_ori) / (end_ori - start_ori);
    PointXYZI q{p.x, p.y, p.z, 0.f};
    q.intensity = id + kScanPeriod * rel_time;  // int + double*float -> double -> float
    rings[id].push_back(q);
  }

  Cloud& cloud = out.full;
  out.scan_start.assign(n_scans, 0);
  out.scan_end.assign(n_scans, 0);
  for (int r = 0; r < n_scans; ++r) {  // :247-252
    out.scan_start[r] = (int)cloud.size() + 5;
    cloud.insert(cloud.end(), rings[r].begin(), rings[r].end());
    out.scan_end[r] = (int)cloud.size() - 6;
  }
  const int n_pts = (int)cloud.size();
  out.times.prepare_ms = now_ms() - t0;
  t0 = now_ms();

  // ---- stage C: curvature over the concatenated cloud (:256-266), left-to-right float sums
  std::vector<float> curv(std::max(n_pts, 1), 0.f);
  std::vector<int> order(std::max(n_pts, 1), 0), picked(std::max(n_pts, 1), 0), label(std::max(n_pts, 1), 0);
  for (int i = 5; i < n_pts - 5; ++i) {
    const PointXYZI* c = &cloud[i];
    float dx = c[-5].x + c[-4].x + c[-3].x + c[-2].x + c[-1].x - 10 * c[0].x + c[1].x + c[2].x + c[3].x + c[4].x + c[5].x;
    float dy = c[-5].y + c[-4].y + c[-3].y + c[-2].y + c[-1].y - 10 * c[0].y + c[1].y + c[2].y + c[3].y + c[4].y + c[5].y;
    float dz = c[-5].z + c[-4].z + c[-3].z + c[-2].z + c[-1].z - 10 * c[0].z + c[1].z + c[2].z + c[3].z + c[4].z + c[5].z;
    curv[i] = dx * dx + dy * dy + dz * dz;
    order[i] = i;
  }
  out.times.curvature_ms = now_ms() - t0;

  // ---- stage D: per ring, six equal index ranges (:277-408)
  double t_sort = 0, t_voxel = 0;
  const double t_pick0 = now_ms();
  const float* cv = curv.data();
  for (int r = 0; r < n_scans; ++r) {
    const int s = out.scan_start[r], e = out.scan_end[r];
    if (e - s < 6) continue;
    Cloud less_flat_ring;
    for (int j = 0; j < 6; ++j) {
      const int sp = s + (e - s) * j / 6;
      const int ep = s + (e - s) * (j + 1) / 6 - 1;

      double ts = now_ms();
      if (mode == SORT_LITERAL)  // :71,288 : comparator on curvature only, libstdc++ introsort decides ties
        std::sort(order.begin() + sp, order.begin() + ep + 1, [cv](int a, int b) { return cv[a] < cv[b]; });
      else
        std::sort(order.begin() + sp, order.begin() + ep + 1,
                  [cv](int a, int b) { return cv[a] != cv[b] ? cv[a] < cv[b] : a < b; });
      t_sort += now_ms() - ts;

      // largest curvature first: 2 sharp (+less sharp), 18 more less-sharp, the 21st ends the walk (:291-344)
      int n_large = 0;
      for (int k = ep; k >= sp; --k) {
        const int ind = order[k];
        if (picked[ind] != 0 || !(curv[ind] > 0.1)) continue;
        ++n_large;
        if (n_large <= 2) {
          label[ind] = 2;
          out.sharp.push_back(cloud[ind]);
          out.less_sharp.push_back(cloud[ind]);
        } else if (n_large <= 20) {
          label[ind] = 1;
          out.less_sharp.push_back(cloud[ind]);
        } else {
          break;
        }
        suppress_around(cloud, picked, ind);
      }
      // smallest curvature first: 4 flat; the 4th is emitted but NOT marked (:346-390)
      int n_small = 0;
      for (int k = sp; k <= ep; ++k) {
        const int ind = order[k];
        if (picked[ind] != 0 || !(curv[ind] < 0.1)) continue;
        label[ind] = -1;
        out.flat.push_back(cloud[ind]);
        if (++n_small >= 4) break;
        suppress_around(cloud, picked, ind);
      }
      for (int k = sp; k <= ep; ++k)  // :392-398 (position index, not sorted index)
        if (label[k] <= 0) less_flat_ring.push_back(cloud[k]);
    }
    double tv = now_ms();
    Cloud ds;
    voxel_grid(less_flat_ring, 0.2f, mode, ds);  // :401-405
    out.less_flat.insert(out.less_flat.end(), ds.begin(), ds.end());
    t_voxel += now_ms() - tv;
  }
  out.times.sort_ms = t_sort;
  out.times.voxel_ms = t_voxel;
  out.times.pick_ms = now_ms() - t_pick0 - t_sort - t_voxel;
  out.times.whole_ms = now_ms() - t_begin;

  curv.resize(n_pts); label.resize(n_pts); picked.resize(n_pts);
  out.curvature.swap(curv);
  out.label.swap(label);
  out.picked.swap(picked);
  return 0;
}

}  // namespace orc
