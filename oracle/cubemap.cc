// ORACLE (test infrastructure, NOT product code) -- see oracle.h.  Pinned to the reference's laserMapping.cpp source by tests/test_oracle_vs_reference_source.py.
//
// The map cube store of alaserMapping and the per-frame loop around it (laserMapping.cpp:74-108 state, :142-163
// pose hand-off, :309-529 centre cube / shift / gather, :541-550 stack filters, :554-734 optimisation through
// orc::Mapping, :736-801 insertion and per-cube re-filter).  SURVEY.md section 8 f-1: the NEXT row of the scope
// table -- this restatement exists so that a device-resident cube store can be checked frame by frame.
#include <cmath>
#include "oracle.h"

namespace orc {

namespace {
constexpr int W = 21, H = 21, D = 11;   // laserCloudWidth / Height / Depth (:77-79)
inline int cube_index(int i, int j, int k) { return i + W * j + W * H * k; }

// int((v + 25.0) / 50.0) + centre, minus one when v + 25.0 < 0  (:316-325, :741-750) ; v is a float promoted to double
// for map points (pointSel.x) and a double for the pose
inline int cube_coord(double v, int centre) {
  int c = int((v + 25.0) / 50.0) + centre;
  if (v + 25.0 < 0) c--;
  return c;
}
}  // namespace

CubeMap::CubeMap() : corner(W * H * D), surf(W * H * D) {}

// :327-509 -- the six shift loops.  The reference rotates POINTERS along one axis and clears the cube that wraps
// around; rotating the vectors (std::swap / move) is the same thing.
void CubeMap::shift_for(int& ci, int& cj, int& ck) {
  auto rot = [&](int axis, bool towards_high) {
    // towards_high: every cube moves one step up along `axis`, the top one wraps to index 0 and is cleared
    const int n[3] = {W, H, D};
    const int a = axis, b = (axis + 1) % 3, c = (axis + 2) % 3;
    for (int u = 0; u < n[b]; ++u) {
      for (int v = 0; v < n[c]; ++v) {
        auto at = [&](int t) { int ijk[3]; ijk[a] = t; ijk[b] = u; ijk[c] = v; return cube_index(ijk[0], ijk[1], ijk[2]); };
        if (towards_high) {
          Cloud kc = std::move(corner[at(n[a] - 1)]), ks = std::move(surf[at(n[a] - 1)]);
          for (int t = n[a] - 1; t >= 1; --t) { corner[at(t)] = std::move(corner[at(t - 1)]); surf[at(t)] = std::move(surf[at(t - 1)]); }
          kc.clear(); ks.clear();
          corner[at(0)] = std::move(kc); surf[at(0)] = std::move(ks);
        } else {
          Cloud kc = std::move(corner[at(0)]), ks = std::move(surf[at(0)]);
          for (int t = 0; t < n[a] - 1; ++t) { corner[at(t)] = std::move(corner[at(t + 1)]); surf[at(t)] = std::move(surf[at(t + 1)]); }
          kc.clear(); ks.clear();
          corner[at(n[a] - 1)] = std::move(kc); surf[at(n[a] - 1)] = std::move(ks);
        }
      }
    }
  };
  while (ci < 3) { rot(0, true); ci++; cen_w++; }              // :327-355
  while (ci >= W - 3) { rot(0, false); ci--; cen_w--; }        // :357-385
  while (cj < 3) { rot(1, true); cj++; cen_h++; }              // :387-415
  while (cj >= H - 3) { rot(1, false); cj--; cen_h--; }        // :417-445
  while (ck < 3) { rot(2, true); ck++; cen_d++; }              // :447-475
  while (ck >= D - 3) { rot(2, false); ck--; cen_d--; }        // :477-509
}

void CubeMap::insert(const Cloud& stack, const double x[7], std::vector<Cloud>& cubes) {
  const Quat q{x[0], x[1], x[2], x[3]};
  const Vec3 t{x[4], x[5], x[6]};
  for (const PointXYZI& p : stack) {
    // pointAssociateToMap :154-163 : double transform, stored back as float
    const Vec3 pw = rotate(q, Vec3{(double)p.x, (double)p.y, (double)p.z}) + t;
    PointXYZI s{(float)pw.x, (float)pw.y, (float)pw.z, p.intensity};
    const int ci = cube_coord((double)s.x, cen_w), cj = cube_coord((double)s.y, cen_h), ck = cube_coord((double)s.z, cen_d);   // :741-750
    if (ci >= 0 && ci < W && cj >= 0 && cj < H && ck >= 0 && ck < D) cubes[cube_index(ci, cj, ck)].push_back(s);            // :752-758
  }
}

int CubeMap::step(const Cloud& corner_last, const Cloud& surf_last, const double q_wodom_curr[4], const double t_wodom_curr[3],
                  float line_res, float plane_res, int outer_iters, const SolveOptions& opt, SortMode mode) {
  double x[7];
  transform_associate_to_map(q_wmap_wodom, t_wmap_wodom, q_wodom_curr, t_wodom_curr, x);   // :311
  int ci = cube_coord(x[4], cen_w), cj = cube_coord(x[5], cen_h), ck = cube_coord(x[6], cen_d);   // :314-325
  shift_for(ci, cj, ck);
  // :511-529 valid cubes, i outermost
  valid.clear();
  for (int i = ci - 2; i <= ci + 2; ++i)
    for (int j = cj - 2; j <= cj + 2; ++j)
      for (int k = ck - 1; k <= ck + 1; ++k)
        if (i >= 0 && i < W && j >= 0 && j < H && k >= 0 && k < D) valid.push_back(cube_index(i, j, k));
  // :531-539
  corner_from_map.clear(); surf_from_map.clear();
  for (int ind : valid) {
    corner_from_map.insert(corner_from_map.end(), corner[ind].begin(), corner[ind].end());
    surf_from_map.insert(surf_from_map.end(), surf[ind].begin(), surf[ind].end());
  }
  // :541-550
  voxel_grid(corner_last, line_res, mode, corner_stack);
  voxel_grid(surf_last, plane_res, mode, surf_stack);
  // :554-733
  Mapping m;
  m.set_map(corner_from_map, surf_from_map);
  const int optimised = m.register_scan(corner_stack, surf_stack, x, outer_iters, opt);
  transform_update(x, q_wodom_curr, t_wodom_curr, q_wmap_wodom, t_wmap_wodom);   // :734
  for (int k = 0; k < 7; ++k) pose[k] = x[k];
  // :736-788
  insert(corner_stack, x, corner);
  insert(surf_stack, x, surf);
  for (int ind : valid) {
    Cloud tc, ts;
    voxel_grid(corner[ind], line_res, mode, tc); corner[ind].swap(tc);
    voxel_grid(surf[ind], plane_res, mode, ts); surf[ind].swap(ts);
  }
  ++frames;
  return optimised;
}

}  // namespace orc
