// Scan-to-scan correspondence search on the GPU -- replaces laserOdometry.cpp:111-129 (TransformToStart),
// :299-483 (the two association loops) and :567-568 (the kd-tree builds).
//
// Index (k_tile_bounds): the "last" clouds are ring-major and, inside a ring, azimuth ordered, so 32 consecutive
// points are spatially compact.  One AABB per 32-point tile + the ring offset table is the whole index; it costs
// one pass over the cloud instead of two O(M log M) kd-tree builds per frame.
// Search (k_odom_assoc): ONE WARP PER QUERY.  Lanes test 32 tile boxes at a time against the current best
// (exact lower bound: same float expression, monotone rounding => no slack needed), surviving tiles are read
// with one coalesced 512-byte float4 load per tile, candidates are reduced with warp REDUX arg-min.
// The result is the exact nearest neighbour under the (distance, index) order, i.e. what FLANN returns up to
// exact-distance ties.  The ring-window scans of :312-361 / :402-455 reuse the same routine on the index
// ranges the ring offset table gives, with the reference's visiting order as tie-break
// (forward ascending first, then backward descending, strict '<').
#include <climits>
#include <cfloat>
#include "common.cuh"
#include "kernels.h"

namespace aloam {

__global__ void k_ring_offsets(const Pt4* __restrict__ pts, int n, int* __restrict__ ring_start, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = (int)pts[i].i;  // int(intensity) = scan line (laserOdometry.cpp:308,398)
  if (r < 0 || r > 63) { atomicExch(err, 1); return; }
  if (i == 0) {
    for (int k = 0; k <= r; ++k) ring_start[k] = 0;
  } else {
    const int rp = (int)pts[i - 1].i;
    if (r < rp) atomicExch(err, 1);
    for (int k = max(rp, 0) + 1; k <= r; ++k) ring_start[k] = i;
  }
  if (i == n - 1)
    for (int k = r + 1; k <= 64; ++k) ring_start[k] = n;
}

__global__ void __launch_bounds__(256) k_tile_bounds(const Pt4* __restrict__ pts, const int* __restrict__ n_ptr,
                                                     float* __restrict__ tile_lo, float* __restrict__ tile_hi) {
  const int n = *n_ptr;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int i = t * ALOAM_TILE + lane;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (i < n) {
    Pt4 p = pts[i];
    lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
    }
  if (lane == 0) {  // tiles past the end of the cloud get an empty box (never selected)
    reinterpret_cast<float4*>(tile_lo)[t] = make_float4(lo[0], lo[1], lo[2], 0.f);
    reinterpret_cast<float4*>(tile_hi)[t] = make_float4(hi[0], hi[1], hi[2], 0.f);
  }
}

namespace {

// lower bound of the float squared distance from q to any point inside the box (same expression shape as sqdist3)
__device__ __forceinline__ float box_bound(float4 lo, float4 hi, float qx, float qy, float qz) {
  float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f);
  float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f);
  float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
  return dx * dx + dy * dy + dz * dz;
}

// Exact arg-min of the float squared distance over indices [j0, j1) of `c`, restricted to d2 <= limit on entry.
// descending = false : ties -> smallest index ; true : ties -> largest index (reference visiting order).
// Returns the warp-uniform (d2, j) ; j = -1 when nothing is at distance <= limit.
__device__ __forceinline__ void range_argmin(const LastCloud& c, float qx, float qy, float qz, int j0, int j1,
                                             bool descending, float limit, float& out_d, int& out_j) {
  const unsigned lane = lane_id();
  float best_d = limit;  // per lane
  int best_j = -1;
  float wbest = limit;   // warp-uniform pruning bound
  if (j1 > j0) {
    const int t_first = j0 / ALOAM_TILE, t_last = (j1 - 1) / ALOAM_TILE;
    const float4* tlo = reinterpret_cast<const float4*>(c.tile_lo);
    const float4* thi = reinterpret_cast<const float4*>(c.tile_hi);
    const int ntl = t_last - t_first + 1;
    for (int g = 0; g < ntl; g += 32) {
      // lane k looks at the k-th tile of this group in visiting order
      int k = g + (int)lane;
      int t = descending ? (t_last - k) : (t_first + k);
      float bound = FLT_MAX;
      if (k < ntl) bound = box_bound(__ldg(tlo + t), __ldg(thi + t), qx, qy, qz);
      unsigned m = __ballot_sync(0xffffffffu, bound <= wbest);
      while (m) {
        int src = __ffs(m) - 1;
        m &= m - 1;
        float b = __shfl_sync(0xffffffffu, bound, src);
        if (b > wbest) continue;
        int tt = descending ? (t_last - (g + src)) : (t_first + g + src);
        int j = tt * ALOAM_TILE + (int)lane;
        float d2 = FLT_MAX;
        if (j >= j0 && j < j1) {
          Pt4 p = c.pts[j];
          d2 = sqdist3(p.x, p.y, p.z, qx, qy, qz);
          // a lane sees its indices in visiting order, so strict '<' keeps the first visited among equals;
          // d2 == limit on first hit is allowed in (callers re-check the strict threshold)
          if (d2 < best_d || (best_j < 0 && d2 == best_d)) { best_d = d2; best_j = j; }
        }
        unsigned mb = __reduce_min_sync(0xffffffffu, __float_as_uint(d2));
        wbest = fminf(wbest, __uint_as_float(mb));
      }
    }
  }
  // combine lanes: (d2, visiting rank)
  int rank = best_j < 0 ? INT_MAX : (descending ? (j1 - 1 - best_j) : (best_j - j0));
  float d = best_j < 0 ? FLT_MAX : best_d;
  warp_argmin(d, rank);
  out_d = d;
  out_j = (rank == INT_MAX || d == FLT_MAX) ? -1 : (descending ? (j1 - 1 - rank) : (j0 + rank));
}

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 cross3(const D3& a, const D3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// laserOdometry.cpp:111-129 with DISTORTION == 0: slerp(1, q) == q exactly; Eigen q*v = v + w*uv + u x uv, uv = 2 u x v
__device__ __forceinline__ void transform_to_start(const double* pose, float px, float py, float pz, float& ox, float& oy, float& oz) {
  const D3 u{pose[0], pose[1], pose[2]};
  const double w = pose[3];
  const D3 v{(double)px, (double)py, (double)pz};
  D3 uv = cross3(u, v);
  uv.x = uv.x + uv.x; uv.y = uv.y + uv.y; uv.z = uv.z + uv.z;
  const D3 c2 = cross3(u, uv);
  ox = (float)(((v.x + w * uv.x) + c2.x) + pose[4]);
  oy = (float)(((v.y + w * uv.y) + c2.y) + pose[5]);
  oz = (float)(((v.z + w * uv.z) + c2.z) + pose[6]);
}

__device__ __forceinline__ void store_none(BlockRec* b, int* corr) {
  b->type = -1;
  if (corr) { corr[0] = -1; corr[1] = -1; corr[2] = -1; corr[3] = 0; }
}

}  // namespace

__global__ void __launch_bounds__(256) k_odom_assoc(const Pt4* __restrict__ sharp, const Pt4* __restrict__ flat,
                                                    const int* __restrict__ feat_counts, LastCloud corner,
                                                    LastCloud surf, const double* __restrict__ pose7, OdomParams prm,
                                                    BlockRec* __restrict__ blocks, int* __restrict__ corr, int max_sharp) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = lane_id();
  const bool is_corner = wid < max_sharp;
  const int qi = is_corner ? wid : wid - max_sharp;
  const int nq = is_corner ? feat_counts[0] : feat_counts[2];
  BlockRec* out = blocks + wid;
  int* co = corr ? corr + 4 * wid : nullptr;
  if (qi >= nq) { if (lane == 0) store_none(out, co); return; }
  const LastCloud& L = is_corner ? corner : surf;
  const int n_last = *L.n;
  const Pt4 cur = is_corner ? sharp[qi] : flat[qi];
  double pose[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) pose[k] = pose7[k];
  float qx, qy, qz;
  transform_to_start(pose, cur.x, cur.y, cur.z, qx, qy, qz);
  const float thr = (float)prm.dist_sq_thresh;

  // nearest neighbour (kdtree*Last->nearestKSearch(pointSel, 1, ...), :302,390) then `< DISTANCE_SQ_THRESHOLD`
  float d1; int closest;
  range_argmin(L, qx, qy, qz, 0, n_last, false, thr, d1, closest);
  if (closest < 0 || !((double)d1 < prm.dist_sq_thresh)) { if (lane == 0) store_none(out, co); return; }
  const int rc = (int)L.pts[closest].i;                       // closestPointScanID
  // rings rc-2 .. rc+2 survive the `> rc + NEARBY_SCAN` / `< rc - NEARBY_SCAN` break tests (:319,345,405,433)
  int up = 0; while ((double)(rc + up + 1) <= (double)rc + prm.nearby_scan) ++up;
  const int r_lo = max(rc - up, 0), r_hi = min(rc + up + 1, 64);
  const int* rs = L.ring_start;
  const int fwd_other0 = rs[min(rc + 1, 64)], fwd_other1 = rs[r_hi];
  const int bwd_other0 = rs[r_lo], bwd_other1 = rs[rc];

  if (is_corner) {
    // minPointInd2: forward over higher rings (:312-335), then backward over lower rings (:338-361)
    float df, db; int jf, jb;
    range_argmin(L, qx, qy, qz, fwd_other0, fwd_other1, false, thr, df, jf);
    if (jf >= 0 && !(df < thr)) jf = -1;
    range_argmin(L, qx, qy, qz, bwd_other0, bwd_other1, true, jf >= 0 ? df : thr, db, jb);
    int second = jf;
    if (jb >= 0 && db < (jf >= 0 ? df : thr)) second = jb;
    if (second < 0) { if (lane == 0) store_none(out, co); return; }
    if (lane == 0) {
      const Pt4 a = L.pts[closest], b = L.pts[second];
      out->cp[0] = cur.x; out->cp[1] = cur.y; out->cp[2] = cur.z;
      out->a[0] = a.x; out->a[1] = a.y; out->a[2] = a.z;
      out->b[0] = b.x; out->b[1] = b.y; out->b[2] = b.z;
      const double ex = (double)a.x - (double)b.x, ey = (double)a.y - (double)b.y, ez = (double)a.z - (double)b.z;
      out->s = sqrt(ex * ex + ey * ey + ez * ez);  // de.norm(), lidarFactor.hpp:36-40
      out->type = 0;
      if (co) { co[0] = closest; co[1] = second; co[2] = -1; co[3] = 1; }
    }
  } else {
    // minPointInd2: same ring, forward part then backward part ; minPointInd3: other rings (:402-455)
    float df, db; int jf, jb;
    range_argmin(L, qx, qy, qz, closest + 1, rs[min(rc + 1, 64)], false, thr, df, jf);
    if (jf >= 0 && !(df < thr)) jf = -1;
    range_argmin(L, qx, qy, qz, rs[rc], closest, true, jf >= 0 ? df : thr, db, jb);
    int m2 = jf;
    if (jb >= 0 && db < (jf >= 0 ? df : thr)) m2 = jb;
    range_argmin(L, qx, qy, qz, fwd_other0, fwd_other1, false, thr, df, jf);
    if (jf >= 0 && !(df < thr)) jf = -1;
    range_argmin(L, qx, qy, qz, bwd_other0, bwd_other1, true, jf >= 0 ? df : thr, db, jb);
    int m3 = jf;
    if (jb >= 0 && db < (jf >= 0 ? df : thr)) m3 = jb;
    if (m2 < 0 || m3 < 0) { if (lane == 0) store_none(out, co); return; }
    if (lane == 0) {
      const Pt4 pj = L.pts[closest], pl = L.pts[m2], pm = L.pts[m3];
      out->cp[0] = cur.x; out->cp[1] = cur.y; out->cp[2] = cur.z;
      out->a[0] = pj.x; out->a[1] = pj.y; out->a[2] = pj.z;
      // ljm_norm = (j - l) x (j - m), normalised (lidarFactor.hpp:64-65)
      const D3 jl{(double)pj.x - (double)pl.x, (double)pj.y - (double)pl.y, (double)pj.z - (double)pl.z};
      const D3 jm{(double)pj.x - (double)pm.x, (double)pj.y - (double)pm.y, (double)pj.z - (double)pm.z};
      D3 nrm = cross3(jl, jm);
      const double z = nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z;
      if (z > 0) { const double nn = sqrt(z); nrm.x /= nn; nrm.y /= nn; nrm.z /= nn; }
      out->b[0] = nrm.x; out->b[1] = nrm.y; out->b[2] = nrm.z;
      out->s = 1.0;
      out->type = 1;
      if (co) { co[0] = closest; co[1] = m2; co[2] = m3; co[3] = 1; }
    }
  }
}

// exact 1-NN of arbitrary queries against a "last" cloud (aloam_knn, which = 0/1)
__global__ void __launch_bounds__(256) k_knn_last(LastCloud cloud, const Pt4* __restrict__ queries, int nq,
                                                  int* __restrict__ idx, float* __restrict__ sqd) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= nq) return;
  const Pt4 q = queries[wid];
  float d; int j;
  range_argmin(cloud, q.x, q.y, q.z, 0, *cloud.n, false, FLT_MAX, d, j);
  if (lane_id() == 0) { idx[wid] = j; sqd[wid] = j < 0 ? __int_as_float(0x7f800000) : d; }
}

}  // namespace aloam
