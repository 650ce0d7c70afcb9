// Scan-to-map association on the GPU -- replaces laserMapping.cpp:154-163 (pointAssociateToMap), :558-559 (the two
// kd-tree builds over the gathered submap), :577-622 (corner: 5-NN, covariance, eigen line test, LidarEdgeFactor)
// and :643-687 (surf: 5-NN, plane fit, LidarPlaneNormFactor).
//
// K0  map index = hash grid with cell edge 1 m * (1 + 1e-5): the reference only uses a neighbourhood whose 5th
//     member is closer than 1 m (`pointSearchSqDis[4] < 1.0`, :584,652), so every neighbour that can matter lies in
//     the 3 x 3 x 3 cells around the query and one probe round (27 lanes) replaces the O(M log M) kd-tree build +
//     descent.  Build = clear / insert (atomicCAS on the key, atomicAdd for the rank) / alloc / fill: one pass over
//     the submap.  Order inside a cell depends on atomic timing; selection is on (distance, index) so results do not.
// K5  one warp per stack point: probe 27 cells, flatten their points with a warp scan, every lane keeps its own
//     sorted top-5, five REDUX arg-min rounds merge them into the exact 5-NN in ascending (distance, index) order.
// K6  lane 0 fits the line (3x3 symmetric eigen, Jacobi) or the plane (5x3 least squares, column-pivoted Householder)
//     in double precision exactly as the oracle does and writes the residual block the LM kernel consumes.
// Multi-GPU: a rank only fits the queries whose cell it owns (slabs of 8 cells along x, round-robin over ranks); its
//     shard of the map holds those slabs plus a one-cell halo, so its 27-cell neighbourhoods are complete.
#include <cfloat>
#include <climits>
#include "ctx.h"

namespace aloam {

namespace {
constexpr unsigned long long kEmpty = ~0ull;
constexpr int kSlab = 8;  // cells per ownership slab (x direction)

__device__ __forceinline__ unsigned long long cell_key(int cx, int cy, int cz) {
  return ((unsigned long long)(unsigned)(cx + (1 << 20)) << 42) | ((unsigned long long)(unsigned)(cy + (1 << 20)) << 21) |
         (unsigned long long)(unsigned)(cz + (1 << 20));
}
__device__ __forceinline__ unsigned hash_key(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (unsigned)k;
}
__device__ __forceinline__ int cell_of(float v, float inv_cs) { return (int)floorf(v * inv_cs); }
__device__ __forceinline__ int owner_of(int cx, int count) {
  int slab = (cx + (1 << 20)) / kSlab;
  return slab % count;
}
}  // namespace

__global__ void k_grid_clear(GridTable a, GridTable b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= (int)a.mask) { a.keys[i] = kEmpty; a.cnt[i] = 0; }
  if (i <= (int)b.mask) { b.keys[i] = kEmpty; b.cnt[i] = 0; }
  if (i == 0) { *a.cursor = 0; *b.cursor = 0; }
}

// blockIdx.y selects the cloud (0 = a, 1 = b)
__global__ void k_grid_insert(GridTable a, const Pt4* __restrict__ pa, int na, GridTable b, const Pt4* __restrict__ pb, int nb) {
  const GridTable& g = blockIdx.y == 0 ? a : b;
  const Pt4* __restrict__ pts = blockIdx.y == 0 ? pa : pb;
  const int n = blockIdx.y == 0 ? na : nb;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Pt4 p = pts[i];
  const unsigned long long key = cell_key(cell_of(p.x, g.inv_cs), cell_of(p.y, g.inv_cs), cell_of(p.z, g.inv_cs));
  unsigned h = hash_key(key) & g.mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&g.keys[h], kEmpty, key);
    if (prev == kEmpty || prev == key) break;
    h = (h + 1) & g.mask;
  }
  g.slot_of[i] = (int)h;
  g.rank_of[i] = atomicAdd(&g.cnt[h], 1);
}

// per-cell storage: block-wide exclusive scan of the counts, ONE atomicAdd per CTA and table on the cursor (a per-slot
// atomicAdd serialises ~500k updates of a single address: 324 us for a 1M-point map, ncu r01)
__global__ void __launch_bounds__(256) k_grid_alloc(GridTable a, GridTable b) {
  __shared__ int s_w[2][8];
  __shared__ int s_base[2];
  const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int c[2], incl[2];
  c[0] = i <= (int)a.mask ? a.cnt[i] : 0;
  c[1] = i <= (int)b.mask ? b.cnt[i] : 0;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    int v = c[t];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
    incl[t] = v;
    if (lane == 31) s_w[t][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    int tot = 0;
    for (int k = 0; k < 8; ++k) { const int x = s_w[threadIdx.x][k]; s_w[threadIdx.x][k] = tot; tot += x; }
    s_base[threadIdx.x] = tot > 0 ? atomicAdd(threadIdx.x == 0 ? a.cursor : b.cursor, tot) : 0;
  }
  __syncthreads();
  if (i <= (int)a.mask && c[0] > 0) a.start[i] = s_base[0] + s_w[0][w] + incl[0] - c[0];
  if (i <= (int)b.mask && c[1] > 0) b.start[i] = s_base[1] + s_w[1][w] + incl[1] - c[1];
}

__global__ void k_grid_fill(GridTable a, const Pt4* __restrict__ pa, int na, GridTable b, const Pt4* __restrict__ pb, int nb) {
  const GridTable& g = blockIdx.y == 0 ? a : b;
  const Pt4* __restrict__ pts = blockIdx.y == 0 ? pa : pb;
  const int n = blockIdx.y == 0 ? na : nb;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Pt4 p = pts[i];
  g.gpts[g.start[g.slot_of[i]] + g.rank_of[i]] = make_float4(p.x, p.y, p.z, __int_as_float(i));
}

namespace {

__device__ __forceinline__ int warp_incl_scan(int v) {
  const unsigned lane = lane_id();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (unsigned)d) v += t;
  }
  return v;
}

template <int K>
struct TopK {
  float d[K]; int i[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < K; ++k) { d[k] = FLT_MAX; i[k] = INT_MAX; }
  }
  __device__ __forceinline__ void push(float dd, int ii) {   // keep ascending (d, i)
    if (!(dd < d[K - 1] || (dd == d[K - 1] && ii < i[K - 1]))) return;
    d[K - 1] = dd; i[K - 1] = ii;
#pragma unroll
    for (int k = K - 1; k > 0; --k) {
      const bool sw = d[k] < d[k - 1] || (d[k] == d[k - 1] && i[k] < i[k - 1]);
      if (sw) { float td = d[k]; d[k] = d[k - 1]; d[k - 1] = td; int ti = i[k]; i[k] = i[k - 1]; i[k - 1] = ti; }
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int k = 0; k < K - 1; ++k) { d[k] = d[k + 1]; i[k] = i[k + 1]; }
    d[K - 1] = FLT_MAX; i[K - 1] = INT_MAX;
  }
};

// all points of the 27 cells around q go through f(x, y, z, index), each exactly once, on some lane
template <typename F>
__device__ __forceinline__ void visit_block27(const GridTable& g, float qx, float qy, float qz, F&& f) {
  const int lane = (int)lane_id();
  const int cx = cell_of(qx, g.inv_cs), cy = cell_of(qy, g.inv_cs), cz = cell_of(qz, g.inv_cs);
  int start = 0, cnt = 0;
  if (lane < 27) {
    const int dz = lane / 9 - 1, dy = (lane % 9) / 3 - 1, dx = lane % 3 - 1;
    const unsigned long long key = cell_key(cx + dx, cy + dy, cz + dz);
    unsigned h = hash_key(key) & g.mask;
    for (;;) {
      const unsigned long long k = g.keys[h];
      if (k == key) { start = g.start[h]; cnt = g.cnt[h]; break; }
      if (k == kEmpty) break;
      h = (h + 1) & g.mask;
    }
  }
  const int incl = warp_incl_scan(cnt);
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  for (int e0 = 0; e0 < total; e0 += 32) {
    const int e = e0 + lane;
    int c = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const int v = __shfl_sync(0xffffffffu, incl, c + step - 1);
      if (v <= e) c += step;
    }
    c = min(c, 31);
    const int c_incl = __shfl_sync(0xffffffffu, incl, c);
    const int c_cnt = __shfl_sync(0xffffffffu, cnt, c);
    const int c_start = __shfl_sync(0xffffffffu, start, c);
    if (e < total) {
      const float4 p = __ldg(g.gpts + c_start + (e - (c_incl - c_cnt)));
      f(p.x, p.y, p.z, __float_as_int(p.w));
    }
  }
}

// merges the lanes' sorted lists: out[k] = k-th smallest (d, i) of the warp, warp-uniform ; returns how many exist
template <int K>
__device__ __forceinline__ int warp_merge_topk(TopK<K>& t, float* out_d, int* out_i) {
  int found = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float d = t.d[0]; int i = t.i[0];
    warp_argmin(d, i);
    out_d[k] = d; out_i[k] = i;
    if (i != INT_MAX && d != FLT_MAX) { ++found; if (t.i[0] == i && t.d[0] == d) t.pop(); }
  }
  return found;
}

// 3x3 symmetric eigen-decomposition, cyclic Jacobi (same rotations as the oracle's eig3_sym): ascending eigenvalues,
// vmax = unit eigenvector of the largest one.  Fully unrolled: every array index is a compile-time constant, so the
// matrices live in registers.
__device__ __forceinline__ void eig3_sym(const double (&Ain)[9], double (&evals)[3], double (&vmax)[3]) {
  double A[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { A[i] = Ain[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    const double dsum = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
    if (off <= 1e-32 * dsum || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[3 * p + q];
        if (apq != 0.0) {
          const double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
          const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
          for (int k = 0; k < 3; ++k) { const double akp = A[3 * k + p], akq = A[3 * k + q]; A[3 * k + p] = c * akp - s * akq; A[3 * k + q] = s * akp + c * akq; }
#pragma unroll
          for (int k = 0; k < 3; ++k) { const double apk = A[3 * p + k], aqk = A[3 * q + k]; A[3 * p + k] = c * apk - s * aqk; A[3 * q + k] = s * apk + c * aqk; }
#pragma unroll
          for (int k = 0; k < 3; ++k) { const double vkp = V[3 * k + p], vkq = V[3 * k + q]; V[3 * k + p] = c * vkp - s * vkq; V[3 * k + q] = s * vkp + c * vkq; }
        }
      }
    }
  }
  // sort the three (eigenvalue, eigenvector) pairs ascending with static compare-exchanges
  double e0 = A[0], e1 = A[4], e2 = A[8];
  double v0[3] = {V[0], V[3], V[6]}, v1[3] = {V[1], V[4], V[7]}, v2[3] = {V[2], V[5], V[8]};
#define ALOAM_CSWAP(ea, va, eb, vb) if (eb < ea) { double te = ea; ea = eb; eb = te; for (int k = 0; k < 3; ++k) { double tv = va[k]; va[k] = vb[k]; vb[k] = tv; } }
  ALOAM_CSWAP(e0, v0, e1, v1)
  ALOAM_CSWAP(e1, v1, e2, v2)
  ALOAM_CSWAP(e0, v0, e1, v1)
#undef ALOAM_CSWAP
  evals[0] = e0; evals[1] = e1; evals[2] = e2;
#pragma unroll
  for (int k = 0; k < 3; ++k) vmax[k] = v2[k];
}

// least squares A n = b (5x3) by Householder QR, fully unrolled (registers).  The oracle pivots columns (Eigen's
// colPivHouseholderQr); for a full-rank 5x3 system both give the least-squares solution to rounding (compared at 1e-10).
__device__ __forceinline__ void lsq_5x3(const double (&Ain)[15], const double (&bin)[5], double (&n)[3]) {
  double A[15], b[5];
#pragma unroll
  for (int i = 0; i < 15; ++i) A[i] = Ain[i];
#pragma unroll
  for (int i = 0; i < 5; ++i) b[i] = bin[i];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double norm2 = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) if (i >= k) norm2 += A[3 * i + k] * A[3 * i + k];
    const double nrm = sqrt(norm2);
    const double akk = A[3 * k + k];
    const double alpha = akk > 0 ? -nrm : nrm;
    const double v0 = akk - alpha;
    const double vn2 = norm2 - akk * akk + v0 * v0;
    A[3 * k + k] = v0;
    if (vn2 > 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j > k) {
          double s = 0;
#pragma unroll
          for (int i = 0; i < 5; ++i) if (i >= k) s += A[3 * i + k] * A[3 * i + j];
          s = 2.0 * s / vn2;
#pragma unroll
          for (int i = 0; i < 5; ++i) if (i >= k) A[3 * i + j] -= s * A[3 * i + k];
        }
      }
      double s = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) if (i >= k) s += A[3 * i + k] * b[i];
      s = 2.0 * s / vn2;
#pragma unroll
      for (int i = 0; i < 5; ++i) if (i >= k) b[i] -= s * A[3 * i + k];
    }
    A[3 * k + k] = alpha;
  }
  n[2] = b[2] / A[8];
  n[1] = (b[1] - A[5] * n[2]) / A[4];
  n[0] = (b[0] - A[1] * n[1] - A[2] * n[2]) / A[0];
}

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 cross3(const D3& a, const D3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

}  // namespace

__global__ void __launch_bounds__(256) k_map_assoc(const Pt4* __restrict__ corner_stack, int n_corner,
                                                   const Pt4* __restrict__ surf_stack, int n_surf, MapCloud corner_map,
                                                   MapCloud surf_map, const double* __restrict__ pose7,
                                                   BlockRec* __restrict__ blocks, double* __restrict__ fits, int shard_rank,
                                                   int shard_count) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = lane_id();
  if (wid >= n_corner + n_surf) return;
  const bool is_corner = wid < n_corner;
  const Pt4 ori = is_corner ? corner_stack[wid] : surf_stack[wid - n_corner];
  const MapCloud& M = is_corner ? corner_map : surf_map;
  BlockRec* out = blocks + wid;
  double* fo = fits ? fits + (size_t)wid * 14 : nullptr;
  // pointAssociateToMap (:154-163): q * p + t in double, stored to float
  float qx, qy, qz;
  {
    const D3 u{pose7[0], pose7[1], pose7[2]};
    const double w = pose7[3];
    const D3 v{(double)ori.x, (double)ori.y, (double)ori.z};
    D3 uv = cross3(u, v);
    uv.x = uv.x + uv.x; uv.y = uv.y + uv.y; uv.z = uv.z + uv.z;
    const D3 c2 = cross3(u, uv);
    qx = (float)(((v.x + w * uv.x) + c2.x) + pose7[4]);
    qy = (float)(((v.y + w * uv.y) + c2.y) + pose7[5]);
    qz = (float)(((v.z + w * uv.z) + c2.z) + pose7[6]);
  }
  bool reject = false;
  if (shard_count > 1 && owner_of(cell_of(qx, M.grid.inv_cs), shard_count) != shard_rank) reject = true;
  float nd[5]; int ni[5];
  if (!reject) {
    TopK<5> t; t.init();
    visit_block27(M.grid, qx, qy, qz, [&](float x, float y, float z, int idx) { t.push(sqdist3(x, y, z, qx, qy, qz), idx); });
    const int found = warp_merge_topk<5>(t, nd, ni);
    reject = !(found == 5 && (double)nd[4] < 1.0);   // pointSearchSqDis[4] < 1.0 (:584,652)
  }
  if (reject) {
    if (lane == 0) { out->type = -1; if (fo) fo[1] = -1.0; }
    return;
  }
  if (lane != 0) return;
  double P[15];
  for (int j = 0; j < 5; ++j) { const Pt4 p = M.pts[ni[j]]; P[3 * j] = p.x; P[3 * j + 1] = p.y; P[3 * j + 2] = p.z; }
  out->cp[0] = ori.x; out->cp[1] = ori.y; out->cp[2] = ori.z;
  if (fo) { fo[0] = is_corner ? wid : wid - n_corner; for (int j = 0; j < 5; ++j) fo[9 + j] = ni[j]; }
  if (is_corner) {
    // :586-616  centre, scatter matrix, eigen-decomposition, line test
    double c[3] = {0, 0, 0};
    for (int j = 0; j < 5; ++j) { c[0] = c[0] + P[3 * j]; c[1] = c[1] + P[3 * j + 1]; c[2] = c[2] + P[3 * j + 2]; }
    c[0] = c[0] / 5.0; c[1] = c[1] / 5.0; c[2] = c[2] / 5.0;
    double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 5; ++j) {
      const double d[3] = {P[3 * j] - c[0], P[3 * j + 1] - c[1], P[3 * j + 2] - c[2]};
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) cov[3 * a + b] += d[a] * d[b];
    }
    double ev[3], dir[3];
    eig3_sym(cov, ev, dir);
    if (!(ev[2] > 3 * ev[1])) { out->type = -1; if (fo) fo[1] = -1.0; return; }
    for (int k = 0; k < 3; ++k) { out->a[k] = 0.1 * dir[k] + c[k]; out->b[k] = -0.1 * dir[k] + c[k]; }
    const double ex = out->a[0] - out->b[0], ey = out->a[1] - out->b[1], ez = out->a[2] - out->b[2];
    out->s = 1.0 / sqrt(ex * ex + ey * ey + ez * ez);
    out->type = 0;
    if (fo) { fo[1] = 0.0; for (int k = 0; k < 3; ++k) { fo[2 + k] = out->a[k]; fo[5 + k] = out->b[k]; } fo[8] = 0.0; }
  } else {
    // :650-684  plane A n = -1, normalise, fit check, LidarPlaneNormFactor
    const double rhs[5] = {-1, -1, -1, -1, -1};
    double nv[3];
    lsq_5x3(P, rhs, nv);
    const double nn = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
    const double neg_oa = 1 / nn;
    nv[0] /= nn; nv[1] /= nn; nv[2] /= nn;
    bool valid = true;
    for (int j = 0; j < 5; ++j)
      if (fabs(nv[0] * P[3 * j] + nv[1] * P[3 * j + 1] + nv[2] * P[3 * j + 2] + neg_oa) > 0.2) { valid = false; break; }
    if (!valid) { out->type = -1; if (fo) fo[1] = -1.0; return; }
    for (int k = 0; k < 3; ++k) { out->a[k] = nv[k]; out->b[k] = 0.0; }
    out->s = neg_oa;
    out->type = 2;
    if (fo) { fo[1] = 2.0; for (int k = 0; k < 3; ++k) { fo[2 + k] = nv[k]; fo[5 + k] = 0.0; } fo[8] = neg_oa; }
  }
}

// exact k-NN (k <= 8) against a map cloud: 27-cell block first; if its k-th distance is not inside the guaranteed
// radius (one cell), a coalesced sweep over the whole cloud settles it
__global__ void __launch_bounds__(256) k_map_knn(MapCloud map, const Pt4* __restrict__ queries, int nq, int k,
                                                 int* __restrict__ idx, float* __restrict__ sqd) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= nq) return;
  const Pt4 q = queries[wid];
  TopK<8> t; t.init();
  visit_block27(map.grid, q.x, q.y, q.z, [&](float x, float y, float z, int i) { t.push(sqdist3(x, y, z, q.x, q.y, q.z), i); });
  float od[8]; int oi[8];
  TopK<8> t2 = t;
  warp_merge_topk<8>(t2, od, oi);
  const float safe = map.grid.cs * map.grid.cs * 0.9999f;
  if (!(oi[k - 1] != INT_MAX && od[k - 1] < safe)) {
    t.init();
    for (int i = (int)lane_id(); i < map.n; i += 32) {
      const Pt4 p = map.pts[i];
      t.push(sqdist3(p.x, p.y, p.z, q.x, q.y, q.z), i);
    }
    warp_merge_topk<8>(t, od, oi);
  }
  if (lane_id() == 0)
    for (int j = 0; j < k; ++j) {
      idx[(size_t)wid * k + j] = oi[j] == INT_MAX ? -1 : oi[j];
      sqd[(size_t)wid * k + j] = oi[j] == INT_MAX ? __int_as_float(0x7f800000) : od[j];
    }
}

}  // namespace aloam

// ---------------------------------------------------------------------------------------------------------------
// host side of the mapping entry points
using namespace aloam;

namespace {
int ensure_map_buffers(aloam_ctx* c) {
  if (c->map_corner.pts) return ALOAM_OK;
  if (c->cfg.max_map_points <= 0) return ALOAM_ERR_CAPACITY;
  c->max_map = c->cfg.max_map_points;
  c->map_slots = 1024;
  while (c->map_slots < 2 * c->max_map) c->map_slots <<= 1;
  for (MapCloud* m : {&c->map_corner, &c->map_surf}) {
    GridTable& g = m->grid;
    const size_t mm = (size_t)c->max_map;
    if (cudaMalloc((void**)&m->pts, mm * 16) != cudaSuccess || cudaMalloc((void**)&g.keys, (size_t)c->map_slots * 8) != cudaSuccess ||
        cudaMalloc((void**)&g.cnt, (size_t)c->map_slots * 4) != cudaSuccess || cudaMalloc((void**)&g.start, (size_t)c->map_slots * 4) != cudaSuccess ||
        cudaMalloc((void**)&g.cursor, 16) != cudaSuccess || cudaMalloc((void**)&g.slot_of, mm * 4) != cudaSuccess ||
        cudaMalloc((void**)&g.rank_of, mm * 4) != cudaSuccess || cudaMalloc((void**)&g.gpts, mm * 16) != cudaSuccess)
      return ALOAM_ERR_CUDA;
    g.mask = (unsigned)c->map_slots - 1;
    g.cs = 1.0f * (1.0f + 1e-5f);   // > 1 m so that an f32 d^2 < 1.0 neighbour can never sit outside the 27-cell block
    g.inv_cs = 1.0f / g.cs;
    m->n = 0;
  }
  const size_t mp = (size_t)c->max_points;
  if (cudaMalloc((void**)&c->d_stack_corner, mp * 16) != cudaSuccess || cudaMalloc((void**)&c->d_stack_surf, mp * 16) != cudaSuccess ||
      cudaMalloc((void**)&c->d_fits, 2 * mp * 14 * 8) != cudaSuccess || cudaMalloc((void**)&c->d_map_blocks, 2 * mp * sizeof(BlockRec)) != cudaSuccess)
    return ALOAM_ERR_CUDA;
  return ALOAM_OK;
}
}  // namespace

extern "C" {

void aloam_map_free_impl(aloam_ctx* c) {
  for (MapCloud* m : {&c->map_corner, &c->map_surf}) {
    void* ps[] = {m->pts, m->grid.keys, m->grid.cnt, m->grid.start, m->grid.cursor, m->grid.slot_of, m->grid.rank_of, m->grid.gpts};
    for (void* p : ps) if (p) cudaFree(p);
  }
  void* qs[] = {c->d_stack_corner, c->d_stack_surf, c->d_fits, c->d_map_blocks};
  for (void* p : qs) if (p) cudaFree(p);
}

int aloam_map_upload_impl(aloam_ctx* c, aloam_cloud_view corner_map, aloam_cloud_view surf_map) {
  if (!c) return ALOAM_ERR_INVALID_ARG;
  int rc = check_view(corner_map); if (rc) return rc;
  rc = check_view(surf_map); if (rc) return rc;
  CUDA_CHECK_RET(cudaSetDevice(c->cfg.device));
  rc = ensure_map_buffers(c); if (rc) return rc;
  if (corner_map.n > c->max_map || surf_map.n > c->max_map) return ALOAM_ERR_CAPACITY;
  rc = upload_cloud(c, corner_map, c->map_corner.pts, c->max_map); if (rc) return rc;
  rc = upload_cloud(c, surf_map, c->map_surf.pts, c->max_map); if (rc) return rc;
  c->map_corner.n = corner_map.n; c->map_surf.n = surf_map.n;
  const int tb = (c->map_slots + 255) / 256;
  const int pb = (std::max(std::max(corner_map.n, surf_map.n), 1) + 255) / 256;
  LAUNCH(c, KID_MAP_GRID, k_grid_clear, tb, 256, 0, c->map_corner.grid, c->map_surf.grid);
  LAUNCH(c, KID_MAP_GRID, k_grid_insert, dim3(pb, 2), 256, 0, c->map_corner.grid, c->map_corner.pts, corner_map.n, c->map_surf.grid,
         c->map_surf.pts, surf_map.n);
  LAUNCH(c, KID_MAP_GRID, k_grid_alloc, tb, 256, 0, c->map_corner.grid, c->map_surf.grid);
  LAUNCH(c, KID_MAP_GRID, k_grid_fill, dim3(pb, 2), 256, 0, c->map_corner.grid, c->map_corner.pts, corner_map.n, c->map_surf.grid,
         c->map_surf.pts, surf_map.n);
  CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
  CUDA_CHECK_RET(cudaGetLastError());
  prof_collect(c);
  c->have_map = true;
  return ALOAM_OK;
}

int aloam_mapping_register_impl(aloam_ctx* c, aloam_cloud_view corner_stack, aloam_cloud_view surf_stack, double x[7], aloam_stats* stats) {
  if (!c || !x) return ALOAM_ERR_INVALID_ARG;
  if (!c->have_map) return ALOAM_ERR_STATE;
  int rc = check_view(corner_stack); if (rc) return rc;
  rc = check_view(surf_stack); if (rc) return rc;
  CUDA_CHECK_RET(cudaSetDevice(c->cfg.device));
  if (!(c->map_corner.n > 10 && c->map_surf.n > 50)) {   // laserMapping.cpp:554,730-733: pose unchanged
    if (stats) { std::memset(stats, 0, sizeof(*stats)); stats->flags = ALOAM_FLAG_MAP_TOO_THIN; }
    return ALOAM_OK;
  }
  CUDA_CHECK_RET(cudaEventRecord(c->ev0, c->stream));
  rc = upload_cloud(c, corner_stack, c->d_stack_corner, c->max_points); if (rc) return rc;
  rc = upload_cloud(c, surf_stack, c->d_stack_surf, c->max_points); if (rc) return rc;
  for (int k = 0; k < 7; ++k) c->h_dbl[k] = x[k];
  CUDA_CHECK_RET(cudaMemcpyAsync(c->d_pose, c->h_dbl, 56, cudaMemcpyHostToDevice, c->stream));
  const int nq = corner_stack.n + surf_stack.n;
  const LmParams lp = lm_params(c->cfg);
  for (int it = 0; it < c->cfg.outer_iters; ++it) {
    if (nq > 0)
      LAUNCH(c, KID_MAP_KNN_FIT, k_map_assoc, (nq + 7) / 8, 256, 0, c->d_stack_corner, corner_stack.n, c->d_stack_surf, surf_stack.n,
             c->map_corner, c->map_surf, c->d_pose, c->d_map_blocks, c->d_fits, c->shard_rank, c->shard_count);
    launch_lm_step(c, c->d_map_blocks, nq, c->d_pose, lp, c->d_summary + (it & 3));
  }
  CUDA_CHECK_RET(cudaMemcpyAsync(c->h_dbl + 8, c->d_pose, 56, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK_RET(cudaMemcpyAsync(c->h_summary, c->d_summary, sizeof(LmSummary) * 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK_RET(cudaEventRecord(c->ev1, c->stream));
  CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
  CUDA_CHECK_RET(cudaGetLastError());
  prof_collect(c);
  for (int k = 0; k < 7; ++k) x[k] = c->h_dbl[8 + k];
  float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  fill_stats(c, stats, c->cfg.outer_iters, 0, ms);
  return ALOAM_OK;
}

// association only (tests): fits = (n_corner + n_surf) x 14 doubles [query, type (-1 rejected), p0(3), p1(3), d, nn(5)]
int aloam_mapping_associate(aloam_ctx* c, aloam_cloud_view corner_stack, aloam_cloud_view surf_stack, const double x[7], double* fits) {
  if (!c || !x || !fits) return ALOAM_ERR_INVALID_ARG;
  if (!c->have_map) return ALOAM_ERR_STATE;
  int rc = check_view(corner_stack); if (rc) return rc;
  rc = check_view(surf_stack); if (rc) return rc;
  CUDA_CHECK_RET(cudaSetDevice(c->cfg.device));
  rc = upload_cloud(c, corner_stack, c->d_stack_corner, c->max_points); if (rc) return rc;
  rc = upload_cloud(c, surf_stack, c->d_stack_surf, c->max_points); if (rc) return rc;
  for (int k = 0; k < 7; ++k) c->h_dbl[k] = x[k];
  CUDA_CHECK_RET(cudaMemcpyAsync(c->d_pose, c->h_dbl, 56, cudaMemcpyHostToDevice, c->stream));
  const int nq = corner_stack.n + surf_stack.n;
  if (nq > 0) {
    LAUNCH(c, KID_MAP_KNN_FIT, k_map_assoc, (nq + 7) / 8, 256, 0, c->d_stack_corner, corner_stack.n, c->d_stack_surf, surf_stack.n,
           c->map_corner, c->map_surf, c->d_pose, c->d_map_blocks, c->d_fits, c->shard_rank, c->shard_count);
    CUDA_CHECK_RET(cudaMemcpyAsync(fits, c->d_fits, (size_t)nq * 14 * 8, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
  CUDA_CHECK_RET(cudaGetLastError());
  return ALOAM_OK;
}

int aloam_map_knn_impl(aloam_ctx* c, int which, aloam_cloud_view queries, int k, int* idx, float* sqdist) {
  if (!c->have_map) return ALOAM_ERR_STATE;
  if (k < 1 || k > 8) return ALOAM_ERR_INVALID_ARG;
  int rc = upload_cloud(c, queries, c->d_query, c->max_points); if (rc) return rc;
  if (queries.n > 0) {
    if ((size_t)queries.n * k > (size_t)c->max_points) return ALOAM_ERR_CAPACITY;
    LAUNCH(c, KID_MAP_KNN, k_map_knn, (queries.n + 7) / 8, 256, 0, which == 2 ? c->map_corner : c->map_surf, c->d_query, queries.n, k,
           c->d_knn_idx, c->d_knn_d);
    CUDA_CHECK_RET(cudaMemcpyAsync(idx, c->d_knn_idx, (size_t)queries.n * k * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK_RET(cudaMemcpyAsync(sqdist, c->d_knn_d, (size_t)queries.n * k * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
  CUDA_CHECK_RET(cudaGetLastError());
  return ALOAM_OK;
}

int aloam_voxel_filter_impl(aloam_ctx* c, aloam_cloud_view in, float leaf, aloam_cloud_view* out);

}  // extern "C"
