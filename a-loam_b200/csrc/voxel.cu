// pcl::VoxelGrid<PointXYZI>::applyFilter on an arbitrary cloud -- the scan-stack filters of laserMapping.cpp:543-549
// (downSizeFilterCorner / downSizeFilterSurf) as a general GPU routine (the per-ring 0.2 m instance of
// scanRegistration.cpp:401-407 is fused into k_ring_features).  Semantics per SURVEY.md 8a "V":
//   bounding box -> min_b / div_b -> voxel index (float multiply, floor) -> sort by index -> one output per occupied
//   voxel in ascending index order = float32 centroid of x, y, z, intensity accumulated in sorted order.
// PCL's std::sort is unstable; here ties are resolved by point index (a STABLE least-significant-digit radix sort on
// the voxel index, 8 bits per pass, as many passes as the index range needs), which is the CANONICAL order of the oracle.
#include <climits>
#include <cfloat>
#include <cooperative_groups.h>
#include "ctx.h"

namespace aloam {

namespace {
constexpr int VT = 256, VCH = 1024, VIT = VCH / VT;

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : (i ^ 0x7fffffff); }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : (i ^ 0x7fffffff)); }
}  // namespace

// mm6: ordered-int encoded [min x,y,z | max x,y,z], pre-set to +inf / -inf encodings
__global__ void k_vox_bbox(const Pt4* __restrict__ pts, int n, int* __restrict__ mm6) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Pt4 p = pts[i];
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm6[a], f2ord(lo[a])); atomicMax(&mm6[3 + a], f2ord(hi[a])); }
  }
}

struct VoxParams { float inv; int min_b[3]; int div0, div01; };

__global__ void k_vox_keys(const Pt4* __restrict__ pts, int n, VoxParams vp, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Pt4 p = pts[i];
  const int i0 = (int)(floorf(p.x * vp.inv) - (float)vp.min_b[0]);
  const int i1 = (int)(floorf(p.y * vp.inv) - (float)vp.min_b[1]);
  const int i2 = (int)(floorf(p.z * vp.inv) - (float)vp.min_b[2]);
  keys[i] = (unsigned)(i0 + i1 * vp.div0 + i2 * vp.div01);
  vals[i] = i;
}

__global__ void __launch_bounds__(VT) k_radix_hist(const unsigned* __restrict__ keys, const int* __restrict__ n_ptr, int shift, int* __restrict__ hist) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  __shared__ int s_h[256];
  const int tid = threadIdx.x;
  const int n = *n_ptr;
  if (blockIdx.x * VCH >= n) return;   // grids are sized by a host upper bound; the live size is on the device
  s_h[tid] = 0;
  __syncthreads();
#pragma unroll
  for (int it = 0; it < VIT; ++it) {
    const int i = blockIdx.x * VCH + it * VT + tid;
    if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[blockIdx.x * 256 + tid] = s_h[tid];
}

// one CTA, 1024 threads: thread (bin = t >> 2, quarter = t & 3) ; offsets are digit-major over all blocks
__global__ void __launch_bounds__(1024) k_radix_scan(const int* __restrict__ hist, const int* __restrict__ n_ptr, int* __restrict__ offsets) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  const int nblocks = (*n_ptr + VCH - 1) / VCH;
  __shared__ int s_tot[256];
  __shared__ int s_start[257];
  const int t = threadIdx.x, bin = t >> 2, q = t & 3;
  const int per = (nblocks + 3) / 4;
  const int b0 = min(nblocks, q * per), b1 = min(nblocks, b0 + per);
  int local = 0;
  for (int b = b0; b < b1; ++b) local += hist[b * 256 + bin];
  int incl = local;
#pragma unroll
  for (int d = 1; d < 4; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d, 4); if (q >= d) incl += v; }
  if (q == 3) s_tot[bin] = incl;
  __syncthreads();
  if (t < 32) {  // exclusive scan of 256 bin totals, 8 per lane
    int v[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = s_tot[t * 8 + k]; sum += v[k]; }
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (t >= d) inc += u; }
    int run = inc - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s_start[t * 8 + k] = run; run += v[k]; }
  }
  __syncthreads();
  int running = s_start[bin] + (incl - local);
  for (int b = b0; b < b1; ++b) { offsets[b * 256 + bin] = running; running += hist[b * 256 + bin]; }
}

__global__ void __launch_bounds__(VT) k_radix_scatter(const unsigned* __restrict__ keys_in, const int* __restrict__ vals_in, const int* __restrict__ n_ptr,
                                                      int shift, const int* __restrict__ offsets, unsigned* __restrict__ keys_out,
                                                      int* __restrict__ vals_out) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  __shared__ int s_cnt[VIT * (VT / 32)][256];   // [slot = it*8 + warp][digit] -> exclusive prefix over slots
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int n = *n_ptr;
  if (blockIdx.x * VCH >= n) return;
  for (int k = tid; k < VIT * (VT / 32) * 256; k += VT) (&s_cnt[0][0])[k] = 0;
  __syncthreads();
  int dig[VIT], rank[VIT];
#pragma unroll
  for (int it = 0; it < VIT; ++it) {
    const int i = blockIdx.x * VCH + it * VT + tid;
    const int d = i < n ? (int)((keys_in[i] >> shift) & 255u) : -1;
    dig[it] = d;
    const unsigned grp = __match_any_sync(0xffffffffu, d);
    rank[it] = __popc(grp & ((1u << lane) - 1u));
    if (d >= 0 && rank[it] == 0) s_cnt[it * (VT / 32) + w][d] = __popc(grp);
  }
  __syncthreads();
  {
    int run = offsets[blockIdx.x * 256 + tid];
#pragma unroll
    for (int k = 0; k < VIT * (VT / 32); ++k) { const int c = s_cnt[k][tid]; s_cnt[k][tid] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < VIT; ++it) {
    if (dig[it] < 0) continue;
    const int i = blockIdx.x * VCH + it * VT + tid;
    const int dst = s_cnt[it * (VT / 32) + w][dig[it]] + rank[it];
    keys_out[dst] = keys_in[i];
    vals_out[dst] = vals_in[i];
  }
}

// All passes of the LSD radix sort in ONE cooperative launch (the segmented filters of the mapper sort a few 10^4 keys: twelve
// launches of a few microseconds each were launch latency, not work).  Tiles of VCH keys go round-robin over the CTAs; per pass:
// tile histograms -> grid barrier -> every CTA derives the offsets of its own tiles from the histogram table (thread = digit:
// digit totals over all tiles, block-wide exclusive scan, plus the running sum over the tiles before its own) and scatters ->
// grid barrier.  Same stable order as k_radix_hist / k_radix_scan / k_radix_scatter; the result is in buffer (passes & 1).
__global__ void __launch_bounds__(VT) k_radix_sort_all(unsigned* __restrict__ k0, int* __restrict__ v0, unsigned* __restrict__ k1, int* __restrict__ v1,
                                                       const int* __restrict__ n_ptr, int bits, int* __restrict__ hist) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ int s_h[256];
  __shared__ int s_w[VT / 32];
  __shared__ int s_cnt[VIT * (VT / 32)][256];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int n = *n_ptr;
  const int tiles = (n + VCH - 1) / VCH;
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {   // `bits` is a launch parameter: every CTA passes the same barriers
    const unsigned* __restrict__ kin = cur ? k1 : k0;
    const int* __restrict__ vin = cur ? v1 : v0;
    unsigned* __restrict__ kout = cur ? k0 : k1;
    int* __restrict__ vout = cur ? v0 : v1;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      s_h[tid] = 0;
      __syncthreads();
#pragma unroll
      for (int it = 0; it < VIT; ++it) {
        const int i = t * VCH + it * VT + tid;
        if (i < n) atomicAdd(&s_h[(kin[i] >> shift) & 255u], 1);
      }
      __syncthreads();
      hist[t * 256 + tid] = s_h[tid];
      __syncthreads();
    }
    grid.sync();
    // thread = digit: start of the digit in the output = exclusive scan of the digit totals
    int tot = 0, pre = 0, pre_upto = 0;   // pre = sum of hist[t'][digit] over the tiles t' < pre_upto
    for (int t = 0; t < tiles; ++t) tot += hist[t * 256 + tid];
    int incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    int wb = 0;
    for (int k = 0; k < w; ++k) wb += s_w[k];
    const int start = wb + incl - tot;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      for (; pre_upto < t; ++pre_upto) pre += hist[pre_upto * 256 + tid];
      for (int k = tid; k < VIT * (VT / 32) * 256; k += VT) (&s_cnt[0][0])[k] = 0;
      __syncthreads();
      int dig[VIT], rank[VIT];
#pragma unroll
      for (int it = 0; it < VIT; ++it) {
        const int i = t * VCH + it * VT + tid;
        const int d = i < n ? (int)((kin[i] >> shift) & 255u) : -1;
        dig[it] = d;
        const unsigned grp = __match_any_sync(0xffffffffu, d);
        rank[it] = __popc(grp & ((1u << lane) - 1u));
        if (d >= 0 && rank[it] == 0) s_cnt[it * (VT / 32) + w][d] = __popc(grp);
      }
      __syncthreads();
      {
        int run = start + pre;
#pragma unroll
        for (int k = 0; k < VIT * (VT / 32); ++k) { const int c = s_cnt[k][tid]; s_cnt[k][tid] = run; run += c; }
      }
      __syncthreads();
#pragma unroll
      for (int it = 0; it < VIT; ++it) {
        if (dig[it] < 0) continue;
        const int i = t * VCH + it * VT + tid;
        const int dst = s_cnt[it * (VT / 32) + w][dig[it]] + rank[it];
        kout[dst] = kin[i];
        vout[dst] = vin[i];
      }
      __syncthreads();
    }
    grid.sync();
    cur ^= 1;
  }
}

// number of voxel heads (first element of each run of equal keys) per block of VCH sorted elements
__global__ void __launch_bounds__(VT) k_vox_heads(const unsigned* __restrict__ keys, const int* __restrict__ n_ptr, int* __restrict__ block_heads) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  const int n = *n_ptr;
  if (blockIdx.x * VCH >= n) return;
  __shared__ int s_c;
  if (threadIdx.x == 0) s_c = 0;
  __syncthreads();
  int c = 0;
#pragma unroll
  for (int it = 0; it < VIT; ++it) {
    const int i = blockIdx.x * VCH + it * VT + threadIdx.x;
    if (i < n && (i == 0 || keys[i] != keys[i - 1])) ++c;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_c, c);
  __syncthreads();
  if (threadIdx.x == 0) block_heads[blockIdx.x] = s_c;
}

__global__ void __launch_bounds__(1024) k_vox_blockscan(int* __restrict__ block_heads, const int* __restrict__ n_ptr, int* __restrict__ total) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  const int nblocks = (*n_ptr + VCH - 1) / VCH;
  __shared__ int s_w[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int per = (nblocks + 1023) / 1024;
  int sum = 0;
  for (int k = 0; k < per; ++k) { const int b = t * per + k; if (b < nblocks) sum += block_heads[b]; }
  int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  if (w == 0) {
    int x = s_w[lane], inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
    s_w[lane] = inc - x;
    if (lane == 31) *total = inc;
  }
  __syncthreads();
  int run = s_w[w] + incl - sum;
  for (int k = 0; k < per; ++k) { const int b = t * per + k; if (b < nblocks) { const int c = block_heads[b]; block_heads[b] = run; run += c; } }
}

// every head sums its run in sorted order (float, like pcl::CentroidPoint) and writes voxel number (block offset + local rank)
__global__ void __launch_bounds__(VT) k_vox_emit(const Pt4* __restrict__ pts, const unsigned* __restrict__ keys, const int* __restrict__ vals,
                                                 const int* __restrict__ n_ptr, const int* __restrict__ block_offsets, Pt4* __restrict__ out) {
  const int n = *n_ptr;
  if (blockIdx.x * VCH >= n) return;
  __shared__ int s_w[VT / 32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int base = blockIdx.x * VCH + tid * VIT;   // this thread owns VIT consecutive sorted slots
  int heads = 0;
  bool is_head[VIT];
#pragma unroll
  for (int k = 0; k < VIT; ++k) {
    const int i = base + k;
    is_head[k] = i < n && (i == 0 || keys[i] != keys[i - 1]);
    heads += is_head[k] ? 1 : 0;
  }
  int incl = heads;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  int wbase = 0;
  for (int v = 0; v < w; ++v) wbase += s_w[v];
  int slot = block_offsets[blockIdx.x] + wbase + incl - heads;
#pragma unroll
  for (int k = 0; k < VIT; ++k) {
    if (!is_head[k]) continue;
    const int i = base + k;
    const unsigned key = keys[i];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int cnt = 0;
    for (int j = i; j < n && keys[j] == key; ++j) {
      const Pt4 p = pts[vals[j]];
      sx += p.x; sy += p.y; sz += p.z; si += p.i;
      ++cnt;
    }
    const float nf = (float)cnt;
    Pt4 o; o.x = sx / nf; o.y = sy / nf; o.z = sz / nf; o.i = si / nf;
    out[slot++] = o;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Segmented, device-resident form: S independent clouds ("segments", each with its own leaf size) are filtered by ONE
// sequence of launches -- the two scan stacks of a mapping frame (laserMapping.cpp:543-549), or all valid cubes of the map
// after insertion (:787-801, up to 75 corner + 75 surf cubes).  Sizes live in device memory; nothing here synchronises
// with the host.  Key = segment << idx_bits | voxel index (PCL's index inside the segment's own bounding box), so one
// stable LSD radix sort orders every segment by voxel and keeps the segments apart; everything else is the single-cloud
// algorithm above, per segment: same bounding-box rule, same float centroid accumulation in sorted order.
// A segment whose index range overflows int (PCL: "leaf size is too small", input returned unchanged) or idx_bits keeps its
// points as they are (key = position).

__device__ __forceinline__ int seg_count(const SegFilter& f) { return min(max(*f.n_seg - f.seg0, 0), f.seg_cap); }

__global__ void __launch_bounds__(256) k_seg_prep(SegFilter f) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  __shared__ int s_w[8];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int S = seg_count(f);
  const int n = t < S ? max(0, *f.seg[f.seg0 + t].n_in) : 0;
  int incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  int base = 0;
  for (int v = 0; v < w; ++v) base += s_w[v];
  if (t <= ALOAM_MAX_SEGS) f.off[t] = 0;
  __syncthreads();
  if (t < S) f.off[t] = base + incl - n;
  if (t == S - 1 || (S == 0 && t == 0)) {
    const int total = S == 0 ? 0 : base + incl;
    for (int v = max(S, 0); v <= ALOAM_MAX_SEGS; ++v) f.off[v] = total;
    *f.total = total;
  }
  if (t < ALOAM_MAX_SEGS) {
    f.bbox[6 * t + 0] = INT_MAX; f.bbox[6 * t + 1] = INT_MAX; f.bbox[6 * t + 2] = INT_MAX;
    f.bbox[6 * t + 3] = INT_MIN; f.bbox[6 * t + 4] = INT_MIN; f.bbox[6 * t + 5] = INT_MIN;
  }
}

// grid (chunks, segments)
__global__ void __launch_bounds__(256) k_seg_bbox(SegFilter f) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  const int sgm = blockIdx.y;
  if (sgm >= seg_count(f)) return;
  const SegDesc d = f.seg[f.seg0 + sgm];
  const int n = *d.n_in;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  bool any = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Pt4 p = d.src[i];
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    any = true;
  }
  if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], dd));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], dd));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&f.bbox[6 * sgm + a], f2ord(lo[a])); atomicMax(&f.bbox[6 * sgm + 3 + a], f2ord(hi[a])); }
  }
}

// grid (chunks, segments): keys / vals in the compact order (segment after segment)
__global__ void __launch_bounds__(256) k_seg_keys(SegFilter f, unsigned* __restrict__ keys, int* __restrict__ vals) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  const int sgm = blockIdx.y;
  if (sgm >= seg_count(f)) return;
  const SegDesc d = f.seg[f.seg0 + sgm];
  const int n = *d.n_in;
  if (n <= 0) return;
  const int off = f.off[sgm];
  float mn[3], mx[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { mn[a] = ord2f(f.bbox[6 * sgm + a]); mx[a] = ord2f(f.bbox[6 * sgm + 3 + a]); }
  const float inv = 1.0f / d.leaf;
  long long dxyz[3]; int min_b[3], div_b[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    dxyz[a] = (long long)((mx[a] - mn[a]) * inv) + 1;
    min_b[a] = (int)floorf(mn[a] * inv);
    div_b[a] = (int)floorf(mx[a] * inv) - min_b[a] + 1;
  }
  const unsigned long long cells = (unsigned long long)div_b[0] * (unsigned long long)div_b[1] * (unsigned long long)div_b[2];
  const bool pcl_overflow = dxyz[0] * dxyz[1] * dxyz[2] > (long long)INT_MAX;   // PCL: "leaf size is too small", input returned unchanged
  const bool key_overflow = !pcl_overflow && cells > (1ull << f.idx_bits);       // cannot be encoded: flagged, points kept as they are
  const bool pass = pcl_overflow || key_overflow;
  if (key_overflow && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(f.err, 1);
  const unsigned hi_bits = (unsigned)sgm << f.idx_bits;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned idx;
    if (pass) idx = (unsigned)i;
    else {
      const Pt4 p = d.src[i];
      const int i0 = (int)(floorf(p.x * inv) - (float)min_b[0]);
      const int i1 = (int)(floorf(p.y * inv) - (float)min_b[1]);
      const int i2 = (int)(floorf(p.z * inv) - (float)min_b[2]);
      idx = (unsigned)(i0 + i1 * div_b[0] + i2 * div_b[0] * div_b[1]);
    }
    keys[off + i] = hi_bits | idx;
    vals[off + i] = off + i;
  }
}

// heads before the first sorted slot of every segment (grid: segments + 1 CTAs of one warp; CTA s ranks slot off[s])
__device__ __forceinline__ int seg_head_rank(const unsigned* __restrict__ keys, const int* __restrict__ block_offsets, int pos, int total) {
  const int lane = threadIdx.x & 31;
  if (total <= 0) return 0;
  const int p = min(pos, total);
  // heads in [block start, p): for p == total the block is the last one (all of its heads count)
  const int b = p >= total ? (total - 1) / VCH : p / VCH;
  int cnt = 0;
  for (int i = b * VCH + lane; i < p; i += 32) cnt += (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
#pragma unroll
  for (int dd = 16; dd > 0; dd >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, dd);
  return block_offsets[b] + cnt;
}
__global__ void __launch_bounds__(32) k_seg_rank0(SegFilter f, const unsigned* __restrict__ keys, const int* __restrict__ block_offsets) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  const int S = seg_count(f), total = *f.total;
  const int sgm = blockIdx.x;
  if (sgm > S) return;
  const int r0 = seg_head_rank(keys, block_offsets, sgm == S ? total : f.off[sgm], total);
  if (threadIdx.x == 0) f.rank0[sgm] = r0;
  if (sgm < S) {   // the filtered size of this segment needs the rank of the next boundary too
    const int r1 = seg_head_rank(keys, block_offsets, sgm + 1 == S ? total : f.off[sgm + 1], total);
    if (threadIdx.x == 0) *f.seg[f.seg0 + sgm].n_out = r1 - r0;
  }
}

// every head sums its run in sorted order (float, like pcl::CentroidPoint); output slot = off[segment] + rank inside the segment
__global__ void __launch_bounds__(VT) k_seg_emit(SegFilter f, const unsigned* __restrict__ keys, const int* __restrict__ vals,
                                                 const int* __restrict__ block_offsets, Pt4* __restrict__ tmp) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  const int n = *f.total;
  if (blockIdx.x * VCH >= n) return;
  __shared__ int s_w[VT / 32];
  __shared__ Pt4 s_pts[VCH];        // the block's points in sorted order: the runs are summed from shared memory, the gather
  __shared__ unsigned s_keys[VCH];  // through vals[] is issued once, all loads in flight
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b0 = blockIdx.x * VCH;
#pragma unroll
  for (int k = 0; k < VIT; ++k) {
    const int i = b0 + k * VT + tid;
    if (i < n) {
      const unsigned key = keys[i];
      const int sgm = (int)(key >> f.idx_bits);
      s_keys[k * VT + tid] = key;
      s_pts[k * VT + tid] = f.seg[f.seg0 + sgm].src[vals[i] - f.off[sgm]];
    }
  }
  __syncthreads();
  const int base = b0 + tid * VIT;
  int heads = 0;
  bool is_head[VIT];
#pragma unroll
  for (int k = 0; k < VIT; ++k) {
    const int i = base + k;
    is_head[k] = i < n && (i == 0 || keys[i] != keys[i - 1]);
    heads += is_head[k] ? 1 : 0;
  }
  int incl = heads;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  int wbase = 0;
  for (int v = 0; v < w; ++v) wbase += s_w[v];
  int rank = block_offsets[blockIdx.x] + wbase + incl - heads;
  const int bend = min(n, b0 + VCH);
#pragma unroll
  for (int k = 0; k < VIT; ++k) {
    if (!is_head[k]) continue;
    const int i = base + k;
    const unsigned key = s_keys[i - b0];
    const int sgm = (int)(key >> f.idx_bits);
    const int off = f.off[sgm];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int cnt = 0;
    int j = i;
    for (; j < bend && s_keys[j - b0] == key; ++j) {   // float accumulation in sorted order (pcl::CentroidPoint)
      const Pt4 p = s_pts[j - b0];
      sx += p.x; sy += p.y; sz += p.z; si += p.i;
      ++cnt;
    }
    if (j == bend) {   // the run continues in the next block(s): finish it from global memory
      const Pt4* __restrict__ src = f.seg[f.seg0 + sgm].src;
      for (; j < n && keys[j] == key; ++j) {
        const Pt4 p = src[vals[j] - off];
        sx += p.x; sy += p.y; sz += p.z; si += p.i;
        ++cnt;
      }
    }
    const float nf = (float)cnt;
    Pt4 o; o.x = sx / nf; o.y = sy / nf; o.z = sz / nf; o.i = si / nf;
    tmp[off + (rank - f.rank0[sgm])] = o;
    ++rank;
  }
}

// grid (chunks, segments): filtered points back to their destination (which may be the source slab itself)
__global__ void __launch_bounds__(256) k_seg_writeback(SegFilter f, const Pt4* __restrict__ tmp) {
  pdl_launch_dependents();
  pdl_wait();   // may have been launched with a programmatic dependency on the previous kernel of the stream
  const int sgm = blockIdx.y;
  if (sgm >= seg_count(f)) return;
  const SegDesc d = f.seg[f.seg0 + sgm];
  const int m = f.rank0[sgm + 1] - f.rank0[sgm];
  const int off = f.off[sgm];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) d.dst[i] = tmp[off + i];
}

}  // namespace aloam

using namespace aloam;

// segmented filter over up to S_upper segments; n_upper / per_seg_upper are host bounds for grid sizing (total points, points of
// one segment).  A 32-bit key holds f.idx_bits index bits and the segment number: when S_upper segments do not fit beside the
// index, the segments are processed in windows of 2^(32 - idx_bits).
void vox_seg_filter(aloam_ctx* c, const SegFilter& f_in, SegBuffers& b, int S_upper, int n_upper, int per_seg_upper) {
  const int nblk = std::max(1, (n_upper + VCH - 1) / VCH);
  const int chunks = std::max(1, std::min((per_seg_upper + 255) / 256, S_upper > 8 ? 8 : 64));   // grid-stride inside a segment
  const int win = (int)std::min<long long>(1ll << std::max(0, 32 - f_in.idx_bits), ALOAM_MAX_SEGS);
  for (int s0 = 0; s0 < S_upper; s0 += win) {
    SegFilter f = f_in;
    f.seg0 = s0; f.seg_cap = std::min(win, S_upper - s0);
    int seg_bits = 0; while ((1 << seg_bits) < f.seg_cap) ++seg_bits;
    const int bits = f.idx_bits + seg_bits;
    launch_ex(c, KID_VOXEL, k_seg_prep, dim3(1), dim3(256), 0, 1, true, f);
    launch_ex(c, KID_VOXEL, k_seg_bbox, dim3(dim3(chunks, f.seg_cap)), dim3(256), 0, 1, true, f);
    launch_ex(c, KID_VOXEL, k_seg_keys, dim3(dim3(chunks, f.seg_cap)), dim3(256), 0, 1, true, f, b.keys[0], b.vals[0]);
    // the radix passes: one cooperative launch (all CTAs co-resident, at most one per SM)
    static int coop_cap = 0;
    if (!coop_cap) {
      int per_sm = 0, sms = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_radix_sort_all, VT, 0);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->cfg.device);
      coop_cap = std::max(1, std::min(sms, per_sm * sms));
    }
    {
      prof_begin(c, KID_VOXEL);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(std::min(nblk, coop_cap)); cfg.blockDim = dim3(VT); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k_radix_sort_all, b.keys[0], b.vals[0], b.keys[1], b.vals[1], (const int*)f.total, bits, b.hist);
      prof_end(c);
    }
    const int cur = ((bits + 7) / 8) & 1;
    launch_ex(c, KID_VOXEL, k_vox_heads, dim3(nblk), dim3(VT), 0, 1, true, b.keys[cur], f.total, b.block_heads);
    launch_ex(c, KID_VOXEL, k_vox_blockscan, dim3(1), dim3(1024), 0, 1, true, b.block_heads, f.total, b.heads_total);
    launch_ex(c, KID_VOXEL, k_seg_rank0, dim3(f.seg_cap + 1), dim3(32), 0, 1, true, f, b.keys[cur], b.block_heads);
    launch_ex(c, KID_VOXEL, k_seg_emit, dim3(nblk), dim3(VT), 0, 1, true, f, b.keys[cur], b.vals[cur], b.block_heads, b.tmp);
    launch_ex(c, KID_VOXEL, k_seg_writeback, dim3(dim3(chunks, f.seg_cap)), dim3(256), 0, 1, true, f, b.tmp);
  }
}

int vox_seg_alloc(SegBuffers& b, size_t cap) {
  const size_t nb = (cap + VCH - 1) / VCH + 1;
  b.cap = cap;
  if (cudaMalloc((void**)&b.keys[0], cap * 4) != cudaSuccess || cudaMalloc((void**)&b.keys[1], cap * 4) != cudaSuccess ||
      cudaMalloc((void**)&b.vals[0], cap * 4) != cudaSuccess || cudaMalloc((void**)&b.vals[1], cap * 4) != cudaSuccess ||
      cudaMalloc((void**)&b.hist, nb * 256 * 4) != cudaSuccess || cudaMalloc((void**)&b.offs, nb * 256 * 4) != cudaSuccess ||
      cudaMalloc((void**)&b.block_heads, nb * 4) != cudaSuccess || cudaMalloc((void**)&b.heads_total, 16) != cudaSuccess ||
      cudaMalloc((void**)&b.tmp, cap * sizeof(Pt4)) != cudaSuccess)
    return ALOAM_ERR_CUDA;
  return ALOAM_OK;
}
void vox_seg_free(SegBuffers& b) {
  void* ps[] = {b.keys[0], b.keys[1], b.vals[0], b.vals[1], b.hist, b.offs, b.block_heads, b.heads_total, b.tmp};
  for (void* p : ps) if (p) cudaFree(p);
  b = SegBuffers();
}


extern "C" int aloam_voxel_filter_impl(aloam_ctx* c, aloam_cloud_view in, float leaf, aloam_cloud_view* out) {
  if (!c || !out || !(leaf > 0.f)) return ALOAM_ERR_INVALID_ARG;
  int rc = check_view(in); if (rc) return rc;
  if (in.n > c->max_points) return ALOAM_ERR_CAPACITY;
  CUDA_CHECK_RET(cudaSetDevice(c->cfg.device));
  // the result has its OWN pinned buffer: a caller may pass a view returned by aloam_extract_features (h_out[*]) straight
  // back in, so the output must not alias any other ctx-owned view
  if (!c->h_vox_out) CUDA_CHECK_RET(cudaMallocHost((void**)&c->h_vox_out, (size_t)c->max_points * sizeof(Pt4)));
  out->data = reinterpret_cast<const float*>(c->h_vox_out); out->n = 0; out->stride_floats = 4;
  if (in.n == 0) return ALOAM_OK;
  const int n = in.n;
  if (!c->d_vox_keys[0]) {
    const size_t mp = (size_t)c->max_points;
    const int nb_max = (c->max_points + VCH - 1) / VCH;
    CUDA_CHECK_RET(cudaMalloc((void**)&c->d_vox_keys[0], mp * 4)); CUDA_CHECK_RET(cudaMalloc((void**)&c->d_vox_keys[1], mp * 4));
    CUDA_CHECK_RET(cudaMalloc((void**)&c->d_vox_vals[0], mp * 4)); CUDA_CHECK_RET(cudaMalloc((void**)&c->d_vox_vals[1], mp * 4));
    CUDA_CHECK_RET(cudaMalloc((void**)&c->d_vox_hist, (size_t)nb_max * 256 * 4)); CUDA_CHECK_RET(cudaMalloc((void**)&c->d_vox_offs, (size_t)nb_max * 256 * 4));
    CUDA_CHECK_RET(cudaMalloc((void**)&c->d_vox_misc, 64 * 4 + (size_t)nb_max * 4));
  }
  Pt4* d_in = c->d_query;        // staging buffers that already exist in the context
  Pt4* d_out = c->lanes[0].d_full[0];
  rc = upload_cloud(c, in, d_in, c->max_points); if (rc) return rc;
  int* mm6 = c->d_vox_misc; int* d_total = c->d_vox_misc + 8; int* d_n = c->d_vox_misc + 9; int* block_heads = c->d_vox_misc + 64;
  const int init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  std::memcpy(c->h_ints + 32, init, sizeof(init));
  c->h_ints[38] = n;
  CUDA_CHECK_RET(cudaMemcpyAsync(mm6, c->h_ints + 32, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK_RET(cudaMemcpyAsync(d_n, c->h_ints + 38, 4, cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, KID_VOXEL, k_vox_bbox, std::min((n + 255) / 256, 592), 256, 0, d_in, n, mm6);
  CUDA_CHECK_RET(cudaMemcpyAsync(c->h_ints + 40, mm6, 24, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
  // host replica of PCL's index-range computation (needs the box; also decides the number of radix passes)
  float mn[3], mx[3];
  for (int a = 0; a < 3; ++a) {
    int lo = c->h_ints[40 + a], hi = c->h_ints[43 + a];
    lo = lo >= 0 ? lo : (lo ^ 0x7fffffff); hi = hi >= 0 ? hi : (hi ^ 0x7fffffff);
    std::memcpy(&mn[a], &lo, 4); std::memcpy(&mx[a], &hi, 4);
  }
  const float inv = 1.0f / leaf;
  long long dxyz[3]; int min_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    dxyz[a] = (long long)((mx[a] - mn[a]) * inv) + 1;
    min_b[a] = (int)std::floor(mn[a] * inv);
    div_b[a] = (int)std::floor(mx[a] * inv) - min_b[a] + 1;
  }
  const int nblk = (n + VCH - 1) / VCH;
  if (dxyz[0] * dxyz[1] * dxyz[2] > (long long)INT_MAX) {   // "leaf size is too small": PCL returns the input unchanged
    CUDA_CHECK_RET(cudaMemcpyAsync(c->h_vox_out, d_in, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
    out->n = n;
    return ALOAM_OK;
  }
  VoxParams vp; vp.inv = inv; vp.min_b[0] = min_b[0]; vp.min_b[1] = min_b[1]; vp.min_b[2] = min_b[2];
  vp.div0 = div_b[0]; vp.div01 = div_b[0] * div_b[1];
  LAUNCH(c, KID_VOXEL, k_vox_keys, (n + 255) / 256, 256, 0, d_in, n, vp, c->d_vox_keys[0], c->d_vox_vals[0]);
  const unsigned long long cells = (unsigned long long)div_b[0] * (unsigned long long)div_b[1] * (unsigned long long)div_b[2];
  int bits = 1; while (bits < 32 && (1ull << bits) < cells) ++bits;
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    LAUNCH(c, KID_VOXEL, k_radix_hist, nblk, VT, 0, c->d_vox_keys[cur], d_n, shift, c->d_vox_hist);
    LAUNCH(c, KID_VOXEL, k_radix_scan, 1, 1024, 0, c->d_vox_hist, d_n, c->d_vox_offs);
    LAUNCH(c, KID_VOXEL, k_radix_scatter, nblk, VT, 0, c->d_vox_keys[cur], c->d_vox_vals[cur], d_n, shift, c->d_vox_offs, c->d_vox_keys[cur ^ 1],
           c->d_vox_vals[cur ^ 1]);
    cur ^= 1;
  }
  LAUNCH(c, KID_VOXEL, k_vox_heads, nblk, VT, 0, c->d_vox_keys[cur], d_n, block_heads);
  LAUNCH(c, KID_VOXEL, k_vox_blockscan, 1, 1024, 0, block_heads, d_n, d_total);
  LAUNCH(c, KID_VOXEL, k_vox_emit, nblk, VT, 0, d_in, c->d_vox_keys[cur], c->d_vox_vals[cur], d_n, block_heads, d_out);
  CUDA_CHECK_RET(cudaMemcpyAsync(c->h_ints + 48, d_total, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
  CUDA_CHECK_RET(cudaGetLastError());
  const int m = c->h_ints[48];
  if (m > 0) {
    CUDA_CHECK_RET(cudaMemcpyAsync(c->h_vox_out, d_out, (size_t)m * 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
  }
  prof_collect(c);
  out->n = m;
  return ALOAM_OK;
}
