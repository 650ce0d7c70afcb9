// Map cube store of alaserMapping and the per-frame loop around it -- replaces laserMapping.cpp:74-108 (state),
// :142-163 (pose hand-off, pointAssociateToMap), :309-550 (centre cube, ring-buffer shift, 75-cube gather, stack
// filters) and :736-801 (insertion, per-cube VoxelGrid).  SURVEY.md section 8 f-1.
//
// B200 shape: the 21 x 21 x 11 cubes are FIXED-CAPACITY SLABS of one pooled device array per cloud type (sized for
// 180 GB of HBM: 4851 x (16 k + 64 k) points x 16 B = 6.4 GB); the ring-buffer shift of :327-509, which rotates
// 4851 smart pointers on the CPU, becomes a permutation of a 4851-entry slab table on the host -- no point ever moves
// -- and the cubes that scroll out are emptied by zeroing their counts.  The submap never visits the host: the valid
// cubes are gathered device-to-device and handed to the same index build + registration as aloam_map_upload /
// aloam_mapping_register (which accept device views).  The host keeps the 2 x 4851 counts (they change only in two
// places: insertion, read back once per frame, and the re-filter, whose output sizes it sees).
//
// First version of this row: the per-cube re-filter calls the single-cloud voxel filter once per non-empty valid
// cube (a segmented filter over all valid cubes at once is the next step).
#include "common.cuh"
#include "ctx.h"

extern "C" {
int aloam_map_upload_impl(aloam_ctx* c, aloam_cloud_view corner_map, aloam_cloud_view surf_map);
int aloam_mapping_register_impl(aloam_ctx* c, aloam_cloud_view corner_stack, aloam_cloud_view surf_stack, double x[7], aloam_stats* stats);
int aloam_voxel_filter_impl(aloam_ctx* c, aloam_cloud_view in, float leaf, aloam_cloud_view* out);
}

namespace {

constexpr int CW = 21, CH = 21, CD = 11, NCUBE = CW * CH * CD;   // laserCloudWidth / Height / Depth (:77-82)
inline int cube_index(int i, int j, int k) { return i + CW * j + CW * CH * k; }

struct Mapper {
  int slab_of[NCUBE];                       // cube index -> physical slab (the reference permutes pointers instead)
  int cen[3] = {10, 10, 5};                 // laserCloudCenWidth / Height / Depth (:74-76)
  double q_wmap_wodom[4] = {0, 0, 0, 1}, t_wmap_wodom[3] = {0, 0, 0};   // :116-117
  int cap[2] = {0, 0};
  Pt4* d_pts[2] = {nullptr, nullptr};       // [NCUBE * cap] slabs
  int* d_cnt[2] = {nullptr, nullptr};       // per-frame device copy of the counts, by slab
  std::vector<int> h_cnt[2];                // authoritative counts, by slab
  Pt4* d_sub[2] = {nullptr, nullptr};       // gathered submap
  Pt4* d_world = nullptr;                   // insertion scratch: transformed stack points
  int* d_slab = nullptr;                    //                    their slabs (-1 = outside the block)
  int* d_slab_of = nullptr;
  int* d_err = nullptr;
  std::vector<int> valid;                   // laserCloudValidInd of the last step
  std::vector<float> stack[2];              // voxel-filtered current clouds (host copies)
  int frames = 0;
};

// ---- Eigen-order quaternion helpers (operation order of Eigen::Quaternion: the pose hand-off is compared bit for bit)
struct Qd { double x, y, z, w; };
inline Qd qmul(const Qd& a, const Qd& b) {
  Qd r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
inline Qd qinv(const Qd& a) {
  const double n2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  return Qd{-a.x / n2, -a.y / n2, -a.z / n2, a.w / n2};
}
inline void qrot(const Qd& q, const double v[3], double o[3]) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] = uv[0] + uv[0]; uv[1] = uv[1] + uv[1]; uv[2] = uv[2] + uv[2];
  const double c2[3] = {q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0]};
  o[0] = v[0] + q.w * uv[0] + c2[0]; o[1] = v[1] + q.w * uv[1] + c2[1]; o[2] = v[2] + q.w * uv[2] + c2[2];
}

// int((v + 25.0) / 50.0) + centre, minus one when v + 25.0 < 0   (:314-325, :741-750)
__host__ __device__ inline int cube_coord(double v, int centre) {
  int c = int((v + 25.0) / 50.0) + centre;
  if (v + 25.0 < 0) c--;
  return c;
}

struct Pose7 { double v[7]; };
struct Cen3 { int v[3]; };

// pointAssociateToMap (:154-163) in double, stored as float, then the cube of the stored point (:741-758)
__global__ void k_cube_ids(const Pt4* __restrict__ stack, int n, Pose7 x, Cen3 cen, const int* __restrict__ slab_of,
                           Pt4* __restrict__ world, int* __restrict__ slab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Pt4 p = stack[i];
  const double ux = x.v[0], uy = x.v[1], uz = x.v[2], w = x.v[3];
  const double vx = (double)p.x, vy = (double)p.y, vz = (double)p.z;
  double uvx = uy * vz - uz * vy, uvy = uz * vx - ux * vz, uvz = ux * vy - uy * vx;
  uvx = uvx + uvx; uvy = uvy + uvy; uvz = uvz + uvz;
  const double cx = uy * uvz - uz * uvy, cy = uz * uvx - ux * uvz, cz = ux * uvy - uy * uvx;
  Pt4 s;
  s.x = (float)(((vx + w * uvx) + cx) + x.v[4]);
  s.y = (float)(((vy + w * uvy) + cy) + x.v[5]);
  s.z = (float)(((vz + w * uvz) + cz) + x.v[6]);
  s.i = p.i;
  const int ci = cube_coord((double)s.x, cen.v[0]), cj = cube_coord((double)s.y, cen.v[1]), ck = cube_coord((double)s.z, cen.v[2]);
  world[i] = s;
  slab[i] = (ci >= 0 && ci < CW && cj >= 0 && cj < CH && ck >= 0 && ck < CD) ? slab_of[ci + CW * cj + CW * CH * ck] : -1;
}

// push_back order = stack order: the rank of a point inside its cube is the number of EARLIER stack points of the same
// cube (brute force over <= a few thousand points; the slab ids of a chunk are staged in shared memory)
__global__ void __launch_bounds__(256) k_cube_place(const Pt4* __restrict__ world, const int* __restrict__ slab, int n,
                                                    const int* __restrict__ cnt, int cap, Pt4* __restrict__ pts, int* __restrict__ err) {
  __shared__ int s_slab[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int mine = i < n ? slab[i] : -1;
  int rank = 0;
  const int last_chunk = blockIdx.x;   // only points with a smaller index matter
  for (int ch = 0; ch <= last_chunk; ++ch) {
    const int j = ch * 256 + threadIdx.x;
    __syncthreads();
    s_slab[threadIdx.x] = j < n ? slab[j] : -2;
    __syncthreads();
    const int lim = ch == last_chunk ? (int)threadIdx.x : 256;
    if (mine >= 0)
      for (int t = 0; t < lim; ++t) rank += (s_slab[t] == mine) ? 1 : 0;
  }
  if (mine < 0) return;
  const int pos = cnt[mine] + rank;
  if (pos >= cap) { atomicExch(err, 1); return; }
  pts[(size_t)mine * cap + pos] = world[i];
}

__global__ void k_cube_bump(const int* __restrict__ slab, int n, int cap, int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = slab[i];
  if (s >= 0) { const int old = atomicAdd(&cnt[s], 1); if (old >= cap) atomicSub(&cnt[s], 1); }
}

int ensure_mapper(aloam_ctx* c) {
  if (c->mapper) return ALOAM_OK;
  Mapper* m = new (std::nothrow) Mapper();
  if (!m) return ALOAM_ERR_CUDA;
  for (int i = 0; i < NCUBE; ++i) m->slab_of[i] = i;
  m->cap[0] = std::min(16384, c->max_points);
  m->cap[1] = std::min(65536, c->max_points);
  c->mapper = m;
  for (int t = 0; t < 2; ++t) {
    m->h_cnt[t].assign(NCUBE, 0);
    CUDA_CHECK_RET(cudaMalloc((void**)&m->d_pts[t], (size_t)NCUBE * m->cap[t] * sizeof(Pt4)));
    CUDA_CHECK_RET(cudaMalloc((void**)&m->d_cnt[t], NCUBE * sizeof(int)));
    CUDA_CHECK_RET(cudaMalloc((void**)&m->d_sub[t], (size_t)std::max(c->cfg.max_map_points, 1) * sizeof(Pt4)));
  }
  CUDA_CHECK_RET(cudaMalloc((void**)&m->d_world, (size_t)c->max_points * sizeof(Pt4)));
  CUDA_CHECK_RET(cudaMalloc((void**)&m->d_slab, (size_t)c->max_points * sizeof(int)));
  CUDA_CHECK_RET(cudaMalloc((void**)&m->d_slab_of, NCUBE * sizeof(int)));
  CUDA_CHECK_RET(cudaMalloc((void**)&m->d_err, sizeof(int)));
  return ALOAM_OK;
}

// :327-509 -- one step of the ring buffer along `axis`; towards_high: every cube moves one index up, the top cube wraps
// to index 0 and is emptied
void rotate_axis(Mapper* m, int axis, bool towards_high) {
  const int n[3] = {CW, CH, CD};
  const int a = axis, b = (axis + 1) % 3, cc = (axis + 2) % 3;
  for (int u = 0; u < n[b]; ++u) {
    for (int v = 0; v < n[cc]; ++v) {
      auto at = [&](int t) { int ijk[3]; ijk[a] = t; ijk[b] = u; ijk[cc] = v; return cube_index(ijk[0], ijk[1], ijk[2]); };
      if (towards_high) {
        const int wrapped = m->slab_of[at(n[a] - 1)];
        for (int t = n[a] - 1; t >= 1; --t) m->slab_of[at(t)] = m->slab_of[at(t - 1)];
        m->slab_of[at(0)] = wrapped;
        m->h_cnt[0][wrapped] = 0; m->h_cnt[1][wrapped] = 0;
      } else {
        const int wrapped = m->slab_of[at(0)];
        for (int t = 0; t < n[a] - 1; ++t) m->slab_of[at(t)] = m->slab_of[at(t + 1)];
        m->slab_of[at(n[a] - 1)] = wrapped;
        m->h_cnt[0][wrapped] = 0; m->h_cnt[1][wrapped] = 0;
      }
    }
  }
}

}  // namespace

extern "C" void aloam_mapper_free_impl(aloam_ctx* c) {
  Mapper* m = static_cast<Mapper*>(c->mapper);
  if (!m) return;
  for (int t = 0; t < 2; ++t) { if (m->d_pts[t]) cudaFree(m->d_pts[t]); if (m->d_cnt[t]) cudaFree(m->d_cnt[t]); if (m->d_sub[t]) cudaFree(m->d_sub[t]); }
  if (m->d_world) cudaFree(m->d_world);
  if (m->d_slab) cudaFree(m->d_slab);
  if (m->d_slab_of) cudaFree(m->d_slab_of);
  if (m->d_err) cudaFree(m->d_err);
  delete m;
  c->mapper = nullptr;
}

extern "C" {

int aloam_mapper_reset(aloam_ctx* c) {
  if (!c) return ALOAM_ERR_INVALID_ARG;
  CUDA_CHECK_RET(cudaSetDevice(c->cfg.device));
  int rc = ensure_mapper(c); if (rc) return rc;
  Mapper* m = static_cast<Mapper*>(c->mapper);
  for (int i = 0; i < NCUBE; ++i) m->slab_of[i] = i;
  for (int t = 0; t < 2; ++t) std::fill(m->h_cnt[t].begin(), m->h_cnt[t].end(), 0);
  m->cen[0] = 10; m->cen[1] = 10; m->cen[2] = 5;
  m->q_wmap_wodom[0] = m->q_wmap_wodom[1] = m->q_wmap_wodom[2] = 0; m->q_wmap_wodom[3] = 1;
  m->t_wmap_wodom[0] = m->t_wmap_wodom[1] = m->t_wmap_wodom[2] = 0;
  m->valid.clear(); m->frames = 0;
  return ALOAM_OK;
}

int aloam_mapper_step(aloam_ctx* c, aloam_cloud_view corner_last, aloam_cloud_view surf_last, const double q_wodom_curr[4],
                      const double t_wodom_curr[3], double q_w_curr[4], double t_w_curr[3], aloam_stats* stats) {
  if (!c || !q_wodom_curr || !t_wodom_curr || !q_w_curr || !t_w_curr) return ALOAM_ERR_INVALID_ARG;
  CUDA_CHECK_RET(cudaSetDevice(c->cfg.device));
  int rc = ensure_mapper(c); if (rc) return rc;
  Mapper* m = static_cast<Mapper*>(c->mapper);
  // ---- transformAssociateToMap (:142-146)
  const Qd qm{m->q_wmap_wodom[0], m->q_wmap_wodom[1], m->q_wmap_wodom[2], m->q_wmap_wodom[3]};
  const Qd qo{q_wodom_curr[0], q_wodom_curr[1], q_wodom_curr[2], q_wodom_curr[3]};
  const Qd q0 = qmul(qm, qo);
  double r[3]; qrot(qm, t_wodom_curr, r);
  double x[7] = {q0.x, q0.y, q0.z, q0.w, r[0] + m->t_wmap_wodom[0], r[1] + m->t_wmap_wodom[1], r[2] + m->t_wmap_wodom[2]};
  // ---- centre cube and ring-buffer shift (:314-509)
  int ctr[3] = {cube_coord(x[4], m->cen[0]), cube_coord(x[5], m->cen[1]), cube_coord(x[6], m->cen[2])};
  const int dims[3] = {CW, CH, CD};
  for (int a = 0; a < 3; ++a) {
    while (ctr[a] < 3) { rotate_axis(m, a, true); ctr[a]++; m->cen[a]++; }
    while (ctr[a] >= dims[a] - 3) { rotate_axis(m, a, false); ctr[a]--; m->cen[a]--; }
  }
  // ---- valid cubes (:511-529) and device-to-device gather (:531-539)
  m->valid.clear();
  for (int i = ctr[0] - 2; i <= ctr[0] + 2; ++i)
    for (int j = ctr[1] - 2; j <= ctr[1] + 2; ++j)
      for (int k = ctr[2] - 1; k <= ctr[2] + 1; ++k)
        if (i >= 0 && i < CW && j >= 0 && j < CH && k >= 0 && k < CD) m->valid.push_back(cube_index(i, j, k));
  int n_sub[2] = {0, 0};
  for (int t = 0; t < 2; ++t) {
    long long tot = 0;
    for (int ind : m->valid) tot += m->h_cnt[t][m->slab_of[ind]];
    if (tot > c->cfg.max_map_points) return ALOAM_ERR_CAPACITY;
    for (int ind : m->valid) {
      const int s = m->slab_of[ind], n = m->h_cnt[t][s];
      if (n > 0) CUDA_CHECK_RET(cudaMemcpyAsync(m->d_sub[t] + n_sub[t], m->d_pts[t] + (size_t)s * m->cap[t], (size_t)n * sizeof(Pt4), cudaMemcpyDeviceToDevice, c->stream));
      n_sub[t] += n;
    }
  }
  rc = aloam_map_upload_impl(c, aloam_cloud_view{reinterpret_cast<const float*>(m->d_sub[0]), n_sub[0], 4},
                             aloam_cloud_view{reinterpret_cast<const float*>(m->d_sub[1]), n_sub[1], 4});
  if (rc) return rc;
  // ---- stack filters (:541-550)
  const aloam_cloud_view last[2] = {corner_last, surf_last};
  const float leaf[2] = {c->cfg.line_res, c->cfg.plane_res};
  for (int t = 0; t < 2; ++t) {
    aloam_cloud_view out;
    rc = aloam_voxel_filter_impl(c, last[t], leaf[t], &out); if (rc) return rc;
    m->stack[t].assign(out.data, out.data + (size_t)out.n * 4);
  }
  const aloam_cloud_view st[2] = {aloam_cloud_view{m->stack[0].data(), (int)(m->stack[0].size() / 4), 4},
                                  aloam_cloud_view{m->stack[1].data(), (int)(m->stack[1].size() / 4), 4}};
  // ---- optimisation (:554-733)
  rc = aloam_mapping_register_impl(c, st[0], st[1], x, stats); if (rc) return rc;
  // ---- transformUpdate (:148-152)
  const Qd qw{x[0], x[1], x[2], x[3]};
  const Qd qn = qmul(qw, qinv(qo));
  qrot(qn, t_wodom_curr, r);
  m->q_wmap_wodom[0] = qn.x; m->q_wmap_wodom[1] = qn.y; m->q_wmap_wodom[2] = qn.z; m->q_wmap_wodom[3] = qn.w;
  m->t_wmap_wodom[0] = x[4] - r[0]; m->t_wmap_wodom[1] = x[5] - r[1]; m->t_wmap_wodom[2] = x[6] - r[2];
  for (int k = 0; k < 4; ++k) q_w_curr[k] = x[k];
  for (int k = 0; k < 3; ++k) t_w_curr[k] = x[4 + k];
  // ---- insertion (:736-767)
  Pose7 px; for (int k = 0; k < 7; ++k) px.v[k] = x[k];
  Cen3 cen; for (int a = 0; a < 3; ++a) cen.v[a] = m->cen[a];
  CUDA_CHECK_RET(cudaMemcpyAsync(m->d_slab_of, m->slab_of, sizeof(m->slab_of), cudaMemcpyHostToDevice, c->stream));
  CUDA_CHECK_RET(cudaMemsetAsync(m->d_err, 0, sizeof(int), c->stream));
  Pt4* d_stack = c->d_stack_corner;
  for (int t = 0; t < 2; ++t) {
    const int n = st[t].n;
    CUDA_CHECK_RET(cudaMemcpyAsync(m->d_cnt[t], m->h_cnt[t].data(), NCUBE * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (n > 0) {
      rc = upload_cloud(c, st[t], d_stack, c->max_points); if (rc) return rc;
      const int nb = (n + 255) / 256;
      k_cube_ids<<<nb, 256, 0, c->stream>>>(d_stack, n, px, cen, m->d_slab_of, m->d_world, m->d_slab);
      k_cube_place<<<nb, 256, 0, c->stream>>>(m->d_world, m->d_slab, n, m->d_cnt[t], m->cap[t], m->d_pts[t], m->d_err);
      k_cube_bump<<<nb, 256, 0, c->stream>>>(m->d_slab, n, m->cap[t], m->d_cnt[t]);
      c->launches += 3;
    }
    CUDA_CHECK_RET(cudaMemcpyAsync(m->h_cnt[t].data(), m->d_cnt[t], NCUBE * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));   // the host stack buffer is reused, the counts are needed below
  }
  int err = 0;
  CUDA_CHECK_RET(cudaMemcpy(&err, m->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) return ALOAM_ERR_CAPACITY;
  // ---- per-cube re-filter of the valid cubes (:770-788)
  for (int ind : m->valid) {
    const int s = m->slab_of[ind];
    for (int t = 0; t < 2; ++t) {
      const int n = m->h_cnt[t][s];
      if (n == 0) continue;
      Pt4* slab = m->d_pts[t] + (size_t)s * m->cap[t];
      aloam_cloud_view out;
      rc = aloam_voxel_filter_impl(c, aloam_cloud_view{reinterpret_cast<const float*>(slab), n, 4}, leaf[t], &out); if (rc) return rc;
      CUDA_CHECK_RET(cudaMemcpyAsync(slab, out.data, (size_t)out.n * sizeof(Pt4), cudaMemcpyHostToDevice, c->stream));
      CUDA_CHECK_RET(cudaStreamSynchronize(c->stream));
      m->h_cnt[t][s] = out.n;
    }
  }
  ++m->frames;
  return ALOAM_OK;
}

int aloam_mapper_debug_state(aloam_ctx* c, int centre[3], int* n_valid, int valid[125], double q_wmap_wodom[4], double t_wmap_wodom[3],
                             long long totals[2]) {
  if (!c || !c->mapper) return ALOAM_ERR_STATE;
  Mapper* m = static_cast<Mapper*>(c->mapper);
  if (centre) for (int a = 0; a < 3; ++a) centre[a] = m->cen[a];
  if (n_valid) *n_valid = (int)m->valid.size();
  if (valid) for (size_t i = 0; i < m->valid.size() && i < 125; ++i) valid[i] = m->valid[i];
  if (q_wmap_wodom) for (int k = 0; k < 4; ++k) q_wmap_wodom[k] = m->q_wmap_wodom[k];
  if (t_wmap_wodom) for (int k = 0; k < 3; ++k) t_wmap_wodom[k] = m->t_wmap_wodom[k];
  if (totals) for (int t = 0; t < 2; ++t) { totals[t] = 0; for (int v : m->h_cnt[t]) totals[t] += v; }
  return ALOAM_OK;
}

int aloam_mapper_debug_cube(aloam_ctx* c, int which, int cube, aloam_cloud_view* out) {
  if (!c || !c->mapper || !out || which < 0 || which > 1 || cube < 0 || cube >= NCUBE) return ALOAM_ERR_INVALID_ARG;
  Mapper* m = static_cast<Mapper*>(c->mapper);
  CUDA_CHECK_RET(cudaSetDevice(c->cfg.device));
  const int s = m->slab_of[cube], n = m->h_cnt[which][s];
  out->data = reinterpret_cast<const float*>(c->h_out[4]); out->n = n; out->stride_floats = 4;
  if (n > 0) CUDA_CHECK_RET(cudaMemcpy(c->h_out[4], m->d_pts[which] + (size_t)s * m->cap[which], (size_t)n * sizeof(Pt4), cudaMemcpyDeviceToHost));
  return ALOAM_OK;
}

}  // extern "C"
