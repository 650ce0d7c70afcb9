// Device-resident Levenberg-Marquardt -- replaces the ceres::Solve calls at laserOdometry.cpp:494-499 and
// laserMapping.cpp:712-720 together with the residual functors of lidarFactor.hpp (:12-55 edge, :57-104 plane,
// :106-138 plane-norm) and everything Ceres drives around them (AutoDiff Jacobians, HuberLoss(0.1) + Corrector,
// EigenQuaternionParameterization, trust-region loop, DENSE_QR step).
//
// B200 shape: the whole solve is ONE kernel launch and never returns to the host.  Each pass evaluates every
// residual block with the closed-form tangent Jacobian (SURVEY.md 8a "Residual math"), accumulates the 28
// numbers that the 6-dof problem reduces to -- upper triangle of J^T J (21), J^T r (6), cost (1) -- with a
// fixed-shape warp-shuffle + shared-memory tree (deterministic, no float atomics), and thread 0 takes the
// trust-region decision exactly as Ceres' TrustRegionMinimizer / LevenbergMarquardtStrategy would:
//   Jacobi scaling 1/(1+||J_j||) fixed at iteration 0, D^2 = clamp(diag(Js^T Js), 1e-6, 1e32) (re-used after a
//   rejected step), (Js^T Js + D^2/radius) y = Js^T r  [normal-equation form of Ceres' QR on [Js; sqrt(D^2/radius)]],
//   model_cost_change, Plus(), parameter / function tolerance tests, rho > 1e-3 accept with radius/(max(1/3,1-(2rho-1)^3)),
//   reject with radius/decrease_factor, decrease_factor *= 2.
// Because the candidate evaluation already carries J^T J and J^T r, an accepted step needs no second pass
// (Ceres evaluates cost-only, then re-evaluates with Jacobians): <= 1 + max_iters passes per solve instead of <= 1 + 2*max_iters.
#include <cfloat>
#include "common.cuh"
#include "kernels.h"

namespace aloam {

namespace {

constexpr int NT = ALOAM_LM_THREADS;
constexpr int NW = NT / 32;

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 crossd(const V3& a, const V3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// acc layout: [0..20] upper triangle of J^T J row-major (00,01,..05,11,..,55), [21..26] J^T r, [27] cost
__device__ __forceinline__ void accumulate_row(double* acc, const double j[6], double r) {
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
#pragma unroll
    for (int b = a; b < 6; ++b) acc[k++] += j[a] * j[b];
  }
#pragma unroll
  for (int a = 0; a < 6; ++a) acc[21 + a] += j[a] * r;
}

__device__ __forceinline__ void eval_block(const BlockRec& rb, const double* x, double huber_a, double* acc) {
  // lp = R(q) cp + t   (s == 1 for every block the reference builds; slerp(1, q) == q)
  const V3 u{x[0], x[1], x[2]};
  const double w = x[3];
  const V3 cp{rb.cp[0], rb.cp[1], rb.cp[2]};
  V3 uv = crossd(u, cp);
  uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
  const V3 c2 = crossd(u, uv);
  const V3 Rp{cp.x + w * uv.x + c2.x, cp.y + w * uv.y + c2.y, cp.z + w * uv.z + c2.z};
  const V3 lp{Rp.x + x[4], Rp.y + x[5], Rp.z + x[6]};
  // d lp / d dtheta = -2 [Rp]x  (Ceres Plus is delta_q (x) q with a half-angle delta) ; d lp / d t = I
  // row of J for a residual with gradient n wrt lp:  [ n^T (-2[Rp]x) , n^T ] = [ 2 (Rp x n)^T , n^T ]
  if (rb.type == 0) {
    const V3 a{rb.a[0], rb.a[1], rb.a[2]}, b{rb.b[0], rb.b[1], rb.b[2]};
    const V3 la{lp.x - a.x, lp.y - a.y, lp.z - a.z}, lb{lp.x - b.x, lp.y - b.y, lp.z - b.z};
    const V3 nu = crossd(la, lb);
    const double dn = rb.s;  // |a - b|
    double r[3] = {nu.x / dn, nu.y / dn, nu.z / dn};
    const V3 wv{(b.x - a.x) / dn, (b.y - a.y) / dn, (b.z - a.z) / dn};
    // d r / d lp = [wv]x ; rows: n0 = (0,-wz,wy), n1 = (wz,0,-wx), n2 = (-wy,wx,0)
    const V3 n0{0.0, -wv.z, wv.y}, n1{wv.z, 0.0, -wv.x}, n2{-wv.y, wv.x, 0.0};
    const double sq = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double rho0, rho1;
    const double bb = huber_a * huber_a;
    if (sq > bb) { const double rr = sqrt(sq); rho0 = 2.0 * huber_a * rr - bb; rho1 = fmax(DBL_MIN, huber_a / rr); }
    else { rho0 = sq; rho1 = 1.0; }
    acc[27] += 0.5 * rho0;
    const double sr = sqrt(rho1);
    const V3 ns[3] = {n0, n1, n2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const V3 t = crossd(Rp, ns[k]);  // n^T (-2 [Rp]x) = 2 (Rp x n)^T
      double j[6] = {2.0 * t.x * sr, 2.0 * t.y * sr, 2.0 * t.z * sr, ns[k].x * sr, ns[k].y * sr, ns[k].z * sr};
      accumulate_row(acc, j, r[k] * sr);
    }
  } else {
    V3 n; double r;
    if (rb.type == 1) {
      n = V3{rb.b[0], rb.b[1], rb.b[2]};
      r = (lp.x - rb.a[0]) * n.x + (lp.y - rb.a[1]) * n.y + (lp.z - rb.a[2]) * n.z;
    } else {
      n = V3{rb.a[0], rb.a[1], rb.a[2]};
      r = n.x * lp.x + n.y * lp.y + n.z * lp.z + rb.s;
    }
    const double sq = r * r;
    double rho0, rho1;
    const double bb = huber_a * huber_a;
    if (sq > bb) { const double rr = sqrt(sq); rho0 = 2.0 * huber_a * rr - bb; rho1 = fmax(DBL_MIN, huber_a / rr); }
    else { rho0 = sq; rho1 = 1.0; }
    acc[27] += 0.5 * rho0;
    const double sr = sqrt(rho1);
    const V3 t = crossd(Rp, n);
    double j[6] = {2.0 * t.x * sr, 2.0 * t.y * sr, 2.0 * t.z * sr, n.x * sr, n.y * sr, n.z * sr};
    accumulate_row(acc, j, r * sr);
  }
}

// ceres::EigenQuaternionParameterization::Plus + plain addition on t
__device__ void plus7(const double* x, const double* d, double* o) {
  const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n > 0.0) {
    const double s = sin(n) / n;
    const double ax = s * d[0], ay = s * d[1], az = s * d[2], aw = cos(n);
    const double bx = x[0], by = x[1], bz = x[2], bw = x[3];
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
  } else {
    o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3];
  }
  o[4] = x[4] + d[3]; o[5] = x[5] + d[4]; o[6] = x[6] + d[5];
}

__device__ __forceinline__ int tri(int a, int b) {  // index of (a,b), a<=b, in the packed upper triangle
  return a * 6 - a * (a - 1) / 2 + (b - a);
}

// solve (A) y = b, A symmetric positive definite 6x6 (full storage), Cholesky
__device__ bool chol_solve6(double A[6][6], const double b[6], double y[6]) {
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
    if (!(d > 0.0)) return false;
    d = sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= A[i][k] * A[j][k];
      A[i][j] = s / d;
    }
  }
  double z[6];
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[i][k] * z[k];
    z[i] = s / A[i][i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < 6; ++k) s -= A[k][i] * y[k];
    y[i] = s / A[i][i];
  }
  for (int i = 0; i < 6; ++i)
    if (!isfinite(y[i])) return false;
  return true;
}

}  // namespace

__global__ void k_pack_blocks(const double* __restrict__ packed, int n, BlockRec* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* p = packed + (size_t)i * 11;
  BlockRec r;
  r.type = (int)p[0];
  for (int k = 0; k < 3; ++k) { r.cp[k] = p[1 + k]; r.a[k] = p[4 + k]; r.b[k] = p[7 + k]; }
  r.s = p[10];
  if (r.type == 0) {  // edge: the kernel wants |a-b| ; s of the packed form is the (always 1) interpolation ratio
    const double ex = r.a[0] - r.b[0], ey = r.a[1] - r.b[1], ez = r.a[2] - r.b[2];
    r.s = sqrt(ex * ex + ey * ey + ez * ez);
  }
  r.pad = 0;
  out[i] = r;
}

__global__ void __launch_bounds__(NT, 1) k_lm_solve(const BlockRec* __restrict__ blocks, const int* __restrict__ n_blocks_ptr,
                                                    int n_blocks_host, double* __restrict__ x7, LmParams prm,
                                                    LmSummary* __restrict__ summary, int mode, double* __restrict__ out28,
                                                    double* __restrict__ world7, int integrate) {
  __shared__ double s_part[NW][28];
  __shared__ double s_tot[28];
  __shared__ double s_x[7];
  __shared__ int s_go;
  __shared__ int s_cnt[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = n_blocks_ptr ? *n_blocks_ptr : n_blocks_host;

  if (tid < 7) s_x[tid] = x7[tid];
  if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; }
  __syncthreads();

  // residual-block census (n_edge / n_plane), once
  {
    int ne = 0, np = 0;
    for (int b = tid; b < n; b += NT) { int t = blocks[b].type; ne += (t == 0); np += (t > 0); }
    for (int d = 16; d > 0; d >>= 1) { ne += __shfl_xor_sync(0xffffffffu, ne, d); np += __shfl_xor_sync(0xffffffffu, np, d); }
    if (lane == 0) { atomicAdd(&s_cnt[0], ne); atomicAdd(&s_cnt[1], np); }
  }

  auto evaluate = [&]() {
    double x[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) x[k] = s_x[k];
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
    for (int b = tid; b < n; b += NT) {
      const BlockRec rb = blocks[b];
      if (rb.type >= 0) eval_block(rb, x, prm.huber_a, acc);
    }
#pragma unroll
    for (int k = 0; k < 28; ++k) {
      double v = acc[k];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) s_part[warp][k] = v;
    }
    __syncthreads();
    if (tid < 28) {
      double v = 0.0;
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2) v += s_part[w2][tid];
      s_tot[tid] = v;
    }
    __syncthreads();
  };

  evaluate();

  if (mode == 1) {
    if (tid < 28) out28[tid] = s_tot[tid];
    return;
  }

  // ---------------- thread-0 trust-region state
  double x[7], H[21], g[6], cost = 0, scale[6], diag[6], radius = prm.initial_radius, decrease_factor = 2.0;
  double xc[7], mcc = 0, gmax = 0, x_norm = 0;
  bool reuse_diag = false, last_successful = false;
  int iteration = 0, num_invalid = 0, num_successful = 0, num_evals = 1, termination = 0, trace_rows = 0;
  const int n_res = s_cnt[0] + s_cnt[1];

  auto gradient_max = [&](const double* xx, const double* gg) {
    double ng[6], xp[7];
    for (int k = 0; k < 6; ++k) ng[k] = -gg[k];
    plus7(xx, ng, xp);
    double m = 0;
    for (int k = 0; k < 7; ++k) m = fmax(m, fabs(xp[k] - xx[k]));
    return m;
  };
  auto trace = [&](double c, double cc, double gm, double sn, double rd, double rad, int valid, int succ) {
    if (trace_rows < ALOAM_LM_MAX_TRACE) {
      double* o = summary->trace[trace_rows++];
      o[0] = c; o[1] = cc; o[2] = gm; o[3] = sn; o[4] = rd; o[5] = rad; o[6] = valid; o[7] = succ;
    }
  };
  // produce the next candidate (-> s_x, s_go = 1) or stop (s_go = 0)
  auto next_candidate = [&]() {
    for (;;) {
      if (iteration >= prm.max_iters) { termination = 0; s_go = 0; return; }
      if (last_successful && gmax <= prm.gradient_tolerance) { termination = 1; s_go = 0; return; }
      if (radius < prm.min_radius) { termination = 5; s_go = 0; return; }
      ++iteration;
      last_successful = false;
      if (!reuse_diag)
        for (int j = 0; j < 6; ++j) diag[j] = fmin(fmax(scale[j] * scale[j] * H[tri(j, j)], prm.min_lm_diagonal), prm.max_lm_diagonal);
      double A[6][6], Hs[6][6], b[6], y[6];
      for (int a = 0; a < 6; ++a) {
        for (int c = a; c < 6; ++c) { double v = scale[a] * H[tri(a, c)] * scale[c]; Hs[a][c] = v; Hs[c][a] = v; }
        b[a] = scale[a] * g[a];
      }
      for (int a = 0; a < 6; ++a)
        for (int c = 0; c < 6; ++c) A[a][c] = Hs[a][c] + (a == c ? diag[a] / radius : 0.0);
      bool ok = chol_solve6(A, b, y);
      reuse_diag = true;
      double step[6];
      mcc = 0;
      if (ok) {
        for (int k = 0; k < 6; ++k) step[k] = -y[k];
        // model_cost_change = -(Js step)^T (r + Js step / 2) = -step^T Js^T r - 1/2 step^T Js^T Js step
        double sb = 0, shs = 0;
        for (int a = 0; a < 6; ++a) {
          sb += step[a] * b[a];
          double t = 0;
          for (int c = 0; c < 6; ++c) t += Hs[a][c] * step[c];
          shs += step[a] * t;
        }
        mcc = -sb - 0.5 * shs;
      }
      if (!(ok && mcc > 0.0)) {  // invalid step
        if (++num_invalid >= prm.max_invalid) { trace(cost, 0, gmax, 0, 0, radius, 0, 0); termination = 5; s_go = 0; return; }
        radius *= 0.5;
        reuse_diag = true;
        trace(cost, 0, gmax, 0, 0, radius, 0, 0);
        continue;
      }
      num_invalid = 0;
      double delta[6];
      for (int k = 0; k < 6; ++k) delta[k] = step[k] * scale[k];
      plus7(x, delta, xc);
      for (int k = 0; k < 7; ++k) s_x[k] = xc[k];
      s_go = 1;
      return;
    }
  };

  if (tid == 0) {
    for (int k = 0; k < 7; ++k) x[k] = s_x[k];
    for (int k = 0; k < 21; ++k) H[k] = s_tot[k];
    for (int k = 0; k < 6; ++k) g[k] = s_tot[21 + k];
    cost = s_tot[27];
    summary->initial_cost = cost;
    summary->n_edge = s_cnt[0]; summary->n_plane = s_cnt[1];
    if (n_res == 0) {  // Ceres: nothing to optimise, parameters untouched
      termination = 4; s_go = 0;
    } else {
      for (int j = 0; j < 6; ++j) scale[j] = 1.0 / (1.0 + sqrt(H[tri(j, j)]));
      gmax = gradient_max(x, g);
      x_norm = 0; for (int k = 0; k < 7; ++k) x_norm += x[k] * x[k]; x_norm = sqrt(x_norm);
      trace(cost, 0, gmax, 0, 0, radius, 0, 0);
      if (gmax <= prm.gradient_tolerance) { termination = 1; s_go = 0; }
      else next_candidate();
    }
  }
  __syncthreads();

  while (s_go) {
    evaluate();
    if (tid == 0) {
      ++num_evals;
      const double cand_cost = s_tot[27];
      double sn = 0;
      for (int k = 0; k < 7; ++k) sn += (x[k] - xc[k]) * (x[k] - xc[k]);
      sn = sqrt(sn);
      const double cost_change = cost - cand_cost;
      if (sn <= prm.parameter_tolerance * (x_norm + prm.parameter_tolerance)) {
        trace(cost, 0, gmax, sn, 0, radius, 1, 0); termination = 2; s_go = 0;
      } else if (fabs(cost_change) <= prm.function_tolerance * cost) {
        trace(cost, cost_change, gmax, sn, 0, radius, 1, 0); termination = 3; s_go = 0;
      } else {
        const double rho = cost_change / mcc;
        if (rho > prm.min_relative_decrease) {
          for (int k = 0; k < 7; ++k) x[k] = xc[k];
          x_norm = 0; for (int k = 0; k < 7; ++k) x_norm += x[k] * x[k]; x_norm = sqrt(x_norm);
          for (int k = 0; k < 21; ++k) H[k] = s_tot[k];
          for (int k = 0; k < 6; ++k) g[k] = s_tot[21 + k];
          cost = cand_cost;
          gmax = gradient_max(x, g);
          last_successful = true;
          ++num_successful;
          radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3.0));
          radius = fmin(prm.max_radius, radius);
          decrease_factor = 2.0;
          reuse_diag = false;
          trace(cost, cost_change, gmax, sn, rho, radius, 1, 1);
        } else {
          radius = radius / decrease_factor;
          decrease_factor *= 2.0;
          reuse_diag = true;
          trace(cost, cost_change, gmax, sn, rho, radius, 1, 0);
        }
        next_candidate();
      }
    }
    __syncthreads();
  }

  if (tid == 0) {
    for (int k = 0; k < 7; ++k) x7[k] = x[k];
    summary->termination = termination;
    summary->num_iterations = iteration;
    summary->num_successful = num_successful;
    summary->num_jac_evals = num_evals;
    summary->final_cost = cost;
    summary->trace_rows = trace_rows;
    if (integrate && world7) {
      // laserOdometry.cpp:504-505  t_w += q_w * t_last_curr ; q_w = q_w * q_last_curr
      const V3 u{world7[0], world7[1], world7[2]};
      const double w = world7[3];
      const V3 v{x[4], x[5], x[6]};
      V3 uv = crossd(u, v);
      uv.x += uv.x; uv.y += uv.y; uv.z += uv.z;
      const V3 c2 = crossd(u, uv);
      world7[4] += v.x + w * uv.x + c2.x;
      world7[5] += v.y + w * uv.y + c2.y;
      world7[6] += v.z + w * uv.z + c2.z;
      const double ax = world7[0], ay = world7[1], az = world7[2], aw = world7[3];
      const double bx = x[0], by = x[1], bz = x[2], bw = x[3];
      world7[3] = aw * bw - ax * bx - ay * by - az * bz;
      world7[0] = aw * bx + ax * bw + ay * bz - az * by;
      world7[1] = aw * by + ay * bw + az * bx - ax * bz;
      world7[2] = aw * bz + az * bw + ax * by - ay * bx;
    }
  }
}

}  // namespace aloam
