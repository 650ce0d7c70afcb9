// ORACLE (test infrastructure) -- the three residual functors of src/lidarFactor.hpp (:12-55 LidarEdgeFactor,
// :57-104 LidarPlaneFactor, :106-138 LidarPlaneNormFactor) and a restatement of what ceres::Solve does with them
// at laserOdometry.cpp:284-291,494-499 and laserMapping.cpp:565-572,712-720:
//   AutoDiffCostFunction<F,rows,4,3> (Jet<double,7>), HuberLoss(0.1) + Corrector, EigenQuaternionParameterization,
//   TrustRegionMinimizer + LevenbergMarquardtStrategy + DENSE_QR, max_num_iterations = 4, all other options default.
// Ceres is NOT in /root/reference (un-vendored; 1.12.0 pinned only by docker/Dockerfile:3): the control flow below
// follows Ceres' published trust_region_minimizer.cc / levenberg_marquardt_strategy.cc / residual_block.cc /
// corrector.cc / loss_function.cc / local_parameterization.cc (SURVEY.md 8a "R7 spec").  The functors are pinned to the reference's lidarFactor.hpp source; the minimiser is a restatement of Ceres that cannot be checked here.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include "oracle.h"

namespace orc {

namespace {
template <typename T> inline T vnorm(const Vec3T<T>& v) { return sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }

// lidarFactor.hpp:19-43
template <typename T> void edge_functor(const ResidualBlock& rb, const T* q, const T* t, T* residual) {
  Vec3T<T> cp{T(rb.cp[0]), T(rb.cp[1]), T(rb.cp[2])};
  Vec3T<T> lpa{T(rb.a[0]), T(rb.a[1]), T(rb.a[2])};
  Vec3T<T> lpb{T(rb.b[0]), T(rb.b[1]), T(rb.b[2])};
  QuatT<T> q_last_curr{q[0], q[1], q[2], q[3]};
  QuatT<T> q_identity{T(0.0), T(0.0), T(0.0), T(1.0)};
  q_last_curr = slerp(q_identity, T(rb.s), q_last_curr);
  Vec3T<T> t_last_curr{T(rb.s) * t[0], T(rb.s) * t[1], T(rb.s) * t[2]};
  Vec3T<T> lp = rotate(q_last_curr, cp) + t_last_curr;
  Vec3T<T> nu = cross(lp - lpa, lp - lpb);
  Vec3T<T> de = lpa - lpb;
  residual[0] = nu.x / vnorm(de);
  residual[1] = nu.y / vnorm(de);
  residual[2] = nu.z / vnorm(de);
}
// lidarFactor.hpp:69-90 ; ljm_norm precomputed in the constructor (:64-65)
template <typename T> void plane_functor(const ResidualBlock& rb, const T* q, const T* t, T* residual) {
  Vec3T<T> cp{T(rb.cp[0]), T(rb.cp[1]), T(rb.cp[2])};
  Vec3T<T> lpj{T(rb.a[0]), T(rb.a[1]), T(rb.a[2])};
  Vec3T<T> ljm{T(rb.b[0]), T(rb.b[1]), T(rb.b[2])};
  QuatT<T> q_last_curr{q[0], q[1], q[2], q[3]};
  QuatT<T> q_identity{T(0.0), T(0.0), T(0.0), T(1.0)};
  q_last_curr = slerp(q_identity, T(rb.s), q_last_curr);
  Vec3T<T> t_last_curr{T(rb.s) * t[0], T(rb.s) * t[1], T(rb.s) * t[2]};
  Vec3T<T> lp = rotate(q_last_curr, cp) + t_last_curr;
  residual[0] = dot(lp - lpj, ljm);
}
// lidarFactor.hpp:114-125
template <typename T> void plane_norm_functor(const ResidualBlock& rb, const T* q, const T* t, T* residual) {
  QuatT<T> q_w_curr{q[0], q[1], q[2], q[3]};
  Vec3T<T> t_w_curr{t[0], t[1], t[2]};
  Vec3T<T> cp{T(rb.cp[0]), T(rb.cp[1]), T(rb.cp[2])};
  Vec3T<T> point_w = rotate(q_w_curr, cp) + t_w_curr;
  Vec3T<T> norm{T(rb.a[0]), T(rb.a[1]), T(rb.a[2])};
  residual[0] = dot(norm, point_w) + T(rb.s);
}
template <typename T> void functor(const ResidualBlock& rb, const T* q, const T* t, T* residual) {
  if (rb.type == FACTOR_EDGE) edge_functor(rb, q, t, residual);
  else if (rb.type == FACTOR_PLANE) plane_functor(rb, q, t, residual);
  else plane_norm_functor(rb, q, t, residual);
}

// closed-form tangent Jacobian (SURVEY.md 8a "Residual math"): lp = R p + s t (s = 1 for every block the
// reference builds), d lp / d dtheta = -2 [R p]x , d lp / d t = I ; d r / d lp = [b-a]x / |a-b| (edge) or n^T.
void analytic_block(const ResidualBlock& rb, const double x[7], double* r, double* Jloc /*rows x 6*/) {
  Quat q{x[0], x[1], x[2], x[3]};
  Vec3 cp{rb.cp[0], rb.cp[1], rb.cp[2]};
  Vec3 Rp = rotate(q, cp);
  Vec3 lp{Rp.x + x[4], Rp.y + x[5], Rp.z + x[6]};
  // dlp/ddelta = -2 [Rp]x
  const double D[9] = {0, 2 * Rp.z, -2 * Rp.y, -2 * Rp.z, 0, 2 * Rp.x, 2 * Rp.y, -2 * Rp.x, 0};
  if (rb.type == FACTOR_EDGE) {
    Vec3 a{rb.a[0], rb.a[1], rb.a[2]}, b{rb.b[0], rb.b[1], rb.b[2]};
    Vec3 nu = cross(lp - a, lp - b);
    Vec3 de = a - b;
    const double n = vnorm(de);
    r[0] = nu.x / n; r[1] = nu.y / n; r[2] = nu.z / n;
    Vec3 w{(b.x - a.x) / n, (b.y - a.y) / n, (b.z - a.z) / n};
    const double W[9] = {0, -w.z, w.y, w.z, 0, -w.x, -w.y, w.x, 0};  // [w]x
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += W[3 * i + k] * D[3 * k + j];
        Jloc[6 * i + j] = s;
        Jloc[6 * i + 3 + j] = W[3 * i + j];
      }
  } else {
    Vec3 nrm;
    if (rb.type == FACTOR_PLANE) {
      nrm = Vec3{rb.b[0], rb.b[1], rb.b[2]};
      Vec3 j{rb.a[0], rb.a[1], rb.a[2]};
      r[0] = dot(lp - j, nrm);
    } else {
      nrm = Vec3{rb.a[0], rb.a[1], rb.a[2]};
      r[0] = dot(nrm, lp) + rb.s;
    }
    const double nv[3] = {nrm.x, nrm.y, nrm.z};
    for (int j = 0; j < 3; ++j) {
      Jloc[j] = nv[0] * D[j] + nv[1] * D[3 + j] + nv[2] * D[6 + j];
      Jloc[3 + j] = nv[j];
    }
  }
}
}  // namespace

ResidualBlock make_edge(const double cp[3], const double a[3], const double b[3], double s) {
  ResidualBlock rb; rb.type = FACTOR_EDGE; rb.s = s;
  for (int i = 0; i < 3; ++i) { rb.cp[i] = cp[i]; rb.a[i] = a[i]; rb.b[i] = b[i]; }
  return rb;
}
ResidualBlock make_plane(const double cp[3], const double j[3], const double l[3], const double m[3], double s) {
  ResidualBlock rb; rb.type = FACTOR_PLANE; rb.s = s;
  Vec3 vj{j[0], j[1], j[2]}, vl{l[0], l[1], l[2]}, vm{m[0], m[1], m[2]};
  Vec3 n = cross(vj - vl, vj - vm);             // lidarFactor.hpp:64
  const double z = n.x * n.x + n.y * n.y + n.z * n.z;  // Eigen normalize(): /= sqrt(squaredNorm) if > 0
  if (z > 0) { const double nn = std::sqrt(z); n.x /= nn; n.y /= nn; n.z /= nn; }
  for (int i = 0; i < 3; ++i) { rb.cp[i] = cp[i]; rb.a[i] = j[i]; }
  rb.b[0] = n.x; rb.b[1] = n.y; rb.b[2] = n.z;
  return rb;
}
ResidualBlock make_plane_norm(const double cp[3], const double n[3], double d) {
  ResidualBlock rb; rb.type = FACTOR_PLANE_NORM; rb.s = d;
  for (int i = 0; i < 3; ++i) { rb.cp[i] = cp[i]; rb.a[i] = n[i]; rb.b[i] = 0; }
  return rb;
}

double evaluate(const std::vector<ResidualBlock>& blocks, const double x[7], double huber_a, bool autodiff,
                std::vector<double>* residuals, std::vector<double>* jacobian, double* gradient) {
  int rows = 0;
  for (const ResidualBlock& rb : blocks) rows += rb.rows();
  if (residuals) residuals->assign(rows, 0.0);
  if (jacobian) jacobian->assign((size_t)rows * 6, 0.0);
  if (gradient) for (int i = 0; i < 6; ++i) gradient[i] = 0.0;
  const bool want_j = jacobian != nullptr || gradient != nullptr;
  double Jp[12];
  quat_plus_jacobian(x, Jp);
  double cost = 0.0;
  int row = 0;
  for (const ResidualBlock& rb : blocks) {
    const int nr = rb.rows();
    double r[3];
    double Jl[18];
    if (!want_j) {
      functor<double>(rb, x, x + 4, r);  // AutoDiffCostFunction::Evaluate with jacobians == NULL: plain doubles
    } else if (autodiff || (rb.type != FACTOR_PLANE_NORM && rb.s != 1.0)) {   // the closed form below is written for s == 1
      typedef Jet<7> J7;
      J7 q[4], t[3], rr[3];
      for (int i = 0; i < 4; ++i) q[i] = J7(x[i], i);
      for (int i = 0; i < 3; ++i) t[i] = J7(x[4 + i], 4 + i);
      functor<J7>(rb, q, t, rr);
      for (int k = 0; k < nr; ++k) {
        r[k] = rr[k].a;
        // ResidualBlock::Evaluate: local = global(rows x 4) * ComputeJacobian(4 x 3)
        for (int j = 0; j < 3; ++j) {
          double s = 0;
          for (int c = 0; c < 4; ++c) s += rr[k].v[c] * Jp[3 * c + j];
          Jl[6 * k + j] = s;
          Jl[6 * k + 3 + j] = rr[k].v[4 + j];
        }
      }
    } else {
      analytic_block(rb, x, r, Jl);
    }
    double sq = 0;
    for (int k = 0; k < nr; ++k) sq += r[k] * r[k];
    double rho[3];
    huber(huber_a, sq, rho);
    cost += 0.5 * rho[0];
    if (residuals || want_j) {
      // Corrector: rho'' <= 0 for Huber => residual_scaling = sqrt(rho'), jacobian *= sqrt(rho')
      const double sr = std::sqrt(rho[1]);
      for (int k = 0; k < nr; ++k) {
        if (want_j)
          for (int j = 0; j < 6; ++j) Jl[6 * k + j] *= sr;
        r[k] *= sr;
      }
      for (int k = 0; k < nr; ++k) {
        if (residuals) (*residuals)[row + k] = r[k];
        if (jacobian)
          for (int j = 0; j < 6; ++j) (*jacobian)[(size_t)(row + k) * 6 + j] = Jl[6 * k + j];
        if (gradient)
          for (int j = 0; j < 6; ++j) gradient[j] += Jl[6 * k + j] * r[k];
      }
    }
    row += nr;
  }
  return cost;
}

double normal_equations(const std::vector<ResidualBlock>& blocks, const double x[7], double huber_a,
                        bool autodiff, double JtJ[36], double Jtr[6]) {
  std::vector<double> r, J;
  double cost = evaluate(blocks, x, huber_a, autodiff, &r, &J, Jtr);
  for (int i = 0; i < 36; ++i) JtJ[i] = 0;
  for (size_t k = 0; k < r.size(); ++k)
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) JtJ[6 * i + j] += J[6 * k + i] * J[6 * k + j];
  return cost;
}

namespace {
// min || A y - b ||, A m x 6 row-major (destroyed), b length m (destroyed): Householder QR + back substitution
bool householder_lsq(std::vector<double>& A, std::vector<double>& b, int m, double y[6]) {
  const int n = 6;
  for (int k = 0; k < n; ++k) {
    double norm2 = 0;
    for (int i = k; i < m; ++i) norm2 += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    double nrm = std::sqrt(norm2);
    if (nrm == 0.0) return false;
    double akk = A[(size_t)k * n + k];
    double alpha = akk > 0 ? -nrm : nrm;
    // v = x - alpha e1 (stored in column k, rows k..m)
    double v0 = akk - alpha;
    A[(size_t)k * n + k] = v0;
    double vnorm2 = norm2 - akk * akk + v0 * v0;
    if (vnorm2 == 0.0) { A[(size_t)k * n + k] = alpha; continue; }
    for (int j = k + 1; j < n; ++j) {
      double s = 0;
      for (int i = k; i < m; ++i) s += A[(size_t)i * n + k] * A[(size_t)i * n + j];
      s = 2.0 * s / vnorm2;
      for (int i = k; i < m; ++i) A[(size_t)i * n + j] -= s * A[(size_t)i * n + k];
    }
    double s = 0;
    for (int i = k; i < m; ++i) s += A[(size_t)i * n + k] * b[i];
    s = 2.0 * s / vnorm2;
    for (int i = k; i < m; ++i) b[i] -= s * A[(size_t)i * n + k];
    A[(size_t)k * n + k] = alpha;  // R(k,k)
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[(size_t)k * n + j] * y[j];
    if (A[(size_t)k * n + k] == 0.0) return false;
    y[k] = s / A[(size_t)k * n + k];
  }
  for (int k = 0; k < n; ++k)
    if (!std::isfinite(y[k])) return false;
  return true;
}

void plus7(const double x[7], const double delta[6], double out[7]) {
  quat_plus(x, delta, out);
  for (int i = 0; i < 3; ++i) out[4 + i] = x[4 + i] + delta[3 + i];
}
double norm7(const double x[7]) { double s = 0; for (int i = 0; i < 7; ++i) s += x[i] * x[i]; return std::sqrt(s); }
}  // namespace

void solve(const std::vector<ResidualBlock>& blocks, double x[7], const SolveOptions& opt, SolveSummary* summary) {
  SolveSummary local;
  SolveSummary& S = summary ? *summary : local;
  S = SolveSummary();
  if (blocks.empty()) { S.termination = 4; return; }  // preprocessor: nothing to optimise, parameters untouched

  std::vector<double> r, J;
  double g[6];
  double cost = evaluate(blocks, x, opt.huber_a, opt.autodiff, &r, &J, g);
  S.num_jacobian_evals = 1;
  S.initial_cost = S.final_cost = cost;
  const int m = (int)r.size();

  double scale[6];
  for (int j = 0; j < 6; ++j) scale[j] = 1.0;
  if (opt.jacobi_scaling) {  // computed once at iteration 0: 1 / (1 + sqrt(squared column norm))
    for (int j = 0; j < 6; ++j) {
      double s = 0;
      for (int i = 0; i < m; ++i) s += J[(size_t)i * 6 + j] * J[(size_t)i * 6 + j];
      scale[j] = 1.0 / (1.0 + std::sqrt(s));
    }
  }
  auto scale_columns = [&]() {
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < 6; ++j) J[(size_t)i * 6 + j] *= scale[j];
  };
  scale_columns();
  auto gradient_max_norm = [&]() {
    double ng[6], xp[7];
    for (int j = 0; j < 6; ++j) ng[j] = -g[j];
    plus7(x, ng, xp);
    double mx = 0;
    for (int i = 0; i < 7; ++i) mx = std::max(mx, std::fabs(xp[i] - x[i]));
    return mx;
  };
  double gmax = gradient_max_norm();
  double x_norm = norm7(x);

  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  double diag[6];
  int num_invalid = 0;
  bool last_successful = false;
  S.iters.push_back({cost, 0, gmax, 0, 0, radius, 0, 0});
  if (gmax <= opt.gradient_tolerance) { S.termination = 1; return; }

  int iteration = 0;
  std::vector<double> A, rhs, model(m);
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iteration >= opt.max_num_iterations) { S.termination = 0; break; }
    if (last_successful && gmax <= opt.gradient_tolerance) { S.termination = 1; break; }
    if (radius < opt.min_trust_region_radius) { S.termination = 5; break; }
    ++iteration;
    S.num_iterations = iteration;
    last_successful = false;
    IterationRecord rec{cost, 0, gmax, 0, 0, radius, 0, 0};

    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += J[(size_t)i * 6 + j] * J[(size_t)i * 6 + j];
        diag[j] = std::min(std::max(s, opt.min_lm_diagonal), opt.max_lm_diagonal);
      }
    }
    double lm_diag[6];
    for (int j = 0; j < 6; ++j) lm_diag[j] = std::sqrt(diag[j] / radius);
    // DenseQRSolver: [J ; diag(D)] y = [r ; 0]  (solves J y = r, step = -y)
    A.assign((size_t)(m + 6) * 6, 0.0);
    rhs.assign(m + 6, 0.0);
    std::memcpy(A.data(), J.data(), sizeof(double) * (size_t)m * 6);
    for (int j = 0; j < 6; ++j) A[(size_t)(m + j) * 6 + j] = lm_diag[j];
    std::memcpy(rhs.data(), r.data(), sizeof(double) * m);
    double y[6], step[6];
    bool ok = householder_lsq(A, rhs, m + 6, y);
    reuse_diagonal = true;
    double model_cost_change = 0;
    if (ok) {
      for (int j = 0; j < 6; ++j) step[j] = -y[j];
      // model_cost_change = -(J step)^T (r + J step / 2)
      for (int i = 0; i < m; ++i) {
        double s = 0;
        for (int j = 0; j < 6; ++j) s += J[(size_t)i * 6 + j] * step[j];
        model[i] = s;
      }
      for (int i = 0; i < m; ++i) model_cost_change -= model[i] * (r[i] + model[i] / 2.0);
    }
    const bool valid = ok && model_cost_change > 0.0;
    rec.valid = valid;
    if (!valid) {  // HandleInvalidStep
      if (++num_invalid >= opt.max_num_consecutive_invalid_steps) { S.iters.push_back(rec); S.termination = 5; break; }
      radius *= 0.5;  // StepIsInvalid
      reuse_diagonal = true;
      rec.radius = radius;
      S.iters.push_back(rec);
      continue;
    }
    num_invalid = 0;
    double delta[6];
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * scale[j];
    double xc[7];
    plus7(x, delta, xc);
    double cand_cost = evaluate(blocks, xc, opt.huber_a, opt.autodiff, nullptr, nullptr, nullptr);
    ++S.num_cost_evals;

    // ParameterToleranceReached
    double sn = 0;
    for (int i = 0; i < 7; ++i) sn += (x[i] - xc[i]) * (x[i] - xc[i]);
    rec.step_norm = std::sqrt(sn);
    if (rec.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
      S.iters.push_back(rec); S.termination = 2; break;
    }
    // FunctionToleranceReached
    rec.cost_change = cost - cand_cost;
    if (std::fabs(rec.cost_change) <= opt.function_tolerance * cost) {
      S.iters.push_back(rec); S.termination = 3; break;
    }
    rec.relative_decrease = rec.cost_change / model_cost_change;
    if (rec.relative_decrease > opt.min_relative_decrease) {  // HandleSuccessfulStep
      for (int i = 0; i < 7; ++i) x[i] = xc[i];
      x_norm = norm7(x);
      cost = evaluate(blocks, x, opt.huber_a, opt.autodiff, &r, &J, g);
      ++S.num_jacobian_evals;
      scale_columns();
      gmax = gradient_max_norm();
      rec.successful = 1;
      last_successful = true;
      ++S.num_successful_steps;
      // StepAccepted
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rec.relative_decrease - 1.0, 3));
      radius = std::min(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      rec.cost = cost; rec.gradient_max_norm = gmax;
    } else {  // HandleUnsuccessfulStep -> StepRejected
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
    rec.radius = radius;
    S.iters.push_back(rec);
  }
  S.final_cost = cost;
}

}  // namespace orc
