// lidarFactor.hpp -- drop-in for the reference's src/lidarFactor.hpp (SURVEY.md 8b "preserved surface 2").
//
// Same struct names, constructor arguments, member names, templated
//     bool operator()(const T* q, const T* t, T* residual) const          (q in Eigen / Ceres order x, y, z, w)
// and block sizes <3,4,3> (LidarEdgeFactor, reference :12-55), <1,4,3> (LidarPlaneFactor, :57-104), <1,4,3>
// (LidarPlaneNormFactor, :106-138) -- but the header itself needs neither Ceres nor Eigen nor PCL:
//   * vectors are anything with .x() .y() .z() (Eigen::Vector3d qualifies; aloam::Vec3d below is the built-in stand-in);
//   * operator() is written on plain scalars with the reference's operation order (Eigen's slerp and q * v restated), so it
//     works for double, for ceres::Jet and for any other scalar type with + - * / sqrt sin acos abs and comparisons;
//   * bool Evaluate(q, t, residual, jacobian) gives the ANALYTIC Jacobian in the 6-dim tangent Ceres optimises in with
//     EigenQuaternionParameterization ([dtheta(3) | dt(3)], row-major rows x 6) -- including the interpolation ratio s != 1
//     of the DISTORTION == 1 build (laserOdometry.cpp:59,115-116,376-379) -- so no autodiff is needed to build J^T J;
//   * PackBlock(double[11]) emits the record aloam_normal_equations / aloam_solve (include/aloam_b200.h) consume, which is
//     how a caller hands its residual blocks to the GPU instead of to ceres::Problem;
//   * the Ceres factories Create(...) of the reference (:45-51,92-99,127-133) exist under #ifdef ALOAM_WITH_CERES, unchanged
//     in signature, for callers that keep Ceres in the loop (define ALOAM_WITH_CERES before including; needs Eigen too).
// LidarDistanceFactor (:141-172) is dead code in the reference (only referenced from comments) and is not provided.
#ifndef ALOAM_LIDAR_FACTOR_HPP_
#define ALOAM_LIDAR_FACTOR_HPP_

#include <cmath>
#include <limits>

#ifdef ALOAM_WITH_CERES
#include <ceres/ceres.h>
#include <eigen3/Eigen/Dense>
#endif

namespace aloam {

// stand-in for Eigen::Vector3d where Eigen is not available (same accessors)
struct Vec3d {
  double v[3];
  Vec3d() : v{0, 0, 0} {}
  Vec3d(double x_, double y_, double z_) : v{x_, y_, z_} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};

#ifdef ALOAM_WITH_CERES
typedef Eigen::Vector3d FactorVec3;
#else
typedef Vec3d FactorVec3;
#endif

namespace factor_detail {

using std::abs;
using std::acos;
using std::sin;
using std::sqrt;

// Eigen QuaternionBase::slerp(s, q) called on the identity: (x, y, z, w) of the interpolated rotation.
// d = identity . q = q.w ; |d| >= 1 - eps -> linear blend, else spherical ; d < 0 flips the second weight.
template <typename T>
inline void slerp_from_identity(const T* q, const T& s, T out[4]) {
  const T one = T(1.0) - T(std::numeric_limits<double>::epsilon());
  const T d = q[3];
  const T ad = abs(d);
  T w0, w1;
  if (ad >= one) {
    w0 = T(1.0) - s;
    w1 = s;
  } else {
    const T th = acos(ad);
    const T sth = sin(th);
    w0 = sin((T(1.0) - s) * th) / sth;
    w1 = sin(s * th) / sth;
  }
  if (d < T(0.0)) w1 = -w1;
  out[0] = w1 * q[0];
  out[1] = w1 * q[1];
  out[2] = w1 * q[2];
  out[3] = w0 + w1 * q[3];
}

// Eigen q * v:  uv = 2 (u x v) ;  v + w uv + u x uv   (no normalisation, as Eigen)
template <typename T>
inline void rotate(const T* q, const T* v, T out[3]) {
  T uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] = uv[0] + uv[0]; uv[1] = uv[1] + uv[1]; uv[2] = uv[2] + uv[2];
  out[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
  out[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
  out[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
}

// lp = slerp(I, q, s) * p + s t   -- the point of the current sweep carried to the sweep start (reference :27-33,79-85)
template <typename T>
inline void point_to_start(const T* q, const T* t, const double p[3], double s, T lp[3]) {
  T qs[4];
  slerp_from_identity(q, T(s), qs);
  const T pv[3] = {T(p[0]), T(p[1]), T(p[2])};
  T r[3];
  rotate(qs, pv, r);
  lp[0] = r[0] + T(s) * t[0];
  lp[1] = r[1] + T(s) * t[1];
  lp[2] = r[2] + T(s) * t[2];
}

// d lp / d(dtheta, dt) for lp = R(q)^s p + s t with Ceres' left-multiplicative Plus(q, dtheta) = Exp(dtheta) (x) q, where
// Exp(d) = (sin|d| d/|d|, cos|d|) is a rotation by 2|d|:
//   rotation vector phi of q (|phi| = th), q^s = Exp_so3(s phi) ;  a left perturbation eps = 2 dtheta of q moves q^s by the left
//   perturbation  M eps,  M = s Jl(s phi) Jl(phi)^-1   (Jl = left Jacobian of SO(3)) ;  hence
//   d lp / d dtheta = -2 [R^s p]x M ,   d lp / d t = s I.
// Jl and its inverse are polynomials in K = [phi / th]x (K^3 = -K), so M = m0 I + m1 K + m2 K^2 with scalar coefficients.
// For s == 1 M is the identity and the expression reduces to SURVEY.md 8a "Residual math".
struct StartJacobian {
  double Rp[3];      // R^s p
  double k[3];       // rotation axis of q (unit; anything when th == 0)
  double m0, m1, m2; // M = m0 I + m1 K + m2 K^2
  double s;
  // row = g^T d lp / d(dtheta, dt) for a residual with gradient g wrt lp
  void row(const double g[3], double out[6]) const {
    // g^T (-2 [Rp]x) = 2 (Rp x g)^T =: h^T ; then h^T M = m0 h + m1 (h x k) ... careful: h^T K = (K^T h)^T = -(k x h)^T = (h x k)^T
    const double h[3] = {2.0 * (Rp[1] * g[2] - Rp[2] * g[1]), 2.0 * (Rp[2] * g[0] - Rp[0] * g[2]), 2.0 * (Rp[0] * g[1] - Rp[1] * g[0])};
    const double hk[3] = {h[1] * k[2] - h[2] * k[1], h[2] * k[0] - h[0] * k[2], h[0] * k[1] - h[1] * k[0]};          // h^T K
    const double hkk[3] = {hk[1] * k[2] - hk[2] * k[1], hk[2] * k[0] - hk[0] * k[2], hk[0] * k[1] - hk[1] * k[0]};   // h^T K^2
    for (int a = 0; a < 3; ++a) {
      out[a] = m0 * h[a] + m1 * hk[a] + m2 * hkk[a];
      out[3 + a] = s * g[a];
    }
  }
};

inline StartJacobian start_jacobian(const double* q, const double p[3], double s) {
  StartJacobian J;
  J.s = s;
  double qs[4];
  slerp_from_identity(q, s, qs);
  rotate(qs, p, J.Rp);
  // rotation vector of q: the shorter arc, as the slerp takes it (q and -q are the same rotation)
  const double sgn = q[3] < 0.0 ? -1.0 : 1.0;
  const double vx = sgn * q[0], vy = sgn * q[1], vz = sgn * q[2], w = sgn * q[3];
  const double vn = std::sqrt(vx * vx + vy * vy + vz * vz);
  const double th = 2.0 * std::atan2(vn, w);
  if (vn < 1e-12 || s == 1.0) {   // M -> s I as th -> 0 (and exactly I for s == 1)
    J.k[0] = 1.0; J.k[1] = 0.0; J.k[2] = 0.0;
    J.m0 = s; J.m1 = 0.0; J.m2 = 0.0;
    return J;
  }
  J.k[0] = vx / vn; J.k[1] = vy / vn; J.k[2] = vz / vn;
  // Jl(s phi) = I + a1 K + a2 K^2 ;  Jl(phi)^-1 = I + b1 K + b2 K^2
  const double u = s * th;
  const double a1 = std::abs(u) < 1e-8 ? 0.5 * u : (1.0 - std::cos(u)) / u;
  const double a2 = std::abs(u) < 1e-4 ? u * u / 6.0 : 1.0 - std::sin(u) / u;
  const double b1 = -0.5 * th;
  const double b2 = 1.0 - 0.5 * th * std::cos(0.5 * th) / std::sin(0.5 * th);
  // (I + a1 K + a2 K^2)(I + b1 K + b2 K^2) with K^3 = -K, K^4 = -K^2
  J.m0 = s;
  J.m1 = s * (b1 + a1 - a1 * b2 - a2 * b1);
  J.m2 = s * (b2 + a2 + a1 * b1 - a2 * b2);
  return J;
}

}  // namespace factor_detail
}  // namespace aloam

// ----------------------------------------------------------------------------------------------------------------------
// point-to-line residual (3 rows): ((lp - a) x (lp - b)) / |a - b|, lp = current point carried to the sweep start
struct LidarEdgeFactor {
  typedef aloam::FactorVec3 Vec;
  template <class V>
  LidarEdgeFactor(const V& curr_point_, const V& last_point_a_, const V& last_point_b_, double s_)
      : curr_point(curr_point_.x(), curr_point_.y(), curr_point_.z()), last_point_a(last_point_a_.x(), last_point_a_.y(), last_point_a_.z()),
        last_point_b(last_point_b_.x(), last_point_b_.y(), last_point_b_.z()), s(s_) {}

  template <typename T>
  bool operator()(const T* q, const T* t, T* residual) const {
    using namespace aloam::factor_detail;
    const double cp[3] = {curr_point.x(), curr_point.y(), curr_point.z()};
    T lp[3];
    point_to_start(q, t, cp, s, lp);
    const T da[3] = {lp[0] - T(last_point_a.x()), lp[1] - T(last_point_a.y()), lp[2] - T(last_point_a.z())};
    const T db[3] = {lp[0] - T(last_point_b.x()), lp[1] - T(last_point_b.y()), lp[2] - T(last_point_b.z())};
    const T nu[3] = {da[1] * db[2] - da[2] * db[1], da[2] * db[0] - da[0] * db[2], da[0] * db[1] - da[1] * db[0]};
    const T de[3] = {T(last_point_a.x()) - T(last_point_b.x()), T(last_point_a.y()) - T(last_point_b.y()), T(last_point_a.z()) - T(last_point_b.z())};
    const T den = sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    residual[0] = nu[0] / den;
    residual[1] = nu[1] / den;
    residual[2] = nu[2] / den;
    return true;
  }

  // residual[3]; jacobian (may be null): 3 x 6 row-major in the tangent [dtheta | dt]
  bool Evaluate(const double* q, const double* t, double* residual, double* jacobian) const {
    (*this)(q, t, residual);
    if (!jacobian) return true;
    const double cp[3] = {curr_point.x(), curr_point.y(), curr_point.z()};
    const aloam::factor_detail::StartJacobian J = aloam::factor_detail::start_jacobian(q, cp, s);
    const double e[3] = {last_point_b.x() - last_point_a.x(), last_point_b.y() - last_point_a.y(), last_point_b.z() - last_point_a.z()};
    const double n = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const double w[3] = {e[0] / n, e[1] / n, e[2] / n};
    // d r / d lp = [w]x : rows (0, -wz, wy), (wz, 0, -wx), (-wy, wx, 0)
    const double g[3][3] = {{0.0, -w[2], w[1]}, {w[2], 0.0, -w[0]}, {-w[1], w[0], 0.0}};
    for (int r = 0; r < 3; ++r) J.row(g[r], jacobian + 6 * r);
    return true;
  }

  // record for aloam_normal_equations / aloam_solve: [type 0, curr(3), a(3), b(3), s]
  void PackBlock(double out[11]) const {
    out[0] = 0.0;
    out[1] = curr_point.x(); out[2] = curr_point.y(); out[3] = curr_point.z();
    out[4] = last_point_a.x(); out[5] = last_point_a.y(); out[6] = last_point_a.z();
    out[7] = last_point_b.x(); out[8] = last_point_b.y(); out[9] = last_point_b.z();
    out[10] = s;
  }

#ifdef ALOAM_WITH_CERES
  static ceres::CostFunction* Create(const Eigen::Vector3d curr_point_, const Eigen::Vector3d last_point_a_, const Eigen::Vector3d last_point_b_,
                                     const double s_) {
    return (new ceres::AutoDiffCostFunction<LidarEdgeFactor, 3, 4, 3>(new LidarEdgeFactor(curr_point_, last_point_a_, last_point_b_, s_)));
  }
#endif

  Vec curr_point, last_point_a, last_point_b;
  double s;
};

// point-to-plane residual (1 row): (lp - j) . n, n = unit normal of the triangle (j, l, m), computed once in the constructor
struct LidarPlaneFactor {
  typedef aloam::FactorVec3 Vec;
  template <class V>
  LidarPlaneFactor(const V& curr_point_, const V& last_point_j_, const V& last_point_l_, const V& last_point_m_, double s_)
      : curr_point(curr_point_.x(), curr_point_.y(), curr_point_.z()), last_point_j(last_point_j_.x(), last_point_j_.y(), last_point_j_.z()),
        last_point_l(last_point_l_.x(), last_point_l_.y(), last_point_l_.z()), last_point_m(last_point_m_.x(), last_point_m_.y(), last_point_m_.z()),
        s(s_) {
    const double jl[3] = {last_point_j.x() - last_point_l.x(), last_point_j.y() - last_point_l.y(), last_point_j.z() - last_point_l.z()};
    const double jm[3] = {last_point_j.x() - last_point_m.x(), last_point_j.y() - last_point_m.y(), last_point_j.z() - last_point_m.z()};
    double n[3] = {jl[1] * jm[2] - jl[2] * jm[1], jl[2] * jm[0] - jl[0] * jm[2], jl[0] * jm[1] - jl[1] * jm[0]};
    const double z = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    if (z > 0.0) { const double nn = std::sqrt(z); n[0] /= nn; n[1] /= nn; n[2] /= nn; }   // Eigen normalize(): only if squaredNorm > 0
    ljm_norm = Vec(n[0], n[1], n[2]);
  }

  template <typename T>
  bool operator()(const T* q, const T* t, T* residual) const {
    using namespace aloam::factor_detail;
    const double cp[3] = {curr_point.x(), curr_point.y(), curr_point.z()};
    T lp[3];
    point_to_start(q, t, cp, s, lp);
    residual[0] = (lp[0] - T(last_point_j.x())) * T(ljm_norm.x()) + (lp[1] - T(last_point_j.y())) * T(ljm_norm.y()) +
                  (lp[2] - T(last_point_j.z())) * T(ljm_norm.z());
    return true;
  }

  // residual[1]; jacobian (may be null): 1 x 6
  bool Evaluate(const double* q, const double* t, double* residual, double* jacobian) const {
    (*this)(q, t, residual);
    if (!jacobian) return true;
    const double cp[3] = {curr_point.x(), curr_point.y(), curr_point.z()};
    const aloam::factor_detail::StartJacobian J = aloam::factor_detail::start_jacobian(q, cp, s);
    const double g[3] = {ljm_norm.x(), ljm_norm.y(), ljm_norm.z()};
    J.row(g, jacobian);
    return true;
  }

  // [type 1, curr(3), j(3), unit normal(3), s]
  void PackBlock(double out[11]) const {
    out[0] = 1.0;
    out[1] = curr_point.x(); out[2] = curr_point.y(); out[3] = curr_point.z();
    out[4] = last_point_j.x(); out[5] = last_point_j.y(); out[6] = last_point_j.z();
    out[7] = ljm_norm.x(); out[8] = ljm_norm.y(); out[9] = ljm_norm.z();
    out[10] = s;
  }

#ifdef ALOAM_WITH_CERES
  static ceres::CostFunction* Create(const Eigen::Vector3d curr_point_, const Eigen::Vector3d last_point_j_, const Eigen::Vector3d last_point_l_,
                                     const Eigen::Vector3d last_point_m_, const double s_) {
    return (new ceres::AutoDiffCostFunction<LidarPlaneFactor, 1, 4, 3>(new LidarPlaneFactor(curr_point_, last_point_j_, last_point_l_, last_point_m_, s_)));
  }
#endif

  Vec curr_point, last_point_j, last_point_l, last_point_m;
  Vec ljm_norm;
  double s;
};

// point-to-fitted-plane residual of the mapping stage (1 row): n . (q p + t) + d, n unit, d = negative_OA_dot_norm
struct LidarPlaneNormFactor {
  typedef aloam::FactorVec3 Vec;
  template <class V>
  LidarPlaneNormFactor(const V& curr_point_, const V& plane_unit_norm_, double negative_OA_dot_norm_)
      : curr_point(curr_point_.x(), curr_point_.y(), curr_point_.z()), plane_unit_norm(plane_unit_norm_.x(), plane_unit_norm_.y(), plane_unit_norm_.z()),
        negative_OA_dot_norm(negative_OA_dot_norm_) {}

  template <typename T>
  bool operator()(const T* q, const T* t, T* residual) const {
    const T cp[3] = {T(curr_point.x()), T(curr_point.y()), T(curr_point.z())};
    T pw[3];
    aloam::factor_detail::rotate(q, cp, pw);
    pw[0] = pw[0] + t[0]; pw[1] = pw[1] + t[1]; pw[2] = pw[2] + t[2];
    residual[0] = T(plane_unit_norm.x()) * pw[0] + T(plane_unit_norm.y()) * pw[1] + T(plane_unit_norm.z()) * pw[2] + T(negative_OA_dot_norm);
    return true;
  }

  bool Evaluate(const double* q, const double* t, double* residual, double* jacobian) const {
    (*this)(q, t, residual);
    if (!jacobian) return true;
    const double cp[3] = {curr_point.x(), curr_point.y(), curr_point.z()};
    aloam::factor_detail::StartJacobian J;
    aloam::factor_detail::rotate(q, cp, J.Rp);
    J.k[0] = 1.0; J.k[1] = 0.0; J.k[2] = 0.0; J.m0 = 1.0; J.m1 = 0.0; J.m2 = 0.0; J.s = 1.0;
    const double g[3] = {plane_unit_norm.x(), plane_unit_norm.y(), plane_unit_norm.z()};
    J.row(g, jacobian);
    return true;
  }

  // [type 2, curr(3), unit normal(3), 0 0 0, negative_OA_dot_norm]
  void PackBlock(double out[11]) const {
    out[0] = 2.0;
    out[1] = curr_point.x(); out[2] = curr_point.y(); out[3] = curr_point.z();
    out[4] = plane_unit_norm.x(); out[5] = plane_unit_norm.y(); out[6] = plane_unit_norm.z();
    out[7] = 0.0; out[8] = 0.0; out[9] = 0.0;
    out[10] = negative_OA_dot_norm;
  }

#ifdef ALOAM_WITH_CERES
  static ceres::CostFunction* Create(const Eigen::Vector3d curr_point_, const Eigen::Vector3d plane_unit_norm_, const double negative_OA_dot_norm_) {
    return (new ceres::AutoDiffCostFunction<LidarPlaneNormFactor, 1, 4, 3>(new LidarPlaneNormFactor(curr_point_, plane_unit_norm_, negative_OA_dot_norm_)));
  }
#endif

  Vec curr_point;
  Vec plane_unit_norm;
  double negative_OA_dot_norm;
};

#endif  // ALOAM_LIDAR_FACTOR_HPP_
