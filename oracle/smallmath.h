// ORACLE (test infrastructure, NOT product code) -- third-party semantics, restated (not checkable here): the reference ships no tests or
// golden vectors and Ceres/PCL/Eigen are absent here, so these are restatements from the published
// semantics of those libraries.
//
// Tiny fixed-size linear algebra restating the slices of Eigen / Ceres the A-LOAM hot path uses:
//   * Eigen::Quaternion product, q*v (Eigen/src/Geometry/Quaternion.h _transformVector),
//     slerp (QuaternionBase::slerp)                      -- used at lidarFactor.hpp:27-33,79-85,
//                                                           laserOdometry.cpp:120-123,504-505
//   * ceres::Jet<double,7>                               -- AutoDiffCostFunction<.,.,4,3>, lidarFactor.hpp:48,96,130
//   * ceres::EigenQuaternionParameterization Plus / ComputeJacobian
//   * ceres::HuberLoss + Corrector
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

namespace orc {

// ---------------------------------------------------------------- Jet<double, N> (ceres/jet.h)
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT
  Jet(double x, int k) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  // ceres/jet.h: g_a_inverse = 1/g.a ; f_a_by_g_a = f.a * g_a_inverse ; (f.v - f_a_by_g_a * g.v) * g_a_inverse
  Jet<N> h; const double gi = 1.0 / g.a; const double fg = f.a * gi; h.a = fg;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
  return h;
}
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { Jet<N> h; h.a = s * f.a; for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { return s * f; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { return f + s; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> h = -f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline bool operator<(const Jet<N>& f, const Jet<N>& g) { return f.a < g.a; }
template <int N> inline bool operator>=(const Jet<N>& f, const Jet<N>& g) { return f.a >= g.a; }
template <int N> inline bool operator<(const Jet<N>& f, double g) { return f.a < g; }
template <int N> inline Jet<N> sqrt(const Jet<N>& f) { Jet<N> h; h.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <int N> inline Jet<N> sin(const Jet<N>& f) { Jet<N> h; h.a = std::sin(f.a); const double c = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> inline Jet<N> acos(const Jet<N>& f) { Jet<N> h; h.a = std::acos(f.a); const double t = -1.0 / std::sqrt(1.0 - f.a * f.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <int N> inline Jet<N> abs(const Jet<N>& f) { return f.a < 0.0 ? -f : f; }
inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double acos(double x) { return std::acos(x); }
inline double abs(double x) { return std::fabs(x); }

// ---------------------------------------------------------------- Eigen-like quaternion, templated on scalar
template <typename T> struct Vec3T { T x, y, z; };
template <typename T> struct QuatT { T x, y, z, w; };

template <typename T> inline Vec3T<T> cross(const Vec3T<T>& a, const Vec3T<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <typename T> inline T dot(const Vec3T<T>& a, const Vec3T<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> inline Vec3T<T> operator+(const Vec3T<T>& a, const Vec3T<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> inline Vec3T<T> operator-(const Vec3T<T>& a, const Vec3T<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> inline Vec3T<T> operator*(const T& s, const Vec3T<T>& a) { return {s * a.x, s * a.y, s * a.z}; }

// Eigen QuaternionBase::_transformVector:  uv = q.vec x v; uv += uv; v + w*uv + q.vec x uv
template <typename T> inline Vec3T<T> rotate(const QuatT<T>& q, const Vec3T<T>& v) {
  Vec3T<T> u{q.x, q.y, q.z};
  Vec3T<T> uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}
// Eigen quaternion product a*b
template <typename T> inline QuatT<T> qmul(const QuatT<T>& a, const QuatT<T>& b) {
  QuatT<T> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
template <typename T> inline QuatT<T> qconj(const QuatT<T>& a) { return {-a.x, -a.y, -a.z, a.w}; }
// Eigen Quaternion::inverse(): conjugate / squaredNorm
template <typename T> inline QuatT<T> qinv(const QuatT<T>& a) {
  T n2 = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  return {-a.x / n2, -a.y / n2, -a.z / n2, a.w / n2};
}
// Eigen QuaternionBase::slerp(t, other), called on *this = a
template <typename T> inline QuatT<T> slerp(const QuatT<T>& a, const T& t, const QuatT<T>& b) {
  const T one = T(1.0) - T(std::numeric_limits<double>::epsilon());
  T d = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;  // Eigen: coeffs().dot(other.coeffs()) (x,y,z,w order)
  T absD = abs(d);
  T scale0, scale1;
  if (absD >= one) {
    scale0 = T(1.0) - t;
    scale1 = t;
  } else {
    T theta = acos(absD);
    T sinTheta = sin(theta);
    scale0 = sin((T(1.0) - t) * theta) / sinTheta;
    scale1 = sin((t * theta)) / sinTheta;
  }
  if (d < T(0.0)) scale1 = -scale1;
  return {scale0 * a.x + scale1 * b.x, scale0 * a.y + scale1 * b.y, scale0 * a.z + scale1 * b.z,
          scale0 * a.w + scale1 * b.w};
}

typedef Vec3T<double> Vec3;
typedef QuatT<double> Quat;

// ---------------------------------------------------------------- ceres::EigenQuaternionParameterization
// x = (qx,qy,qz,qw).  Plus: delta_q = (sin|d|/|d| * d, cos|d|) ; x+ = delta_q * x.
inline void quat_plus(const double x[4], const double delta[3], double out[4]) {
  const double n = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (n > 0.0) {
    const double s = std::sin(n) / n;
    Quat dq{s * delta[0], s * delta[1], s * delta[2], std::cos(n)};
    Quat q{x[0], x[1], x[2], x[3]};
    Quat r = qmul(dq, q);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  } else {
    for (int i = 0; i < 4; ++i) out[i] = x[i];
  }
}
// 4x3 row-major
inline void quat_plus_jacobian(const double x[4], double J[12]) {
  J[0] = x[3];  J[1] = x[2];   J[2] = -x[1];
  J[3] = -x[2]; J[4] = x[3];   J[5] = x[0];
  J[6] = x[1];  J[7] = -x[0];  J[8] = x[3];
  J[9] = -x[0]; J[10] = -x[1]; J[11] = -x[2];
}

// ---------------------------------------------------------------- ceres::HuberLoss(a) ; rho[0..2]
inline void huber(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else {
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}

}  // namespace orc
