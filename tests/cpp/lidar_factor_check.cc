// Test driver for include/lidarFactor.hpp (compiled by tests/test_lidar_factor_header.py with plain g++, no Ceres / Eigen):
//  * the templated operator() is evaluated on doubles and on the oracle's Jet<7> (oracle/smallmath.h, a restatement of ceres::Jet)
//    = what ceres::AutoDiffCostFunction<F, rows, 4, 3> would do; the 4-column quaternion Jacobian is multiplied by the
//    EigenQuaternionParameterization Jacobian to get the 6-dim tangent Jacobian;
//  * Evaluate()'s analytic Jacobian must agree with it;
//  * every case is written to stdout (packed block, pose, residuals, analytic Jacobian) so that the Python side can compare
//    with the oracle's own functors (liboracle) and feed the blocks to the GPU.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "lidarFactor.hpp"
#include "smallmath.h"

using orc::Jet;

template <class F, int ROWS>
static double check(const F& f, const double x[7], double* res_out, double* jac_out) {
  double r[ROWS], Ja[ROWS * 6];
  f.Evaluate(x, x + 4, r, Ja);
  Jet<7> p[7];
  for (int k = 0; k < 7; ++k) p[k] = Jet<7>(x[k], k);
  Jet<7> rj[ROWS];
  f(p, p + 4, rj);
  double P[12];
  orc::quat_plus_jacobian(x, P);
  double worst = 0;
  for (int i = 0; i < ROWS; ++i) {
    worst = std::max(worst, std::fabs(rj[i].a - r[i]));
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += rj[i].v[k] * P[3 * k + j];
      worst = std::max(worst, std::fabs(s - Ja[6 * i + j]));
      worst = std::max(worst, std::fabs(rj[i].v[4 + j] - Ja[6 * i + 3 + j]));
    }
    res_out[i] = r[i];
    for (int j = 0; j < 6; ++j) jac_out[6 * i + j] = Ja[6 * i + j];
  }
  return worst;
}

int main() {
  std::mt19937_64 rng(20240901);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  const double ratios[] = {1.0, 0.0, 0.25, 0.73, 0.999};
  double worst = 0, worst_tiny = 0;   // tiny rotations (1e-7 rad) with s != 1: the AUTODIFF side loses digits in acos / sin(theta) there
  int cases = 0;
  for (int trial = 0; trial < 60; ++trial) {
    // pose: kinds 0 random rotation up to ~0.6 rad, 1 tiny rotation, 2 exact identity, 3 negated quaternion (q.w < 0)
    const int kind = trial % 4;
    double ax[3] = {U(rng), U(rng), U(rng)};
    const double an = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    const double ang = kind == 1 ? 1e-7 * (1 + U(rng)) : kind == 2 ? 0.0 : 0.6 * std::fabs(U(rng)) + 0.01;
    double x[7] = {std::sin(ang / 2) * ax[0] / an, std::sin(ang / 2) * ax[1] / an, std::sin(ang / 2) * ax[2] / an, std::cos(ang / 2), U(rng), U(rng), 0.3 * U(rng)};
    if (kind == 3) for (int k = 0; k < 4; ++k) x[k] = -x[k];
    for (double s : ratios) {
      aloam::Vec3d cp(20 * U(rng), 20 * U(rng), 3 * U(rng)), a(cp.x() + U(rng), cp.y() + U(rng), cp.z() + U(rng)),
          b(a.x() + 0.5 * U(rng), a.y() + 0.5 * U(rng), a.z() + 2.0 + U(rng)), l(a.x() + 1 + U(rng), a.y() + U(rng), a.z() + 0.2 * U(rng)),
          m(a.x() + 0.3 * U(rng), a.y() + 1.5 + U(rng), a.z() + 0.2 * U(rng));
      LidarEdgeFactor fe(cp, a, b, s);
      LidarPlaneFactor fp(cp, a, l, m, s);
      double nrm[3] = {U(rng), U(rng), 1.5 + U(rng)};
      const double nn = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
      LidarPlaneNormFactor fn(cp, aloam::Vec3d(nrm[0] / nn, nrm[1] / nn, nrm[2] / nn), 1.7 + U(rng));
      double blk[11], r[3], J[18];
      std::printf("X %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", x[0], x[1], x[2], x[3], x[4], x[5], x[6]);
      auto dump = [&](int rows) {
        std::printf("B");
        for (int k = 0; k < 11; ++k) std::printf(" %.17g", blk[k]);
        std::printf("\nR");
        for (int k = 0; k < rows; ++k) std::printf(" %.17g", r[k]);
        std::printf("\nJ");
        for (int k = 0; k < rows * 6; ++k) std::printf(" %.17g", J[k]);
        std::printf("\n");
      };
      double& acc = (kind == 1 && s != 1.0) ? worst_tiny : worst;
      acc = std::max(acc, check<LidarEdgeFactor, 3>(fe, x, r, J)); fe.PackBlock(blk); dump(3);
      acc = std::max(acc, check<LidarPlaneFactor, 1>(fp, x, r, J)); fp.PackBlock(blk); dump(1);
      acc = std::max(acc, check<LidarPlaneNormFactor, 1>(fn, x, r, J)); fn.PackBlock(blk); dump(1);
      cases += 3;
    }
  }
  std::printf("WORST %.3e WORST_TINY %.3e CASES %d\n", worst, worst_tiny, cases);
  return 0;
}
