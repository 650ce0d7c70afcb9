// Stand-in for <ceres/ceres.h> (test infrastructure, our code): dual numbers (oracle/smallmath.h Jet) and an
// AutoDiffCostFunction that can be EVALUATED (residuals + Jacobians by forward-mode autodiff over the functor's own
// operator()).  Problem / Solve collect the residual blocks; the minimiser is the restatement in oracle/lm.cc
// (DESIGN.md section 2, rows 1-8).
#pragma once
#include <memory>
#include <vector>
#include "smallmath.h"

namespace ceres {

template <typename T, int N> using JetBase = orc::Jet<N>;
typedef orc::Jet<7> Jet7;

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual int num_residuals() const = 0;
  // parameters: q (x, y, z, w) and t ; jacobians (may be null): rows x 4 and rows x 3, row-major, w.r.t. the AMBIENT parameters
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
};

template <typename Functor, int kNumResiduals, int N0, int N1>
class AutoDiffCostFunction : public CostFunction {
  static_assert(N0 == 4 && N1 == 3, "the reference only uses <., ., 4, 3>");
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) {}
  int num_residuals() const override { return kNumResiduals; }
  const Functor& functor() const { return *functor_; }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (!jacobians) return (*functor_)(parameters[0], parameters[1], residuals);
    Jet7 q[4], t[3], r[kNumResiduals];
    for (int i = 0; i < 4; ++i) { q[i] = Jet7(parameters[0][i]); q[i].v[i] = 1.0; }
    for (int i = 0; i < 3; ++i) { t[i] = Jet7(parameters[1][i]); t[i].v[4 + i] = 1.0; }
    if (!(*functor_)(q, t, r)) return false;
    for (int k = 0; k < kNumResiduals; ++k) {
      residuals[k] = r[k].a;
      if (jacobians[0]) for (int i = 0; i < 4; ++i) jacobians[0][k * 4 + i] = r[k].v[i];
      if (jacobians[1]) for (int i = 0; i < 3; ++i) jacobians[1][k * 3 + i] = r[k].v[4 + i];
    }
    return true;
  }
 private:
  std::unique_ptr<Functor> functor_;
};

}  // namespace ceres
