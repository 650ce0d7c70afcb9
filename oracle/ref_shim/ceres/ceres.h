// Stand-in for <ceres/ceres.h> (test infrastructure, our code): dual numbers (oracle/smallmath.h Jet) and an
// AutoDiffCostFunction that can be EVALUATED (residuals + Jacobians by forward-mode autodiff over the functor's own
// operator()).  Problem / Solve collect the residual blocks; the minimiser is the restatement in oracle/lm.cc
// (DESIGN.md section 2, rows 1-8).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>
#include "oracle.h"
#include "smallmath.h"

namespace ceres {

template <typename T, int N> using JetBase = orc::Jet<N>;
typedef orc::Jet<7> Jet7;

namespace shim {
// A driver that wants ceres::Solve to work declares, BEFORE it includes the reference source,
//   bool ref_block_of(const LidarEdgeFactor&, orc::ResidualBlock*);   (and the other functors)
// which packs a functor's public members into the residual-block record of oracle/lm.cc.  Without such a declaration a cost
// function can only be evaluated.
template <typename F> auto pack_block(const F& f, orc::ResidualBlock* b, int) -> decltype(ref_block_of(f, b)) { return ref_block_of(f, b); }
template <typename F> bool pack_block(const F&, orc::ResidualBlock*, long) { return false; }
}  // namespace shim

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual int num_residuals() const = 0;
  virtual bool as_block(orc::ResidualBlock* b) const = 0;
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
  bool as_block(orc::ResidualBlock* b) const override { return shim::pack_block(*functor_, b, 0); }
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

// ---- the slice of the problem / solver API the reference calls (laserOdometry.cpp:283-499, laserMapping.cpp:571-720)
class LossFunction { public: virtual ~LossFunction() {} virtual double a() const = 0; };
class HuberLoss : public LossFunction { public: explicit HuberLoss(double a) : a_(a) {} double a() const override { return a_; } private: double a_; };
class LocalParameterization { public: virtual ~LocalParameterization() {} };
class EigenQuaternionParameterization : public LocalParameterization {};
enum LinearSolverType { DENSE_QR = 0 };

class Problem {
 public:
  struct Options {};
  Problem() {}
  explicit Problem(const Options&) {}
  ~Problem() { for (auto* c : costs_) delete c; for (auto* l : owned_loss_) delete l; for (auto* p : owned_param_) delete p; }
  void AddParameterBlock(double* values, int size, LocalParameterization* p = nullptr) {
    if (size == 4) q_ = values; else if (size == 3) t_ = values;
    if (p) { bool seen = false; for (auto* o : owned_param_) seen = seen || o == p; if (!seen) owned_param_.push_back(p); }
  }
  void AddResidualBlock(CostFunction* c, LossFunction* loss, double* q, double* t) {
    costs_.push_back(c); q_ = q; t_ = t;
    if (loss) { huber_a_ = loss->a(); bool seen = false; for (auto* o : owned_loss_) seen = seen || o == loss; if (!seen) owned_loss_.push_back(loss); }
  }
  std::vector<CostFunction*> costs_;
  std::vector<LossFunction*> owned_loss_;
  std::vector<LocalParameterization*> owned_param_;
  double* q_ = nullptr; double* t_ = nullptr;
  double huber_a_ = 0.1;
};

class Solver {
 public:
  struct Options {
    LinearSolverType linear_solver_type = DENSE_QR;
    int max_num_iterations = 50;
    bool minimizer_progress_to_stdout = false;
    bool check_gradients = false;
    double gradient_check_relative_precision = 1e-8;
  };
  struct Summary { orc::SolveSummary detail; std::string BriefReport() const { return std::string(); } };
};

// the minimiser is oracle/lm.cc (Ceres' trust-region loop is third-party: DESIGN.md section 2, rows 1-8)
inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  std::vector<orc::ResidualBlock> blocks(problem->costs_.size());
  for (size_t i = 0; i < blocks.size(); ++i)
    if (!problem->costs_[i]->as_block(&blocks[i])) { std::fprintf(stderr, "ref_shim: cost function without ref_block_of()\n"); std::abort(); }
  double x[7] = {0, 0, 0, 1, 0, 0, 0};
  if (problem->q_) for (int k = 0; k < 4; ++k) x[k] = problem->q_[k];
  if (problem->t_) for (int k = 0; k < 3; ++k) x[4 + k] = problem->t_[k];
  orc::SolveOptions opt;
  opt.max_num_iterations = options.max_num_iterations;
  opt.huber_a = problem->huber_a_;
  orc::solve(blocks, x, opt, summary ? &summary->detail : nullptr);
  if (problem->q_) for (int k = 0; k < 4; ++k) problem->q_[k] = x[k];
  if (problem->t_) for (int k = 0; k < 3; ++k) problem->t_[k] = x[4 + k];
}

}  // namespace ceres
