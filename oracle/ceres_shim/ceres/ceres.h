// TEST INFRASTRUCTURE ONLY (oracle/): never linked into the product library.
//
// Minimal stand-in for <ceres/ceres.h> so that the reference's
// lib/utils/extend_utils/src/uncertainty_pnp.cpp compiles UNMODIFIED, from where
// it lies under /root/reference, into oracle/_ref/libuncertainty_pnp_ref.so.
//
// Why: the vendored libceres.so.{1.14.0,2.0.0} cannot be linked in this image
// (needs libglog/SuiteSparse/LAPACK) and <ceres/ceres.h> needs
// glog headers.  The header-only part of the vendored Ceres 2.0 tree
// (ceres/jet.h, ceres/rotation.h, ceres/tiny_solver.h + Eigen 3) does compile,
// so this shim maps the four Ceres API names the reference file touches
// (AutoDiffCostFunction, Problem, Solver::{Options,Summary}, Solve) onto the
// reference's own vendored ceres::TinySolver (LM with Jacobi scaling, same
// default tolerances).  The cost functor, the extern "C" uncertainty_pnp()
// entry and the solver loop are therefore all reference code; only this glue
// is ours.  Difference to full ceres::Solve: TinySolver accepts
// a step when rho > 0 (full Ceres: rho > 1e-3) and has no function_tolerance
// stop, i.e. it converges at least as far.  Parity against it is asserted on
// the converged minimiser, not on iterates.
#ifndef ESA_POSE_B200_ORACLE_CERES_SHIM_H_
#define ESA_POSE_B200_ORACLE_CERES_SHIM_H_

#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "Eigen/Dense"
#include "ceres/jet.h"
#include "ceres/rotation.h"
#include "ceres/tiny_solver.h"

namespace ceres {

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual int num_residuals() const = 0;
  virtual int num_parameters() const = 0;
  // jac: row-major [num_residuals x num_parameters] or nullptr.
  virtual bool Evaluate(const double* x, double* res, double* jac) const = 0;
};

template <typename Functor, int kNumResiduals, int N0>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) {}
  int num_residuals() const override { return kNumResiduals; }
  int num_parameters() const override { return N0; }
  bool Evaluate(const double* x, double* res, double* jac) const override {
    if (jac == nullptr) return (*functor_)(x, res);
    typedef Jet<double, N0> JetT;
    JetT jx[N0];
    JetT jr[kNumResiduals];
    for (int i = 0; i < N0; ++i) jx[i] = JetT(x[i], i);
    if (!(*functor_)(jx, jr)) return false;
    for (int k = 0; k < kNumResiduals; ++k) {
      res[k] = jr[k].a;
      for (int i = 0; i < N0; ++i) jac[k * N0 + i] = jr[k].v[i];
    }
    return true;
  }

 private:
  std::unique_ptr<Functor> functor_;
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, DENSE_SCHUR };

class Problem {
 public:
  void AddResidualBlock(CostFunction* cost, void* /*loss*/, double* params) {
    if (params_ == nullptr) params_ = params;
    blocks_.emplace_back(cost);
    same_block_ = same_block_ && (params == params_);
  }
  double* params_ = nullptr;
  bool same_block_ = true;
  std::vector<std::unique_ptr<CostFunction>> blocks_;
};

struct Solver {
  struct Options {
    LinearSolverType linear_solver_type = DENSE_QR;
    bool minimizer_progress_to_stdout = false;
    int max_num_iterations = 50;
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
    double initial_trust_region_radius = 1e4;
  };
  struct Summary {
    double initial_cost = -1, final_cost = -1;
    int iterations = -1;
    int status = -1;
    std::string FullReport() const {
      return "tiny_solver shim: iterations=" + std::to_string(iterations) +
             " initial_cost=" + std::to_string(initial_cost) +
             " final_cost=" + std::to_string(final_cost);
    }
    std::string BriefReport() const { return FullReport(); }
  };
};

namespace shim_internal {
// Stacks all residual blocks of a single 6-parameter block into the function
// object ceres::TinySolver expects (dynamic residual count, col-major jacobian).
struct StackedFunction {
  typedef double Scalar;
  enum { NUM_RESIDUALS = Eigen::Dynamic, NUM_PARAMETERS = 6 };
  const Problem* problem;
  int NumResiduals() const {
    int n = 0;
    for (const auto& b : problem->blocks_) n += b->num_residuals();
    return n;
  }
  bool operator()(const double* x, double* res, double* jac) const {
    const int total = NumResiduals();
    int row = 0;
    double jrow[16 * 6];
    for (const auto& b : problem->blocks_) {
      const int nr = b->num_residuals();
      if (!b->Evaluate(x, res + row, jac ? jrow : nullptr)) return false;
      if (jac) {
        for (int k = 0; k < nr; ++k)
          for (int i = 0; i < 6; ++i) jac[i * total + row + k] = jrow[k * 6 + i];
      }
      row += nr;
    }
    return true;
  }
};
}  // namespace shim_internal

inline void Solve(const Solver::Options& options, Problem* problem,
                  Solver::Summary* summary) {
  if (problem->params_ == nullptr || !problem->same_block_) {
    summary->status = -2;
    return;
  }
  for (const auto& b : problem->blocks_) {
    if (b->num_parameters() != 6 || b->num_residuals() > 16) {
      summary->status = -3;
      return;
    }
  }
  shim_internal::StackedFunction f{problem};
  TinySolver<shim_internal::StackedFunction> solver;
  solver.options.max_num_iterations = options.max_num_iterations;
  solver.options.gradient_tolerance = options.gradient_tolerance;
  solver.options.parameter_tolerance = options.parameter_tolerance;
  solver.options.initial_trust_region_radius = options.initial_trust_region_radius;
  Eigen::Matrix<double, 6, 1> x;
  std::memcpy(x.data(), problem->params_, 6 * sizeof(double));
  const auto& s = solver.Solve(f, &x);
  std::memcpy(problem->params_, x.data(), 6 * sizeof(double));
  summary->initial_cost = s.initial_cost;
  summary->final_cost = s.final_cost;
  summary->iterations = s.iterations;
  summary->status = static_cast<int>(s.status);
  if (options.minimizer_progress_to_stdout) {
    std::cout << summary->FullReport() << std::endl;
  }
}

}  // namespace ceres

#endif  // ESA_POSE_B200_ORACLE_CERES_SHIM_H_
