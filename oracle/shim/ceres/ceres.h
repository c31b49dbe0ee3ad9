// oracle/shim/ceres/ceres.h — TEST INFRASTRUCTURE ONLY.
// Stand-in for <ceres/ceres.h> so that the reference's src/lidarFactor.hpp AND its node sources (src/laserOdometry.cpp:369-376,
// 462,557,571-576; src/laserMapping.cpp:565-572,619,684,713-720) compile UNMODIFIED into oracle/_ref (Ceres 1.12.0 itself is absent
// from this image).  Provides:
//   * `ceres::Jet<T,N>`, a forward-mode dual number usable as an Eigen scalar, with Ceres' jet.h formulas;
//   * CostFunction / AutoDiffCostFunction<Functor, kRes, 4, 3> (Evaluate = the functor on Jets, as Ceres' autodiff does);
//   * LossFunction / HuberLoss, LocalParameterization / EigenQuaternionParameterization (Ceres 1.12 loss_function.cc,
//     local_parameterization.cc formulas);
//   * Problem (two parameter blocks: q[4] with a parameterization, t[3]) and Solve(): residual-block evaluation with the
//     Corrector, then the trust-region LM loop of oracle/lm_core.hpp (restated from the published algorithm, parity unpinned).
#pragma once
#include <cmath>
#include <cfloat>
#include <limits>
#include <memory>
#include <set>
#include <string>
#include <vector>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include "lm_core.hpp"   // oracle/lm_core.hpp via -I<oracle>

namespace ceres {

template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a(T(0)) { for (int i = 0; i < N; ++i) v[i] = T(0); }
  Jet(const T& value) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(0); }  // NOLINT implicit like ceres::Jet
  Jet(const T& value, int k) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(0); v[k] = T(1); }
  Jet& operator+=(const Jet& o) { a += o.a; for (int i = 0; i < N; ++i) v[i] += o.v[i]; return *this; }
  Jet& operator-=(const Jet& o) { a -= o.a; for (int i = 0; i < N; ++i) v[i] -= o.v[i]; return *this; }
  Jet& operator*=(const Jet& o) { *this = *this * o; return *this; }
  Jet& operator/=(const Jet& o) { *this = *this / o; return *this; }
};
#define LVO_JET template <typename T, int N> inline
LVO_JET Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
LVO_JET Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> r; r.a = -f.a; for (int i = 0; i < N; ++i) r.v[i] = -f.v[i]; return r; }
LVO_JET Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> r; r.a = f.a + g.a; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] + g.v[i]; return r; }
LVO_JET Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> r; r.a = f.a - g.a; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] - g.v[i]; return r; }
LVO_JET Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> r; r.a = f.a * g.a; for (int i = 0; i < N; ++i) r.v[i] = f.a * g.v[i] + f.v[i] * g.a; return r; }
LVO_JET Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r; const T gi = T(1) / g.a; const T q = f.a * gi; r.a = q;
  for (int i = 0; i < N; ++i) r.v[i] = (f.v[i] - q * g.v[i]) * gi;
  return r;
}
LVO_JET Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> r = f; r.a += s; return r; }
LVO_JET Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> r = f; r.a += s; return r; }
LVO_JET Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> r = f; r.a -= s; return r; }
LVO_JET Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> r = -f; r.a += s; return r; }
LVO_JET Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> r; r.a = f.a * s; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * s; return r; }
LVO_JET Jet<T, N> operator*(T s, const Jet<T, N>& f) { return f * s; }
LVO_JET Jet<T, N> operator/(const Jet<T, N>& f, T s) { Jet<T, N> r; const T si = T(1) / s; r.a = f.a * si; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * si; return r; }
LVO_JET Jet<T, N> operator/(T s, const Jet<T, N>& g) { return Jet<T, N>(s) / g; }
#define LVO_JET_CMP(op) \
  LVO_JET bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
  LVO_JET bool operator op(const Jet<T, N>& f, T g) { return f.a op g; }                    \
  LVO_JET bool operator op(T f, const Jet<T, N>& g) { return f op g.a; }
LVO_JET_CMP(<) LVO_JET_CMP(<=) LVO_JET_CMP(>) LVO_JET_CMP(>=) LVO_JET_CMP(==) LVO_JET_CMP(!=)
LVO_JET Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0) ? -f : f; }
LVO_JET Jet<T, N> sqrt(const Jet<T, N>& f) { Jet<T, N> r; r.a = std::sqrt(f.a); const T d = T(1) / (T(2) * r.a); for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * d; return r; }
LVO_JET Jet<T, N> sin(const Jet<T, N>& f) { Jet<T, N> r; r.a = std::sin(f.a); const T c = std::cos(f.a); for (int i = 0; i < N; ++i) r.v[i] = c * f.v[i]; return r; }
LVO_JET Jet<T, N> cos(const Jet<T, N>& f) { Jet<T, N> r; r.a = std::cos(f.a); const T s = -std::sin(f.a); for (int i = 0; i < N; ++i) r.v[i] = s * f.v[i]; return r; }
LVO_JET Jet<T, N> acos(const Jet<T, N>& f) { Jet<T, N> r; r.a = std::acos(f.a); const T d = -T(1) / std::sqrt(T(1) - f.a * f.a); for (int i = 0; i < N; ++i) r.v[i] = d * f.v[i]; return r; }
LVO_JET bool isfinite(const Jet<T, N>& f) { return std::isfinite(f.a); }

class CostFunction {
 public:
  virtual ~CostFunction() {}
  // parameters[b] = block b; residuals[num_residuals]; jacobians[b] (may be null) = num_residuals x block_size, row-major
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  int num_residuals() const { return num_residuals_; }
  const std::vector<int>& parameter_block_sizes() const { return parameter_block_sizes_; }
 protected:
  int num_residuals_ = 0;
  std::vector<int> parameter_block_sizes_;
};
template <typename Functor, int kNumResiduals, int N0, int N1>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) { num_residuals_ = kNumResiduals; parameter_block_sizes_ = {N0, N1}; }
  ~AutoDiffCostFunction() { delete functor_; }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (!jacobians) return (*functor_)(parameters[0], parameters[1], residuals);
    typedef Jet<double, N0 + N1> JetT;
    JetT p0[N0], p1[N1], r[kNumResiduals];
    for (int i = 0; i < N0; ++i) p0[i] = JetT(parameters[0][i], i);
    for (int i = 0; i < N1; ++i) p1[i] = JetT(parameters[1][i], N0 + i);
    if (!(*functor_)(p0, p1, r)) return false;
    for (int k = 0; k < kNumResiduals; ++k) {
      residuals[k] = r[k].a;
      if (jacobians[0]) for (int i = 0; i < N0; ++i) jacobians[0][k * N0 + i] = r[k].v[i];
      if (jacobians[1]) for (int i = 0; i < N1; ++i) jacobians[1][k * N1 + i] = r[k].v[N0 + i];
    }
    return true;
  }
 private:
  Functor* functor_;
};

// ceres/loss_function.h: rho[0] = rho(s), rho[1] = rho'(s), rho[2] = rho''(s), s = squared norm of the block's residual
class LossFunction { public: virtual ~LossFunction() {} virtual void Evaluate(double s, double rho[3]) const = 0; };
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a_ * r - b_;
      rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  }
 private:
  const double a_, b_;
};

class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;   // GlobalSize x LocalSize, row-major
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
// ceres 1.12 local_parameterization.cc: Eigen storage order (x, y, z, w); x_plus_delta = q_delta * x, delta = half rotation vector
class EigenQuaternionParameterization : public LocalParameterization {
 public:
  bool Plus(const double* x_ptr, const double* delta, double* x_plus_delta_ptr) const override {
    Eigen::Map<Eigen::Quaterniond> x_plus_delta(x_plus_delta_ptr);
    Eigen::Map<const Eigen::Quaterniond> x(x_ptr);
    const double norm_delta = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    if (norm_delta > 0.0) {
      const double sin_delta_by_delta = std::sin(norm_delta) / norm_delta;
      Eigen::Quaterniond delta_q(std::cos(norm_delta), sin_delta_by_delta * delta[0], sin_delta_by_delta * delta[1], sin_delta_by_delta * delta[2]);
      x_plus_delta = delta_q * x;
    } else {
      x_plus_delta = x;
    }
    return true;
  }
  bool ComputeJacobian(const double* x, double* jacobian) const override {
    jacobian[0] = x[3];  jacobian[1] = x[2];   jacobian[2] = -x[1];
    jacobian[3] = -x[2]; jacobian[4] = x[3];   jacobian[5] = x[0];
    jacobian[6] = x[1];  jacobian[7] = -x[0];  jacobian[8] = x[3];
    jacobian[9] = -x[0]; jacobian[10] = -x[1]; jacobian[11] = -x[2];
    return true;
  }
  int GlobalSize() const override { return 4; }
  int LocalSize() const override { return 3; }
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };

class Problem {
 public:
  struct Options {};
  Problem() {}
  explicit Problem(const Options&) {}
  // Problem owns cost / loss functions and parameterizations (Ceres' default ownership); shared objects are freed once
  ~Problem() {
    std::set<const CostFunction*> cs; std::set<const LossFunction*> ls;
    for (auto& b : blocks_) { cs.insert(b.cost); if (b.loss) ls.insert(b.loss); }
    for (auto* c : cs) delete c;
    for (auto* l : ls) delete l;
    std::set<const LocalParameterization*> ps;
    for (auto& p : params_) if (p.local) ps.insert(p.local);
    for (auto* p : ps) delete p;
  }
  void AddParameterBlock(double* values, int size) { add_param(values, size, nullptr); }
  void AddParameterBlock(double* values, int size, LocalParameterization* lp) { add_param(values, size, lp); }
  void AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1) {
    add_param(x0, cost->parameter_block_sizes()[0], nullptr);
    add_param(x1, cost->parameter_block_sizes()[1], nullptr);
    blocks_.push_back(Block{cost, loss, x0, x1});
  }
  int NumResidualBlocks() const { return (int)blocks_.size(); }
  struct Param { double* values; int size; LocalParameterization* local; };
  struct Block { CostFunction* cost; LossFunction* loss; double* x0; double* x1; };
  const std::vector<Param>& params() const { return params_; }
  const std::vector<Block>& blocks() const { return blocks_; }
 private:
  void add_param(double* v, int size, LocalParameterization* lp) {
    for (auto& p : params_) if (p.values == v) { if (lp) p.local = lp; return; }
    params_.push_back(Param{v, size, lp});
  }
  std::vector<Param> params_;
  std::vector<Block> blocks_;
};

class Solver {
 public:
  struct Options {
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    int max_num_iterations = 50;
    bool minimizer_progress_to_stdout = false;
    int num_threads = 1;
    bool check_gradients = false;                       // laserMapping.cpp:717-718 (off: no effect on the solve)
    double gradient_check_relative_precision = 1e-8;
  };
  struct Summary {
    double initial_cost = 0, final_cost = 0;
    int num_successful_steps = 0, num_unsuccessful_steps = 0, iterations = 0;
    std::vector<lvo_oracle::LmTraceRow> trace;   // oracle extension: one row per LM iteration (not part of Ceres)
    std::string BriefReport() const { return "lvo shim ceres::Solve"; }
    std::string FullReport() const { return BriefReport(); }
  };
};

// ceres::Solve for the reference's problems: parameter block 0 = q (4 values, with a LocalParameterization of local size 3),
// block 1 = t (3 values); every residual block depends on (q, t) in that order.  Evaluation follows
// ceres::internal::ResidualBlock::Evaluate: global Jacobians -> local (J_q * d Plus / d delta at delta = 0), loss, Corrector
// (rho'' <= 0 branch for Huber: residuals and Jacobian rows scaled by sqrt(rho')), cost = 1/2 sum rho.
inline void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  using namespace lvo_oracle;
  const auto& P = problem->params();
  if (P.size() != 2 || P[0].size != 4 || P[1].size != 3 || !P[0].local || P[0].local->LocalSize() != 3) {
    fprintf(stderr, "lvo shim ceres::Solve: unsupported problem shape\n"); abort();
  }
  if (options.linear_solver_type != DENSE_QR) { fprintf(stderr, "lvo shim ceres::Solve: only DENSE_QR is modelled\n"); abort(); }
  double* q = P[0].values; double* t = P[1].values;
  const LocalParameterization* lp = P[0].local;
  const auto& B = problem->blocks();
  for (const auto& b : B) if (b.x0 != q || b.x1 != t) { fprintf(stderr, "lvo shim ceres::Solve: residual block over foreign parameters\n"); abort(); }
  auto evaluate_fn = [&](const double* x, bool jac, Evaluation& e) {
    e.cost = 0; e.r.clear();
    if (jac) e.J.clear();
    double lj[12];
    if (jac) lp->ComputeJacobian(x, lj);
    for (const auto& b : B) {
      const int k = b.cost->num_residuals();
      double r[3], Jq[12], Jt[9];
      const double* params[2] = {x, x + 4};
      double* jacs[2] = {Jq, Jt};
      b.cost->Evaluate(params, r, jac ? jacs : nullptr);
      double s = 0;
      for (int i = 0; i < k; ++i) s += r[i] * r[i];
      double rho[3] = {s, 1.0, 0.0};
      if (b.loss) b.loss->Evaluate(s, rho);
      e.cost += 0.5 * rho[0];
      const double sc = std::sqrt(rho[1]);
      for (int i = 0; i < k; ++i) {
        e.r.push_back(r[i] * sc);
        if (jac) {
          for (int c = 0; c < 3; ++c) { double a = 0; for (int g = 0; g < 4; ++g) a += Jq[i * 4 + g] * lj[g * 3 + c]; e.J.push_back(a * sc); }
          for (int c = 0; c < 3; ++c) e.J.push_back(Jt[i * 3 + c] * sc);
        }
      }
    }
    if (jac) {
      for (int c = 0; c < 6; ++c) e.g[c] = 0;
      const size_t rows = e.r.size();
      for (size_t i = 0; i < rows; ++i) for (int c = 0; c < 6; ++c) e.g[c] += e.J[i * 6 + c] * e.r[i];
    }
  };
  auto plus_fn = [&](const double* x, const double* delta, double* xp) {
    lp->Plus(x, delta, xp);
    xp[4] = x[4] + delta[3]; xp[5] = x[5] + delta[4]; xp[6] = x[6] + delta[5];
  };
  LmOptions opt;
  opt.max_num_iterations = options.max_num_iterations;
  double x[7] = {q[0], q[1], q[2], q[3], t[0], t[1], t[2]};
  summary->trace.clear();
  const LmSummary s = solve_core(evaluate_fn, plus_fn, B.empty(), x, opt, &summary->trace);
  for (int i = 0; i < 4; ++i) q[i] = x[i];
  for (int i = 0; i < 3; ++i) t[i] = x[4 + i];
  if (lm_trace_hook()) lm_trace_hook()((int)B.size(), summary->trace);
  summary->initial_cost = s.initial_cost; summary->final_cost = s.final_cost; summary->iterations = s.iterations;
  summary->num_successful_steps = s.num_successful;
}
}  // namespace ceres

namespace Eigen {
template <typename T, int N>
struct NumTraits<ceres::Jet<T, N>> {
  typedef ceres::Jet<T, N> Real;
  typedef ceres::Jet<T, N> NonInteger;
  typedef ceres::Jet<T, N> Nested;
  typedef ceres::Jet<T, N> Literal;
  static ceres::Jet<T, N> dummy_precision() { return ceres::Jet<T, N>(1e-12); }
  static inline Real epsilon() { return Real(std::numeric_limits<T>::epsilon()); }
  static inline int digits10() { return NumTraits<T>::digits10(); }
  static inline Real highest() { return Real(std::numeric_limits<T>::max()); }
  static inline Real lowest() { return Real(-std::numeric_limits<T>::max()); }
  enum { IsComplex = 0, IsInteger = 0, IsSigned, ReadCost = 1, AddCost = 1, MulCost = 3, HasFloatingPoint = 1, RequireInitialization = 1 };
};
template <typename BinaryOp, typename T, int N>
struct ScalarBinaryOpTraits<ceres::Jet<T, N>, T, BinaryOp> { typedef ceres::Jet<T, N> ReturnType; };
template <typename BinaryOp, typename T, int N>
struct ScalarBinaryOpTraits<T, ceres::Jet<T, N>, BinaryOp> { typedef ceres::Jet<T, N> ReturnType; };
}  // namespace Eigen
