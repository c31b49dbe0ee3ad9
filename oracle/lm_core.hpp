// oracle/lm_core.hpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).
// The trust-region Levenberg-Marquardt loop of `ceres::Solve` as the reference configures it (Ceres 1.12.0, DENSE_QR,
// max_num_iterations 4, everything else default; laserOdometry.cpp:571-576, laserMapping.cpp:713-720), restated from the published
// algorithm ([3P-mem], parity unpinned: Ceres is not in /root/reference and not in this image).  It is written against an abstract
// evaluator so that TWO callers share it:
//   * oracle/ceres_lm.hpp      — the hand restatement (factor records, analytic or autodiff residuals);
//   * oracle/shim/ceres/ceres.h — the stand-in `ceres::Problem / ceres::Solve` behind which the reference's OWN node sources run
//     unmodified in oracle/_ref (cost functions = the reference's AutoDiffCostFunction<LidarEdgeFactor,...> objects).
// Problem shape (both): parameter blocks q (4, EigenQuaternionParameterization) and t (3): 7 global / 6 local parameters.
#pragma once
#include "common.hpp"
#include "linalg.hpp"
#include <cfloat>

namespace lvo_oracle {

struct LmTraceRow { double x[7]; double cost; double radius; int flags; };
// flags: bit0 step valid, bit1 step accepted, bit2 terminated by parameter tolerance, bit3 function tolerance,
//        bit4 gradient tolerance, bit5 this row is the initial evaluation

struct LmOptions {
  int max_num_iterations = 4;
  double huber = 0.1;
  double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  int max_num_consecutive_invalid_steps = 5;
};

struct LmSummary { int iterations = 0; double initial_cost = 0, final_cost = 0; int num_successful = 0; };

// observer of every solve_core call made through the stand-in ceres::Solve (oracle/refnode publishes the traces to its harness)
typedef void (*LmTraceHook)(int n_residual_blocks, const std::vector<LmTraceRow>&);
inline LmTraceHook& lm_trace_hook() { static LmTraceHook h = nullptr; return h; }

// EigenQuaternionParameterization::Plus (ceres/local_parameterization.cc): x_plus = q_delta * x,
// q_delta = (cos|d|, sin|d|/|d| * d); translation block is plain addition.
inline void plus7(const double* x, const double* delta, double* xp) {
  const double nd = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd > 0.0) {
    const double s = sin(nd) / nd;
    Quat dq{s * delta[0], s * delta[1], s * delta[2], cos(nd)};
    Quat q{x[0], x[1], x[2], x[3]};
    Quat r = qmul(dq, q);
    xp[0] = r.x; xp[1] = r.y; xp[2] = r.z; xp[3] = r.w;
  } else {
    xp[0] = x[0]; xp[1] = x[1]; xp[2] = x[2]; xp[3] = x[3];
  }
  xp[4] = x[4] + delta[3]; xp[5] = x[5] + delta[4]; xp[6] = x[6] + delta[5];
}


// Program evaluation as ceres::internal::ProgramEvaluator + ResidualBlock::Evaluate:
//   cost = 1/2 sum rho(|r|^2); r and J rows scaled by sqrt(rho') (Corrector, rho'' <= 0 branch).
struct Evaluation {
  double cost = 0;
  std::vector<double> r;   // corrected residuals
  std::vector<double> J;   // corrected Jacobian, rows x 6 row-major
  double g[6];             // J^T r
};

// ceres::Solve restated.  x (7) is updated in place; trace (optional) gets one row per iteration incl. row 0.
//   evaluate_fn(x, want_jacobian, Evaluation&) : cost, corrected residuals / local Jacobian rows (6 columns) and gradient at x
//   plus_fn(x, delta6, x_plus)                 : the parameter blocks' Plus (EigenQuaternionParameterization + identity)
//   empty                                      : no residual blocks
template <class EvalFn, class PlusFn>
inline LmSummary solve_core(EvalFn&& evaluate_fn, PlusFn&& plus_fn, bool empty, double* x, const LmOptions& opt, std::vector<LmTraceRow>* trace) {
  LmSummary sum;
  auto push_trace = [&](double cost, double radius, int flags) {
    if (!trace) return;
    LmTraceRow row;
    for (int i = 0; i < 7; ++i) row.x[i] = x[i];
    row.cost = cost; row.radius = radius; row.flags = flags;
    trace->push_back(row);
  };
  if (empty) {  // Ceres: "Terminating: Function tolerance reached. No non-constant parameter blocks found" / nothing to do
    push_trace(0, opt.initial_trust_region_radius, 32);
    return sum;
  }
  Evaluation e;
  evaluate_fn(x, true, e);
  double cost = e.cost;
  sum.initial_cost = sum.final_cost = cost;
  size_t rows = e.r.size();
  double x_norm = 0;
  for (int i = 0; i < 7; ++i) x_norm += x[i] * x[i];
  x_norm = sqrt(x_norm);

  auto gradient_max_norm = [&](const Evaluation& ev) {
    double ng[6], xp[7];
    for (int c = 0; c < 6; ++c) ng[c] = -ev.g[c];
    plus_fn(x, ng, xp);
    double m = 0;
    for (int i = 0; i < 7; ++i) m = std::max(m, fabs(x[i] - xp[i]));
    return m;
  };
  // jacobi scaling, frozen at iteration 0: scale = 1 / (1 + ||J_col||)
  double scale[6];
  for (int c = 0; c < 6; ++c) {
    double s = 0;
    for (size_t i = 0; i < rows; ++i) s += e.J[i * 6 + c] * e.J[i * 6 + c];
    scale[c] = 1.0 / (1.0 + sqrt(s));
  }
  auto scale_columns = [&](Evaluation& ev) {
    for (size_t i = 0; i < rows; ++i)
      for (int c = 0; c < 6; ++c) ev.J[i * 6 + c] *= scale[c];
  };
  double gmax = gradient_max_norm(e);
  scale_columns(e);
  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  double diagonal[6];
  push_trace(cost, radius, 32 | (gmax <= opt.gradient_tolerance ? 16 : 0));
  if (gmax <= opt.gradient_tolerance) return sum;

  int iteration = 0, invalid_run = 0;
  std::vector<double> A, rhs, model;
  while (true) {
    if (iteration >= opt.max_num_iterations) break;  // NO_CONVERGENCE
    ++iteration;
    sum.iterations = iteration;
    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int c = 0; c < 6; ++c) {
        double s = 0;
        for (size_t i = 0; i < rows; ++i) s += e.J[i * 6 + c] * e.J[i * 6 + c];
        diagonal[c] = std::min(std::max(s, opt.min_lm_diagonal), opt.max_lm_diagonal);
      }
    }
    double D[6];
    for (int c = 0; c < 6; ++c) D[c] = sqrt(diagonal[c] / radius);
    // DenseQRSolver: min || [J; diag(D)] y - [r; 0] ||, then step = -y
    A.assign((rows + 6) * 6, 0.0);
    rhs.assign(rows + 6, 0.0);
    std::copy(e.J.begin(), e.J.end(), A.begin());
    for (int c = 0; c < 6; ++c) A[(rows + c) * 6 + c] = D[c];
    std::copy(e.r.begin(), e.r.end(), rhs.begin());
    double step[6];
    bool ok = least_squares_qr(A.data(), (int)rows + 6, 6, rhs.data(), step);
    reuse_diagonal = true;
    bool valid = ok;
    double model_cost_change = 0;
    if (ok) {
      for (int c = 0; c < 6; ++c) { step[c] = -step[c]; if (!std::isfinite(step[c])) valid = false; }
    }
    if (valid) {
      // model_cost_change = -(J step)^T (r + J step / 2)
      for (size_t i = 0; i < rows; ++i) {
        double m = 0;
        for (int c = 0; c < 6; ++c) m += e.J[i * 6 + c] * step[c];
        model_cost_change -= m * (e.r[i] + m / 2.0);
      }
      if (model_cost_change <= 0.0) valid = false;
    }
    if (!valid) {
      if (++invalid_run >= opt.max_num_consecutive_invalid_steps) { push_trace(cost, radius, 0); break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;  // StepIsInvalid
      push_trace(cost, radius, 0);
      if (radius < opt.min_trust_region_radius) break;
      continue;
    }
    invalid_run = 0;
    double delta[6], xp[7];
    for (int c = 0; c < 6; ++c) delta[c] = step[c] * scale[c];
    plus_fn(x, delta, xp);
    Evaluation ec;
    evaluate_fn(xp, false, ec);
    double new_cost = ec.cost;
    double step_norm = 0;
    for (int i = 0; i < 7; ++i) step_norm += (x[i] - xp[i]) * (x[i] - xp[i]);
    step_norm = sqrt(step_norm);
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
      push_trace(cost, radius, 1 | 4);  // candidate discarded
      break;
    }
    double cost_change = cost - new_cost;
    if (fabs(cost_change) <= opt.function_tolerance * cost) {
      push_trace(cost, radius, 1 | 8);  // candidate discarded
      break;
    }
    double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > opt.min_relative_decrease) {
      ++sum.num_successful;
      radius = radius / std::max(1.0 / 3.0, 1.0 - pow(2.0 * relative_decrease - 1.0, 3));  // StepAccepted
      radius = std::min(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      for (int i = 0; i < 7; ++i) x[i] = xp[i];
      x_norm = 0;
      for (int i = 0; i < 7; ++i) x_norm += x[i] * x[i];
      x_norm = sqrt(x_norm);
      evaluate_fn(x, true, e);
      cost = e.cost;
      gmax = gradient_max_norm(e);
      scale_columns(e);
      sum.final_cost = cost;
      bool gconv = gmax <= opt.gradient_tolerance;
      push_trace(cost, radius, 1 | 2 | (gconv ? 16 : 0));
      if (gconv) break;
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;  // StepRejected
      push_trace(cost, radius, 1);
    }
    if (radius < opt.min_trust_region_radius) break;
  }
  return sum;
}

}  // namespace lvo_oracle
