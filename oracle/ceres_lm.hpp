// oracle/ceres_lm.hpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).
// Restates what `ceres::Solve` does for the reference's problems (Ceres 1.12.0, docker/Dockerfile:3; NOT in
// /root/reference — restated from the published algorithm, parity unpinned):
//   * residual blocks = the functors of src/lidarFactor.hpp:12-138 under ceres::HuberLoss(0.1)
//     (laserOdometry.cpp:369, laserMapping.cpp:565) with Ceres' Corrector (rho'' <= 0 branch);
//   * parameter blocks q (4, EigenQuaternionParameterization) and t (3)  (laserOdometry.cpp:370-376);
//   * TrustRegionMinimizer + LevenbergMarquardtStrategy + DENSE_QR, max_num_iterations = 4, all other options at
//     their 1.12 defaults (laserOdometry.cpp:571-576, laserMapping.cpp:713-720).  SURVEY §8a rows A17-A19.
//
// Two builds of this file exist:
//   default                      : residuals/Jacobians analytic (the identity verified in SURVEY §8a A17),
//                                  3x3 eigen / least squares by the small routines in linalg.hpp;
//   -DLVO_ORACLE_USE_REFERENCE   : residuals/Jacobians by forward-mode autodiff of the reference's OWN functors
//                                  (#include "lidarFactor.hpp" from /root/reference/src through oracle/shim), and
//                                  Eigen 3.3.7's SelfAdjointEigenSolver / colPivHouseholderQr / householderQr.
#pragma once
#include "common.hpp"
#include "linalg.hpp"
#include "lm_core.hpp"
#include <cfloat>

#ifdef LVO_ORACLE_USE_REFERENCE
#include "lidarFactor.hpp"  // the reference's file, found through -I/root/reference/src
#endif

namespace lvo_oracle {

enum FactorType { F_EDGE = 0, F_PLANE = 1, F_PLANE_NORM = 2 };

// Parameters of one residual block.
//  F_EDGE       LidarEdgeFactor(curr_point=c, last_point_a=a, last_point_b=b, s=1)        lidarFactor.hpp:12-55
//  F_PLANE      LidarPlaneFactor(curr_point=c, j=a, l=b, m=m, s=1)                         lidarFactor.hpp:57-104
//  F_PLANE_NORM LidarPlaneNormFactor(curr_point=c, plane_unit_norm=a, negative_OA_dot_norm=d)  lidarFactor.hpp:106-138
struct Factor {
  int type;
  Vec3 c, a, b, m;
  double d;
  double s = 1.0;  // interpolation ratio of LidarEdgeFactor / LidarPlaneFactor (lidarFactor.hpp:15,60); 1.0 unless DISTORTION (laserOdometry.cpp:455-459,549-554)
};

// Eigen 3.3.7 QuaternionBase::slerp (Eigen/src/Geometry/Quaternion.h:716-746) of Identity towards q at parameter s, as the
// functors (lidarFactor.hpp:27-30,79-82) and TransformToStart (laserOdometry.cpp:163) call it.  Also returns the derivatives of
// the two blend weights with respect to q.w (the only component they depend on: d = Identity.dot(q) = q.w), which is what
// forward-mode autodiff of the functor propagates (abs: sign of the value; acos: -1 / sqrt(1 - a^2)).
struct SlerpWeights { double scale0, scale1, ds0_dw, ds1_dw; };
inline SlerpWeights slerp_identity_weights(double w, double s) {
  SlerpWeights o{0, 0, 0, 0};
  const double one = 1.0 - DBL_EPSILON;
  const double d = w, absD = fabs(d);
  if (absD >= one) { o.scale0 = 1.0 - s; o.scale1 = s; }
  else {
    const double theta = acos(absD), sinTheta = sin(theta), cosTheta = cos(theta);
    const double a0 = (1.0 - s) * theta, a1 = s * theta;
    o.scale0 = sin(a0) / sinTheta;
    o.scale1 = sin(a1) / sinTheta;
    const double dtheta_dw = (d < 0 ? 1.0 : -1.0) / sqrt(1.0 - absD * absD);
    // d/dtheta [sin(k theta) / sin(theta)] = (k cos(k theta) sin(theta) - sin(k theta) cos(theta)) / sin^2(theta)
    o.ds0_dw = ((1.0 - s) * cos(a0) * sinTheta - sin(a0) * cosTheta) / (sinTheta * sinTheta) * dtheta_dw;
    o.ds1_dw = (s * cos(a1) * sinTheta - sin(a1) * cosTheta) / (sinTheta * sinTheta) * dtheta_dw;
  }
  if (d < 0) { o.scale1 = -o.scale1; o.ds1_dw = -o.ds1_dw; }
  return o;
}
inline Quat slerp_identity(const Quat& q, double s) {
  const SlerpWeights k = slerp_identity_weights(q.w, s);
  return Quat{k.scale0 * 0.0 + k.scale1 * q.x, k.scale0 * 0.0 + k.scale1 * q.y, k.scale0 * 0.0 + k.scale1 * q.z, k.scale0 * 1.0 + k.scale1 * q.w};
}

inline int eval_factor_interp(const Factor& f, const double* x, double* r, double* J);

// Residual (k = 3 or 1 rows) and local 6-column Jacobian of one block at x = (q xyzw, t), before the loss.
inline int eval_factor(const Factor& f, const double* x, double* r, double* J /* k x 6 row-major, may be null */) {
#ifdef LVO_ORACLE_USE_REFERENCE
  typedef ceres::Jet<double, 7> JetT;
  JetT q[4], t[3], res[3];
  for (int i = 0; i < 4; ++i) q[i] = JetT(x[i], i);
  for (int i = 0; i < 3; ++i) t[i] = JetT(x[4 + i], 4 + i);
  int k;
  Eigen::Vector3d c(f.c.x, f.c.y, f.c.z), a(f.a.x, f.a.y, f.a.z), b(f.b.x, f.b.y, f.b.z), m(f.m.x, f.m.y, f.m.z);
  if (f.type == F_EDGE) { LidarEdgeFactor fn(c, a, b, f.s); fn(q, t, res); k = 3; }
  else if (f.type == F_PLANE) { LidarPlaneFactor fn(c, a, b, m, f.s); fn(q, t, res); k = 1; }
  else { LidarPlaneNormFactor fn(c, a, f.d); fn(q, t, res); k = 1; }
  // EigenQuaternionParameterization::ComputeJacobian (4x3, xyzw storage)
  const double P[4][3] = {{x[3], x[2], -x[1]}, {-x[2], x[3], x[0]}, {x[1], -x[0], x[3]}, {-x[0], -x[1], -x[2]}};
  for (int i = 0; i < k; ++i) {
    r[i] = res[i].a;
    if (J) {
      for (int cidx = 0; cidx < 3; ++cidx) {
        double s = 0;
        for (int g = 0; g < 4; ++g) s += res[i].v[g] * P[g][cidx];
        J[i * 6 + cidx] = s;
        J[i * 6 + 3 + cidx] = res[i].v[4 + cidx];
      }
    }
  }
  return k;
#else
  if (f.type != F_PLANE_NORM && f.s != 1.0) return eval_factor_interp(f, x, r, J);
  Quat q{x[0], x[1], x[2], x[3]};
  Vec3 rc = rotate(q, f.c);
  Vec3 lp{rc.x + x[4], rc.y + x[5], rc.z + x[6]};
  // d lp / d delta = -2 [R c]_x  (x (+) delta = q_delta * x, delta = half rotation vector); d lp / d t = I
  const double Jp[3][3] = {{0, 2 * rc.z, -2 * rc.y}, {-2 * rc.z, 0, 2 * rc.x}, {2 * rc.y, -2 * rc.x, 0}};
  if (f.type == F_EDGE) {
    Vec3 u{lp.x - f.a.x, lp.y - f.a.y, lp.z - f.a.z}, v{lp.x - f.b.x, lp.y - f.b.y, lp.z - f.b.z};
    Vec3 nu = cross(u, v);
    Vec3 de{f.a.x - f.b.x, f.a.y - f.b.y, f.a.z - f.b.z};
    double den = sqrt(de.x * de.x + de.y * de.y + de.z * de.z);
    r[0] = nu.x / den; r[1] = nu.y / den; r[2] = nu.z / den;
    if (J) {
      // d r / d lp = -[a-b]_x / |a-b|
      const double D[3][3] = {{0, de.z / den, -de.y / den}, {-de.z / den, 0, de.x / den}, {de.y / den, -de.x / den, 0}};
      for (int i = 0; i < 3; ++i)
        for (int cidx = 0; cidx < 3; ++cidx) {
          J[i * 6 + cidx] = D[i][0] * Jp[0][cidx] + D[i][1] * Jp[1][cidx] + D[i][2] * Jp[2][cidx];
          J[i * 6 + 3 + cidx] = D[i][cidx];
        }
    }
    return 3;
  }
  Vec3 n;
  double r0;
  if (f.type == F_PLANE) {
    Vec3 jl{f.a.x - f.b.x, f.a.y - f.b.y, f.a.z - f.b.z}, jm{f.a.x - f.m.x, f.a.y - f.m.y, f.a.z - f.m.z};
    n = cross(jl, jm);                                   // lidarFactor.hpp:64
    double nn = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);  // :65 normalize()
    n.x /= nn; n.y /= nn; n.z /= nn;
    r0 = (lp.x - f.a.x) * n.x + (lp.y - f.a.y) * n.y + (lp.z - f.a.z) * n.z;
  } else {
    n = f.a;
    r0 = (n.x * lp.x + n.y * lp.y + n.z * lp.z) + f.d;
  }
  r[0] = r0;
  if (J) {
    const double nv[3] = {n.x, n.y, n.z};
    for (int cidx = 0; cidx < 3; ++cidx) {
      J[cidx] = nv[0] * Jp[0][cidx] + nv[1] * Jp[1][cidx] + nv[2] * Jp[2][cidx];
      J[3 + cidx] = nv[cidx];
    }
  }
  return 1;
#endif
}

// LidarEdgeFactor / LidarPlaneFactor with s != 1 (DISTORTION 1): lp = slerp(Identity, q; s) * c + s t, lidarFactor.hpp:27-33,79-85.
// The global 3x4 Jacobian d lp / d(qx,qy,qz,qw) is taken through the functor exactly as written (slerp weights depend on q.w;
// Eigen's _transformVector formula v + w uv + u x uv, uv = 2 u x v, differentiated without assuming |q_s| = 1), then multiplied
// by the 4x3 plus-Jacobian of EigenQuaternionParameterization.
inline int eval_factor_interp(const Factor& f, const double* x, double* r, double* J) {
  const double s = f.s;
  const SlerpWeights k = slerp_identity_weights(x[3], s);
  const Vec3 u{k.scale1 * x[0], k.scale1 * x[1], k.scale1 * x[2]};
  const double ws = k.scale0 + k.scale1 * x[3];
  const Vec3 c = f.c;
  const Vec3 rc = rotate(Quat{u.x, u.y, u.z, ws}, c);
  const Vec3 lp{rc.x + s * x[4], rc.y + s * x[5], rc.z + s * x[6]};
  double Jl[3][6];
  if (J) {
    const double uc = u.x * c.x + u.y * c.y + u.z * c.z;
    const double cv[3] = {c.x, c.y, c.z}, uv_[3] = {u.x, u.y, u.z};
    const double cx[3][3] = {{0, -c.z, c.y}, {c.z, 0, -c.x}, {-c.y, c.x, 0}};
    double A[3][3];   // d lp / d u
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) A[i][j] = -2.0 * ws * cx[i][j] + 2.0 * ((i == j ? uc : 0.0) + uv_[i] * cv[j] - 2.0 * cv[i] * uv_[j]);
    const Vec3 b2 = cross(u, c);
    const double bw[3] = {2.0 * b2.x, 2.0 * b2.y, 2.0 * b2.z};  // d lp / d ws
    double G[3][4];   // d lp / d (qx, qy, qz, qw)
    const double dws_dw = k.ds0_dw + k.ds1_dw * x[3] + k.scale1;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) G[i][j] = A[i][j] * k.scale1;
      G[i][3] = (A[i][0] * x[0] + A[i][1] * x[1] + A[i][2] * x[2]) * k.ds1_dw + bw[i] * dws_dw;
    }
    const double P[4][3] = {{x[3], x[2], -x[1]}, {-x[2], x[3], x[0]}, {x[1], -x[0], x[3]}, {-x[0], -x[1], -x[2]}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Jl[i][j] = G[i][0] * P[0][j] + G[i][1] * P[1][j] + G[i][2] * P[2][j] + G[i][3] * P[3][j];
        Jl[i][3 + j] = i == j ? s : 0.0;
      }
  }
  if (f.type == F_EDGE) {
    Vec3 uu{lp.x - f.a.x, lp.y - f.a.y, lp.z - f.a.z}, vv{lp.x - f.b.x, lp.y - f.b.y, lp.z - f.b.z};
    Vec3 nu = cross(uu, vv);
    Vec3 de{f.a.x - f.b.x, f.a.y - f.b.y, f.a.z - f.b.z};
    double den = sqrt(de.x * de.x + de.y * de.y + de.z * de.z);
    r[0] = nu.x / den; r[1] = nu.y / den; r[2] = nu.z / den;
    if (J) {
      const double D[3][3] = {{0, de.z / den, -de.y / den}, {-de.z / den, 0, de.x / den}, {de.y / den, -de.x / den, 0}};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) J[i * 6 + j] = D[i][0] * Jl[0][j] + D[i][1] * Jl[1][j] + D[i][2] * Jl[2][j];
    }
    return 3;
  }
  Vec3 jl{f.a.x - f.b.x, f.a.y - f.b.y, f.a.z - f.b.z}, jm{f.a.x - f.m.x, f.a.y - f.m.y, f.a.z - f.m.z};
  Vec3 n = cross(jl, jm);
  double nn = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
  n.x /= nn; n.y /= nn; n.z /= nn;
  r[0] = (lp.x - f.a.x) * n.x + (lp.y - f.a.y) * n.y + (lp.z - f.a.z) * n.z;
  if (J) for (int j = 0; j < 6; ++j) J[j] = n.x * Jl[0][j] + n.y * Jl[1][j] + n.z * Jl[2][j];
  return 1;
}

inline void evaluate(const std::vector<Factor>& F, const double* x, double huber, bool jac, Evaluation& e) {
  e.cost = 0;
  e.r.clear();
  if (jac) e.J.clear();
  const double b = huber * huber;
  for (const Factor& f : F) {
    double r[3], J[18];
    int k = eval_factor(f, x, r, jac ? J : nullptr);
    double s = 0;
    for (int i = 0; i < k; ++i) s += r[i] * r[i];
    double rho0, rho1;
    if (s > b) { double rr = sqrt(s); rho0 = 2 * huber * rr - b; rho1 = std::max(DBL_MIN, huber / rr); }
    else { rho0 = s; rho1 = 1.0; }
    e.cost += 0.5 * rho0;
    double sc = sqrt(rho1);
    for (int i = 0; i < k; ++i) {
      e.r.push_back(r[i] * sc);
      if (jac) for (int c = 0; c < 6; ++c) e.J.push_back(J[i * 6 + c] * sc);
    }
  }
  if (jac) {
    for (int c = 0; c < 6; ++c) e.g[c] = 0;
    size_t rows = e.r.size();
    for (size_t i = 0; i < rows; ++i)
      for (int c = 0; c < 6; ++c) e.g[c] += e.J[i * 6 + c] * e.r[i];
  }
}

// ceres::Solve for a list of factor records (the LM itself is lm_core.hpp's solve_core)
inline LmSummary solve(const std::vector<Factor>& F, double* x, const LmOptions& opt, std::vector<LmTraceRow>* trace) {
  return solve_core([&](const double* xx, bool jac, Evaluation& e) { evaluate(F, xx, opt.huber, jac, e); }, plus7, F.empty(), x, opt, trace);
}

}  // namespace lvo_oracle
