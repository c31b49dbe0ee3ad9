// oracle/linalg.hpp — TEST INFRASTRUCTURE ONLY (see oracle/common.hpp).
// The three dense solves on the path:
//   least_squares_qr  <- Eigen householderQr().solve() inside Ceres DenseQRSolver (ceres 1.12 dense_qr_solver.cc) and
//                        Eigen colPivHouseholderQr().solve() at reference src/laserMapping.cpp:663
//   sym_eigen3        <- Eigen::SelfAdjointEigenSolver<Matrix3d> at src/laserMapping.cpp:605
// Default build: plain Householder / cyclic Jacobi.  -DLVO_ORACLE_USE_REFERENCE: the reference's vendored Eigen 3.3.7.
#pragma once
#include "common.hpp"

#ifdef LVO_ORACLE_USE_REFERENCE
#include <Eigen/Dense>
#endif

namespace lvo_oracle {

// Solve min ||A y - b||_2, A is m x n row-major (destroyed), b length m (destroyed), y length n.
inline bool least_squares_qr(double* A, int m, int n, double* b, double* y) {
#ifdef LVO_ORACLE_USE_REFERENCE
  Eigen::Map<Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>> Am(A, m, n);
  Eigen::Map<Eigen::VectorXd> bm(b, m);
  Eigen::MatrixXd Ac = Am;
  Eigen::VectorXd sol = Ac.householderQr().solve(bm);
  for (int j = 0; j < n; ++j) y[j] = sol[j];
  return true;
#else
  for (int k = 0; k < n; ++k) {
    double norm = 0;
    for (int i = k; i < m; ++i) norm += A[i * n + k] * A[i * n + k];
    norm = sqrt(norm);
    if (norm == 0.0) return false;
    double alpha = A[k * n + k] > 0 ? -norm : norm;
    // v = x - alpha e1 ; H = I - 2 v v^T / (v^T v)
    double v0 = A[k * n + k] - alpha;
    double vtv = v0 * v0;
    for (int i = k + 1; i < m; ++i) vtv += A[i * n + k] * A[i * n + k];
    if (vtv == 0.0) return false;
    for (int j = k + 1; j < n; ++j) {
      double dot = v0 * A[k * n + j];
      for (int i = k + 1; i < m; ++i) dot += A[i * n + k] * A[i * n + j];
      double f = 2.0 * dot / vtv;
      A[k * n + j] -= f * v0;
      for (int i = k + 1; i < m; ++i) A[i * n + j] -= f * A[i * n + k];
    }
    {
      double dot = v0 * b[k];
      for (int i = k + 1; i < m; ++i) dot += A[i * n + k] * b[i];
      double f = 2.0 * dot / vtv;
      b[k] -= f * v0;
      for (int i = k + 1; i < m; ++i) b[i] -= f * A[i * n + k];
    }
    A[k * n + k] = alpha;
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[k * n + j] * y[j];
    y[k] = s / A[k * n + k];
  }
  return true;
#endif
}

// Plane fit of laserMapping.cpp:650-665: solve the 5x3 system A n = -1 in the least-squares sense.
inline void plane_fit5(const double A5x3[15], double n[3]) {
#ifdef LVO_ORACLE_USE_REFERENCE
  Eigen::Matrix<double, 5, 3> matA0;
  Eigen::Matrix<double, 5, 1> matB0 = -1 * Eigen::Matrix<double, 5, 1>::Ones();
  for (int j = 0; j < 5; ++j) for (int c = 0; c < 3; ++c) matA0(j, c) = A5x3[j * 3 + c];
  Eigen::Vector3d norm = matA0.colPivHouseholderQr().solve(matB0);
  n[0] = norm[0]; n[1] = norm[1]; n[2] = norm[2];
#else
  double A[15], b[5] = {-1, -1, -1, -1, -1};
  for (int i = 0; i < 15; ++i) A[i] = A5x3[i];
  if (!least_squares_qr(A, 5, 3, b, n)) { n[0] = n[1] = n[2] = 0; }
#endif
}

// Eigen-decomposition of a symmetric 3x3 (row-major).  Eigenvalues ascending, eigenvectors as columns V[r*3+c].
inline void sym_eigen3(const double M[9], double w[3], double V[9]) {
#ifdef LVO_ORACLE_USE_REFERENCE
  Eigen::Matrix3d covMat;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) covMat(r, c) = M[r * 3 + c];
  Eigen::SelfAdjointEigenSolver<Eigen::Matrix3d> saes(covMat);
  for (int c = 0; c < 3; ++c) { w[c] = saes.eigenvalues()[c]; for (int r = 0; r < 3; ++r) V[r * 3 + c] = saes.eigenvectors()(r, c); }
#else
  double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) a[r][c] = M[r * 3 + c];
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  double ev[3] = {a[0][0], a[1][1], a[2][2]};
  std::sort(order, order + 3, [&](int i, int j) { return ev[i] < ev[j]; });
  for (int c = 0; c < 3; ++c) { w[c] = ev[order[c]]; for (int r = 0; r < 3; ++r) V[r * 3 + c] = v[r][order[c]]; }
#endif
}

}  // namespace lvo_oracle
