// oracle/shim/ceres/ceres.h — TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for <ceres/ceres.h> so that the reference's src/lidarFactor.hpp compiles UNMODIFIED into
// oracle/_ref (Ceres itself is absent from this image).  Provides a forward-mode dual number `ceres::Jet<T,N>`
// usable as an Eigen scalar, and just enough of CostFunction / AutoDiffCostFunction for the `Create` statics to
// compile (the oracle evaluates the functors' operator() directly with Jets and never calls Create).
#pragma once
#include <cmath>
#include <limits>
#include <Eigen/Core>

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

class CostFunction { public: virtual ~CostFunction() {} };
template <typename Functor, int kNumResiduals, int N0, int N1>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) {}
  ~AutoDiffCostFunction() { delete functor_; }
 private:
  Functor* functor_;
};
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
