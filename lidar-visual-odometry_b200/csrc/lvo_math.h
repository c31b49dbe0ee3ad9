// lvo_math.h — deterministic transcendental functions shared by host and device code.
//
// Why this exists: the ring id of a point (reference src/scanRegistration.cpp:166-192) and the
// fractional scan time stored in `intensity` (:141-153, :208-239) depend on atan / atan2.  glibc's and
// CUDA's libm differ in the last ulp, which flips ring ids at bin edges and changes intensity bits, so
// "feature labels bit-exact" cannot be met with two different libm's.  These versions use only IEEE
// +,-,*,/ in double (no FMA: device TUs are compiled with -fmad=false, host TUs with -ffp-contract=off)
// and are rounded to float once, so host and device results are bit-identical by construction.
// Accuracy: |error| < 1e-15 relative in double before the final rounding, i.e. the float result is the
// correctly rounded atan/atan2 except in ~1e-8 of cases (double-rounding), at most 1 float ulp away
// from glibc's atanf/atan2f (checked in tests/test_math.py).
#pragma once

#if defined(__CUDACC__)
#define LVO_HD __host__ __device__ __forceinline__
#else
#define LVO_HD inline
#endif

#define LVO_PI 3.14159265358979323846  /* == M_PI */

// atan(k/8), k = 0..8, correctly rounded doubles.
LVO_HD double lvo_atan_tab(int k) {
  switch (k) {
    case 0: return 0.0;
    case 1: return 0.12435499454676143503;
    case 2: return 0.24497866312686415417;
    case 3: return 0.35877067027057222040;
    case 4: return 0.46364760900080611621;
    case 5: return 0.55859931534356243597;
    case 6: return 0.64350110879328438680;
    case 7: return 0.71882999962162450;
    default: return 0.78539816339744830962;
  }
}

// atan for x in [0, 1].
LVO_HD double lvo_atan_unit(double x) {
  int k = (int)(x * 8.0 + 0.5);           // nearest table point c = k/8
  double c = (double)k * 0.125;
  double t = (x - c) / (1.0 + x * c);     // |t| <= 1/16
  double t2 = t * t;
  // odd Taylor series to t^17; remainder < (1/16)^19/19 ~ 7e-25
  double p = 1.0 / 17.0;
  p = p * t2 - 1.0 / 15.0;
  p = p * t2 + 1.0 / 13.0;
  p = p * t2 - 1.0 / 11.0;
  p = p * t2 + 1.0 / 9.0;
  p = p * t2 - 1.0 / 7.0;
  p = p * t2 + 1.0 / 5.0;
  p = p * t2 - 1.0 / 3.0;
  p = p * t2 + 1.0;
  return lvo_atan_tab(k) + t * p;
}

LVO_HD double lvo_atan_d(double x) {
  if (x != x) return x;
  double ax = x < 0.0 ? -x : x;
  double r;
  if (ax <= 1.0) r = lvo_atan_unit(ax);
  else r = 0.5 * LVO_PI - lvo_atan_unit(1.0 / ax);   // 1/inf = 0 -> pi/2
  return x < 0.0 ? -r : r;
}

// Full-quadrant atan2 in double (finite inputs; (0,0) -> 0 like glibc for +0,+0).
LVO_HD double lvo_atan2_d(double y, double x) {
  if (x != x || y != y) return x + y;
  if (y == 0.0) return (x < 0.0) ? LVO_PI : 0.0;
  double ay = y < 0.0 ? -y : y, axx = x < 0.0 ? -x : x;
  double r;
  if (axx >= ay) r = lvo_atan_unit(ay / axx);            // angle from the x axis, in [0, pi/4]
  else           r = 0.5 * LVO_PI - lvo_atan_unit(axx / ay);
  if (x < 0.0) r = LVO_PI - r;
  return y < 0.0 ? -r : r;
}

// The float "overloads" the reference resolves to (std::atan(float), std::atan2(float,float)).
LVO_HD float lvo_atanf(float x) { return (float)lvo_atan_d((double)x); }
LVO_HD float lvo_atan2f(float y, float x) { return (float)lvo_atan2_d((double)y, (double)x); }
