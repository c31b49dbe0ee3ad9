// shim (test infrastructure): laserOdometry.cpp:246-252 creates a vloam::TwoFramePhotometricFunction only to print its pyramid level
// (include/Optimization/FrameTracker.h:18-47); the direct visual tracker is OUT OF SCOPE.
#pragma once
#include <ceres/ceres.h>
namespace vloam {
class TwoFramePhotometricFunction : public ceres::CostFunction {
 public:
  TwoFramePhotometricFunction(void*, void*) {}
  bool Evaluate(double const* const*, double*, double**) const override { return false; }
  mutable int current_level_ = 2;
};
}
