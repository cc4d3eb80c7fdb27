// Stand-in for gtsam::StereoPoint2 (GTSAM 4.2 geometry/StereoPoint2.h): three doubles (uL, uR, v).  TEST INFRASTRUCTURE.
#pragma once
namespace gtsam {
class StereoPoint2 {
 public:
  StereoPoint2() : uL_(0), uR_(0), v_(0) {}
  StereoPoint2(double uL, double uR, double v) : uL_(uL), uR_(uR), v_(v) {}
  double uL() const { return uL_; }
  double uR() const { return uR_; }
  double v() const { return v_; }

 private:
  double uL_, uR_, v_;
};
}  // namespace gtsam
