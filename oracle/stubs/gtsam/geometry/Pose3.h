// Stand-in for gtsam::Pose3: rotation matrix + translation, transformFrom(p) = R p + t (camera -> world for Twc).
// TEST INFRASTRUCTURE.
#pragma once
#include "Point3.h"
namespace gtsam {
class Pose3 {
 public:
  Pose3() {}
  Pose3(const double R[9], const double t[3]) {
    for (int i = 0; i < 9; ++i) R_[i] = R[i];
    for (int i = 0; i < 3; ++i) t_[i] = t[i];
  }
  Point3 transformFrom(const Point3& p) const {
    return Point3(R_[0] * p.x() + R_[1] * p.y() + R_[2] * p.z() + t_[0], R_[3] * p.x() + R_[4] * p.y() + R_[5] * p.z() + t_[1],
                  R_[6] * p.x() + R_[7] * p.y() + R_[8] * p.z() + t_[2]);
  }

 private:
  double R_[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double t_[3] = {0, 0, 0};
};
}  // namespace gtsam
