// Stand-in for gtsam::Point3 (an Eigen 3-vector in GTSAM 4.2).  TEST INFRASTRUCTURE.
#pragma once
namespace gtsam {
struct Point3 {
  double v[3] = {0, 0, 0};
  Point3() {}
  Point3(double x, double y, double z) : v{x, y, z} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};
}  // namespace gtsam
