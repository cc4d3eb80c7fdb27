// Stand-in for the one calib3d call on the path: cv::undistortPoints(raw, undist, K, D, noArray(), K)
// (/root/reference/src/RgbdFrontEnd.cc:32).  The arithmetic is NOT restated here: the call is forwarded to a
// function pointer that the test installs, and the test hands it to the real OpenCV (cv2.undistortPoints).
// TEST INFRASTRUCTURE.
#pragma once
#include <vector>

#include "core.hpp"

namespace cv {
// in [n][2] float, K 3x3 row-major doubles, D nd doubles, P 3x3 row-major doubles, out [n][2] float
using UndistortPointsFn = void (*)(const float* in, int n, const double* K, const double* D, int nd, const double* P, float* out);
inline UndistortPointsFn& undistort_points_hook() {
  static UndistortPointsFn fn = nullptr;
  return fn;
}
inline void undistortPoints(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& K, const Mat& D,
                            NoArrayTag, const Mat& P) {
  dst.resize(src.size());
  double k[9], p[9], d[16] = {0};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) k[3 * r + c] = K.at<double>(r, c), p[3 * r + c] = P.at<double>(r, c);
  const int nd = D.rows * D.cols;
  for (int i = 0; i < nd && i < 16; ++i) d[i] = D.at<double>(i / D.cols, i % D.cols);
  undistort_points_hook()(reinterpret_cast<const float*>(src.data()), static_cast<int>(src.size()), k, d, nd, p,
                          reinterpret_cast<float*>(dst.data()));
}
}  // namespace cv
