// Minimal FUNCTIONAL stand-in for the OpenCV core types that cross the reference's inference interfaces and front
// ends (cv::Mat as a typed 2-D view, cv::KeyPoint, cv::DMatch, cv::Point2f, cv::countNonZero, cv::noArray), so that
// /root/reference/src/StereoFrontEnd.cc and src/RgbdFrontEnd.cc can be compiled and RUN in place by oracle/Makefile.
// Field types follow OpenCV 4.x core/types.hpp (pt is two floats, queryIdx / trainIdx are ints; CV_16U = 2,
// CV_32F = 5, CV_64F = 6).  TEST INFRASTRUCTURE.
#pragma once
#include <cstddef>
#include <cstdint>

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_64F 6

namespace cv {
struct Point2f {
  float x = 0, y = 0;
  Point2f() {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct KeyPoint {
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
  KeyPoint() {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 0;
  DMatch() {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), distance(d) {}
};
// A non-owning single-channel 2-D view: enough for images that pass through, a depth map that is sampled with
// at<T>(row, col), and the small calibration matrices.
struct Mat {
  int rows = 0, cols = 0;
  int type_ = CV_8U;
  unsigned char* data = nullptr;
  size_t step = 0;   // bytes per row
  Mat() {}
  Mat(int r, int c, int type, void* d, size_t step_bytes) : rows(r), cols(c), type_(type), data(static_cast<unsigned char*>(d)), step(step_bytes) {}
  bool empty() const { return rows == 0 || cols == 0 || data == nullptr; }
  int type() const { return type_; }
  template <typename T>
  const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + static_cast<size_t>(r) * step + sizeof(T) * c); }
};
inline int countNonZero(const Mat& m) {
  int n = 0;
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) {
      if (m.type() == CV_64F) n += m.at<double>(r, c) != 0.0;
      else if (m.type() == CV_32F) n += m.at<float>(r, c) != 0.0f;
    }
  return n;
}
struct NoArrayTag {};
inline NoArrayTag noArray() { return NoArrayTag(); }
}  // namespace cv
