// Minimal FUNCTIONAL stand-in for the OpenCV core types that cross the reference's inference interfaces
// (cv::Mat as an opaque pass-through, cv::KeyPoint, cv::DMatch, cv::Point2f), so that
// /root/reference/src/StereoFrontEnd.cc can be compiled and RUN in place by oracle/Makefile.  Field types follow
// OpenCV 4.x core/types.hpp (pt is two floats, queryIdx / trainIdx are ints).  TEST INFRASTRUCTURE.
#pragma once
#include <cstddef>

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
struct Mat {   // images and descriptor matrices only pass through the code under test
  int rows = 0, cols = 0;
  bool empty() const { return rows == 0 || cols == 0; }
};
}  // namespace cv
