// Declaration-only stand-in for the handful of OpenCV core types the reference's interface headers and
// include/superslam_b200_adapter.hpp touch.  TEST INFRASTRUCTURE: it exists so that the adapter can be
// type-checked in this image, which ships no OpenCV C++ headers (tests/test_adapter_compiles.py runs
// `g++ -fsyntax-only`).  Signatures follow OpenCV 4.x core/mat.hpp and core/types.hpp; nothing is implemented.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {

typedef unsigned char uchar;

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;

template <typename T>
struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size;

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

struct MatStep {
  size_t p[2] = {0, 0};
  size_t operator[](int i) const { return p[i]; }
  operator size_t() const { return p[0]; }
};

class Mat {
 public:
  Mat();
  Mat(int rows, int cols, int type);
  Mat(int rows, int cols, int type, void* data, size_t step = 0);
  Mat(const Mat&);
  Mat& operator=(const Mat&);
  ~Mat();
  static Mat zeros(int rows, int cols, int type);
  static Mat zeros(Size size, int type);
  void create(int rows, int cols, int type);
  void create(Size size, int type);
  Mat clone() const;
  Mat reshape(int cn, int rows = 0) const;
  Mat row(int y) const;
  void convertTo(Mat& m, int rtype, double alpha = 1, double beta = 0) const;
  void copyTo(Mat& m) const;
  bool isContinuous() const;
  bool empty() const;
  int type() const;
  int depth() const;
  int channels() const;
  size_t total() const;
  size_t elemSize() const;
  Size size() const;
  template <typename T> T* ptr(int y = 0);
  template <typename T> const T* ptr(int y = 0) const;
  template <typename T> T& at(int y, int x);
  template <typename T> const T& at(int y, int x) const;
  int flags = 0, dims = 0, rows = 0, cols = 0;
  uchar* data = nullptr;
  MatStep step;
};

}  // namespace cv
