// FUNCTIONAL stand-in for the OpenCV core types that cross the reference's inference interfaces, rich enough to RUN
// include/superslam_b200_adapter.hpp under the reference's own callers (oracle/dropin_harness.cpp compiles
// /root/reference/src/StereoFrontEnd.cc in place against it).  Same declarations as the declaration-only stub in
// tests/stubs/ (OpenCV 4.x core/mat.hpp, core/types.hpp), here with bodies: cv::Mat is a ref-counted 2-D array whose
// copies share the buffer, as in OpenCV.  TEST INFRASTRUCTURE - this image ships no OpenCV C++ headers.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
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
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  // header over caller-owned memory (no copy, no ownership), step in bytes; 0 = packed rows
  Mat(int r, int c, int type, void* d, size_t step_bytes = 0) {
    set_header(r, c, type);
    data = static_cast<uchar*>(d);
    if (step_bytes) step.p[0] = step_bytes;
  }
  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    if (m.data) std::memset(m.data, 0, m.step.p[0] * static_cast<size_t>(r));
    return m;
  }
  static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == this->type()) return;   // cv::Mat::create keeps a fitting buffer
    set_header(r, c, type);
    const size_t bytes = step.p[0] * static_cast<size_t>(r);
    buf_ = bytes ? std::shared_ptr<uchar>(new uchar[bytes], std::default_delete<uchar[]>()) : nullptr;
    data = buf_.get();
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  Mat clone() const {
    Mat m;
    copyTo(m);
    return m;
  }
  void copyTo(Mat& m) const {
    if (empty()) {
      m = Mat();
      return;
    }
    m.create(rows, cols, type());
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + y * m.step.p[0], data + y * step.p[0], cols * elemSize());
  }
  // same data, new channel count / row count (continuous matrices only, like cv::Mat::reshape without a copy)
  Mat reshape(int cn, int new_rows = 0) const {
    Mat m = *this;
    // OpenCV throws here (empty matrix: "Bad new number of rows"; non-continuous: "The matrix is not continuous, thus
    // its number of rows can not be changed"); code under test must not get that far, so the stand-in stops the process
    if ((empty() || !isContinuous()) && new_rows != 0 && new_rows != rows) {
      std::fprintf(stderr, "cv::Mat::reshape stand-in: OpenCV would throw (empty or non-continuous matrix)\n");
      std::abort();
    }
    if (empty()) return m;
    const size_t scalars = total() * static_cast<size_t>(channels());
    if (cn == 0) cn = channels();
    if (new_rows == 0) new_rows = rows;
    const int t = CV_MAKETYPE(depth(), cn);
    const int new_cols = static_cast<int>(scalars / (static_cast<size_t>(new_rows) * cn));
    uchar* d = m.data;
    m.set_header(new_rows, new_cols, t);
    m.data = d;
    return m;
  }
  // rows [start, end) of the same buffer
  Mat rowRange(int start, int end) const {
    Mat m = *this;
    m.rows = end - start;
    m.data = data + start * step.p[0];
    return m;
  }
  // appends the rows of `m` (same type and width); like OpenCV the buffer is reallocated, older headers keep the old one
  void push_back(const Mat& m) {
    if (m.empty()) return;
    if (empty()) {
      *this = m.clone();
      return;
    }
    Mat grown(rows + m.rows, cols, type());
    for (int y = 0; y < rows; ++y) std::memcpy(grown.ptr<uchar>(y), ptr<uchar>(y), cols * elemSize());
    for (int y = 0; y < m.rows; ++y) std::memcpy(grown.ptr<uchar>(rows + y), m.ptr<uchar>(y), cols * elemSize());
    *this = grown;
  }
  Mat t() const {   // CV_32F only (what src/PlaceRecognizer.cc transposes)
    Mat r(cols, rows, type());
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols; ++x) r.at<float>(x, y) = at<float>(y, x);
    return r;
  }
  Mat row(int y) const {
    Mat m = *this;
    m.rows = 1;
    m.data = data + y * step.p[0];
    return m;
  }
  void convertTo(Mat& m, int rtype, double alpha = 1, double beta = 0) const {
    const int ddepth = rtype < 0 ? depth() : (rtype & 7);
    Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
    const int n = cols * channels();
    // OpenCV scales in single precision when the destination is CV_32F or narrower (cvt_32f: src * (float)alpha +
    // (float)beta, fused), in double precision for CV_64F
    const float fa = static_cast<float>(alpha), fb = static_cast<float>(beta);
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < n; ++x) {
        if (ddepth == CV_64F) out.store(y, x, load(y, x) * alpha + beta);
        else out.store(y, x, std::fmaf(static_cast<float>(load(y, x)), fa, fb));
      }
    m = out;
  }
  bool isContinuous() const { return rows <= 1 || step.p[0] == cols * elemSize(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return flags & 0xFFF; }
  int depth() const { return flags & 7; }
  int channels() const { return ((flags & 0xFFF) >> 3) + 1; }
  size_t total() const { return static_cast<size_t>(rows) * cols; }
  size_t elemSize() const { return elem_size1(depth()) * channels(); }
  Size size() const { return Size(cols, rows); }
  template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + y * step.p[0]); }
  template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + y * step.p[0]); }
  template <typename T> T& at(int y, int x) { return ptr<T>(y)[x]; }
  template <typename T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }

  int flags = 0, dims = 0, rows = 0, cols = 0;
  uchar* data = nullptr;
  MatStep step;

 private:
  static size_t elem_size1(int depth) { return depth == CV_8U ? 1 : depth == CV_16U ? 2 : depth == CV_32F ? 4 : 8; }
  void set_header(int r, int c, int type) {
    flags = type & 0xFFF;
    dims = 2;
    rows = r;
    cols = c;
    step.p[1] = elemSize();
    step.p[0] = step.p[1] * static_cast<size_t>(c);
  }
  double load(int y, int x) const {   // x counts scalars (channels interleaved)
    const uchar* p = data + y * step.p[0];
    switch (depth()) {
      case CV_8U: return p[x];
      case CV_16U: return reinterpret_cast<const uint16_t*>(p)[x];
      case CV_32F: return reinterpret_cast<const float*>(p)[x];
      default: return reinterpret_cast<const double*>(p)[x];
    }
  }
  void store(int y, int x, double v) {   // float / double targets only need no saturation here
    uchar* p = data + y * step.p[0];
    switch (depth()) {
      case CV_8U: p[x] = static_cast<uchar>(v < 0 ? 0 : v > 255 ? 255 : v + 0.5); break;
      case CV_16U: reinterpret_cast<uint16_t*>(p)[x] = static_cast<uint16_t>(v < 0 ? 0 : v > 65535 ? 65535 : v + 0.5); break;
      case CV_32F: reinterpret_cast<float*>(p)[x] = static_cast<float>(v); break;
      default: reinterpret_cast<double*>(p)[x] = v;
    }
  }
  std::shared_ptr<uchar> buf_;
};

// ---- the three arithmetic calls of src/PlaceRecognizer.cc, CV_32F only.  The arithmetic is the stand-in's (double
// accumulation, one rounding to float), not OpenCV's SIMD kernels: results agree with OpenCV to the last few ulps,
// which is why tests that use them compare scores with a tolerance and only the control flow exactly.
enum { NORM_L2 = 4 };
void normalize(const Mat& src, Mat& dst, double alpha, double beta, int norm_type);   // declared only (see the shims)
inline double norm(const Mat& m) {
  double s = 0;
  for (int y = 0; y < m.rows; ++y)
    for (int x = 0; x < m.cols; ++x) s += static_cast<double>(m.at<float>(y, x)) * m.at<float>(y, x);
  return std::sqrt(s);
}
inline Mat operator/(const Mat& a, double s) {   // MatExpr a * (1 / s), evaluated by convertTo with a float scale
  Mat r(a.rows, a.cols, a.type());
  const float k = static_cast<float>(1.0 / s);
  for (int y = 0; y < a.rows; ++y)
    for (int x = 0; x < a.cols; ++x) r.at<float>(y, x) = a.at<float>(y, x) * k;
  return r;
}
inline Mat operator*(const Mat& a, const Mat& b) {   // gemm
  Mat r(a.rows, b.cols, a.type());
  for (int y = 0; y < a.rows; ++y)
    for (int x = 0; x < b.cols; ++x) {
      double s = 0;
      for (int k = 0; k < a.cols; ++k) s += static_cast<double>(a.at<float>(y, k)) * b.at<float>(k, x);
      r.at<float>(y, x) = static_cast<float>(s);
    }
  return r;
}
}  // namespace cv
