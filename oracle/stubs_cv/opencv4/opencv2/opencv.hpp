// <opencv2/opencv.hpp> of the functional stand-in: core.hpp plus DECLARATIONS of the three imgproc / core functions
// /root/reference/src/SuperPoint.cc calls on paths that are not exercised here (image preprocessing in front of the
// TensorRT engine, host-descriptor normalisation); oracle/ref_nethost_shim.cpp defines them to abort.
// TEST INFRASTRUCTURE.
#pragma once
#include <algorithm>   // the real opencv2/core pulls these in; src/SuperPoint.cc relies on that for std::sort
#include <cmath>
#include <functional>

#include "core.hpp"

namespace cv {
enum { COLOR_BGR2GRAY = 6 };
enum { NORM_L2 = 4 };
void cvtColor(const Mat& src, Mat& dst, int code);
void resize(const Mat& src, Mat& dst, Size dsize);
void normalize(const Mat& src, Mat& dst, double alpha, double beta, int norm_type);
}  // namespace cv
