// <opencv2/opencv.hpp> of the functional stand-in: core.hpp + imgproc.hpp (DECLARATIONS of the imgproc / core functions
// /root/reference/src/SuperPoint.cc calls on paths that are not exercised here (image preprocessing in front of the
// TensorRT engine, host-descriptor normalisation); oracle/ref_nethost_shim.cpp defines them to abort.
// TEST INFRASTRUCTURE.
#pragma once
#include <algorithm>   // the real opencv2/core pulls these in; src/SuperPoint.cc relies on that for std::sort
#include <cmath>
#include <functional>

#include "core.hpp"
#include "imgproc.hpp"
