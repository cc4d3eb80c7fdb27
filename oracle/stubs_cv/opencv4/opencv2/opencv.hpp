// <opencv2/opencv.hpp> of the functional stand-in: core.hpp + imgproc.hpp + the std headers the real one drags in.
// cv::cvtColor / cv::resize / cv::normalize are only DECLARED (imgproc.hpp, core.hpp); each shim that links the reference's
// wrapper classes defines them - to serve the calls its test exercises, to abort on the others.  TEST INFRASTRUCTURE.
#pragma once
#include <algorithm>   // the real opencv2/core pulls these in; src/SuperPoint.cc relies on that for std::sort
#include <cmath>
#include <functional>

#include "core.hpp"
#include "imgproc.hpp"
