// <opencv2/imgproc.hpp> of the functional stand-in: DECLARATIONS of the two imgproc functions the reference's TensorRT
// wrapper classes call (src/SuperPoint.cc, src/EigenPlaces.cc).  oracle/ref_nethost_shim.cpp defines them: the channel
// reorderings exactly, cv::resize forwarded to the real OpenCV (cv2) through a hook.  TEST INFRASTRUCTURE.
#pragma once
#include "core.hpp"

namespace cv {
enum { COLOR_BGR2RGB = 4, COLOR_BGR2GRAY = 6, COLOR_GRAY2RGB = 8 };   // OpenCV 4.x ColorConversionCodes
void cvtColor(const Mat& src, Mat& dst, int code);
void resize(const Mat& src, Mat& dst, Size dsize);
}  // namespace cv
