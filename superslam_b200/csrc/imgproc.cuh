// Image front door either side of the networks, on the device:
//   Rectifier   cv::remap(img, out, M1, M2, INTER_LINEAR) with CV_32FC1 maps and a constant 0 border - the
//               EuRoC stereo rectification of /root/reference/examples/stereo/euroc.cc:118-133,176-177 -
//               bit-exact with OpenCV's fixed-point arithmetic (1/32-pixel coordinates, 15-bit weights).
//   RgbdPost    RgbdFrontEnd::process after the extraction (/root/reference/src/RgbdFrontEnd.cc:27-58):
//               cv::undistortPoints (five Brown-Conrady fixed-point iterations in fp64, no FMA contraction),
//               depth sampled at the raw pixel, uR = uL - bf / Z.
// oracle/imgproc.py is the CPU restatement, pinned bit-for-bit against cv2.
#pragma once

#include <cstdint>

#include "common.cuh"

namespace ssb {

// Host half of the remap: OpenCV's conversion of CV_32F maps to 1/32-pixel fixed point (needs no device).
void convert_remap_maps(const float* map_x, const float* map_y, size_t n, uint32_t* xy, uint16_t* frac);

class Rectifier {
 public:
  ~Rectifier();
  // map_x / map_y: float32 [dst_h][dst_w] source coordinates (initUndistortRectifyMap, CV_32F);
  // source images are src_h x src_w.  The maps are converted to OpenCV's fixed-point form once, here.
  int init(const float* map_x, const float* map_y, int dst_h, int dst_w, int src_h, int src_w, int max_images,
           int device);
  // host images in (u8 gray, row_stride bytes per row), host images out (dst_w bytes per row)
  int remap(const uint8_t* const* images, int count, int row_stride, uint8_t* const* out);
  // device -> device: src [count][src_h][src_w], dst [count][dst_h][dst_w], enqueue only.  src_pitch / dst_pitch:
  // bytes between consecutive images (0 = contiguous), e.g. every second image of an interleaved L/R batch.
  int remap_device(const uint8_t* src_dev, int count, uint8_t* dst_dev, cudaStream_t stream, size_t src_pitch = 0,
                   size_t dst_pitch = 0);
  int src_h() const { return sh_; }
  int src_w() const { return sw_; }
  int device() const { return device_; }
  int dst_h() const { return dh_; }
  int dst_w() const { return dw_; }
  cudaStream_t stream() const { return stream_; }

 private:
  int device_ = 0, dh_ = 0, dw_ = 0, sh_ = 0, sw_ = 0, cap_ = 0;
  uint32_t* xy_ = nullptr;    // [dst_h][dst_w] (uint16)ix | (uint16)iy << 16  (int16 each, saturated)
  uint16_t* frac_ = nullptr;  // [dst_h][dst_w] fy * 32 + fx
  uint8_t *src_ = nullptr, *dst_ = nullptr;          // device staging for the host path
  uint8_t *src_host_ = nullptr, *dst_host_ = nullptr;  // pinned
  cudaStream_t stream_ = nullptr;
};

struct RgbdParams {
  double fx, fy, cx, cy;
  double k[14];        // distortion coefficients, OpenCV order (k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 tx ty)
  int has_dist;        // countNonZero(dist_coeffs) > 0
  double bf, depth_factor, max_depth;
};

class RgbdPost {
 public:
  ~RgbdPost();
  int init(int max_keypoints, int max_h, int max_w, int device);
  // xy: host float [n][2] raw keypoints; depth: host image, depth_type 0 = u16, 1 = f32, row stride in bytes.
  // out_xy float [n][2] undistorted; out_stereo double [n][3] (uL, uR or NaN, v); out_has_depth [n].
  int process(const float* xy, int n, const void* depth, int depth_type, int dh, int dw, int row_stride,
              const RgbdParams& p, float* out_xy, double* out_stereo, uint8_t* out_has_depth);

 private:
  int device_ = 0, kmax_ = 0, hmax_ = 0, wmax_ = 0;
  float *xy_ = nullptr, *oxy_ = nullptr;
  double* stereo_ = nullptr;
  uint8_t* has_ = nullptr;
  uint8_t* depth_ = nullptr;
  uint8_t* host_ = nullptr;   // pinned: [xy | depth] in, [oxy | stereo | has] out
  size_t host_bytes_ = 0;
  cudaStream_t stream_ = nullptr;
};

}  // namespace ssb
