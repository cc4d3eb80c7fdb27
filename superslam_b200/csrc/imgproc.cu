// Device-side image front door: rectification remap and the RGB-D keypoint post-process (imgproc.cuh).
#include "imgproc.cuh"

#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

namespace ssb {

// =================================================================================================
// remap
// =================================================================================================

// One thread per four consecutive destination pixels.  Per pixel (OpenCV remapBilinear, 8u, fixed point):
//   (ix, iy) = top-left tap, (fx, fy) = 5-bit fractions,
//   w = {(32-fx)(32-fy), fx(32-fy), (32-fx)fy, fx fy} * 32      (15-bit, sums to 32768)
//   dst = (sum w_i * tap_i + 2^14) >> 15, taps outside the source read the border constant 0.
// HBM-bound: 6 map bytes + 1 written byte per pixel; the scattered 1-byte taps of neighbouring pixels fall
// into the same few sectors (rectification maps are smooth) and are served by L1/L2.  The maps of a camera
// (6 B/px) stay L2-resident across the images of a batch.
__global__ void __launch_bounds__(256)
remap_linear_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw, size_t src_pitch,
                       const uint32_t* __restrict__ xy, const uint16_t* __restrict__ frac, int npx,
                       uint8_t* __restrict__ dst, size_t dst_pitch) {
  pdl_wait();
  pdl_launch_dependents();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;   // group of four pixels
  const int i0 = q * 4;
  if (i0 >= npx) return;
  const uint8_t* s = src + static_cast<size_t>(blockIdx.y) * src_pitch;   // bytes between consecutive images
  uint8_t* d = dst + static_cast<size_t>(blockIdx.y) * dst_pitch;
  uint32_t pxy[4];
  uint16_t pfr[4];
  if (i0 + 3 < npx) {
    const uint4 a = *reinterpret_cast<const uint4*>(xy + i0);
    const uint2 b = *reinterpret_cast<const uint2*>(frac + i0);
    pxy[0] = a.x, pxy[1] = a.y, pxy[2] = a.z, pxy[3] = a.w;
    pfr[0] = b.x & 0xffffu, pfr[1] = b.x >> 16, pfr[2] = b.y & 0xffffu, pfr[3] = b.y >> 16;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pxy[k] = i0 + k < npx ? xy[i0 + k] : 0u;
      pfr[k] = i0 + k < npx ? frac[i0 + k] : 0;
    }
  }
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = static_cast<int16_t>(pxy[k] & 0xffffu);
    const int y = static_cast<int16_t>(pxy[k] >> 16);
    const int fx = pfr[k] & 31, fy = pfr[k] >> 5;
    int t00 = 0, t01 = 0, t10 = 0, t11 = 0;
    if (static_cast<unsigned>(x) < static_cast<unsigned>(sw - 1) && static_cast<unsigned>(y) < static_cast<unsigned>(sh - 1)) {
      const uint8_t* p = s + static_cast<size_t>(y) * sw + x;
      t00 = p[0], t01 = p[1], t10 = p[sw], t11 = p[sw + 1];
    } else {
      const bool x0 = x >= 0 && x < sw, x1 = x + 1 >= 0 && x + 1 < sw;
      const bool y0 = y >= 0 && y < sh, y1 = y + 1 >= 0 && y + 1 < sh;
      if (y0 && x0) t00 = s[static_cast<size_t>(y) * sw + x];
      if (y0 && x1) t01 = s[static_cast<size_t>(y) * sw + x + 1];
      if (y1 && x0) t10 = s[static_cast<size_t>(y + 1) * sw + x];
      if (y1 && x1) t11 = s[static_cast<size_t>(y + 1) * sw + x + 1];
    }
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32,
              w11 = fx * fy * 32;
    const int v = (t00 * w00 + t01 * w01 + t10 * w10 + t11 * w11 + (1 << 14)) >> 15;
    out |= static_cast<uint32_t>(min(max(v, 0), 255)) << (8 * k);
  }
  if (i0 + 3 < npx) {
    *reinterpret_cast<uint32_t*>(d + i0) = out;
  } else {
    for (int k = 0; k < 4 && i0 + k < npx; ++k) d[i0 + k] = static_cast<uint8_t>(out >> (8 * k));
  }
}

Rectifier::~Rectifier() {
  cudaSetDevice(device_);
  if (xy_) cudaFree(xy_);
  if (frac_) cudaFree(frac_);
  if (src_) cudaFree(src_);
  if (dst_) cudaFree(dst_);
  if (src_host_) cudaFreeHost(src_host_);
  if (dst_host_) cudaFreeHost(dst_host_);
  if (stream_) cudaStreamDestroy(stream_);
}

// cvRound(v * INTER_TAB_SIZE): round half to even; NaN / out-of-range convert to INT_MIN (cvtss2si)
static inline int cv_round_x32(float v) {
  const float t = v * 32.0f;
  if (!(std::fabs(t) < 2147483648.0f)) return INT_MIN;
  return static_cast<int>(std::nearbyintf(t));
}
static inline int sat16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

// cv::remap's float -> fixed-point map conversion (host): (int16 ix | int16 iy << 16) and fy * 32 + fx per pixel.
void convert_remap_maps(const float* map_x, const float* map_y, size_t n, uint32_t* xy, uint16_t* frac) {
  for (size_t i = 0; i < n; ++i) {
    const int sx = cv_round_x32(map_x[i]), sy = cv_round_x32(map_y[i]);
    const int ix = sat16(sx >> 5), iy = sat16(sy >> 5);   // arithmetic shift == floor division by 32
    xy[i] = (static_cast<uint32_t>(ix) & 0xffffu) | (static_cast<uint32_t>(iy) << 16);
    frac[i] = static_cast<uint16_t>((sy & 31) * 32 + (sx & 31));
  }
}

int Rectifier::init(const float* map_x, const float* map_y, int dst_h, int dst_w, int src_h, int src_w,
                    int max_images, int device) {
  SSB_CHECK(map_x && map_y, SSB_ERR_INVALID, "null map");
  SSB_CHECK(dst_h > 0 && dst_w > 0 && src_h > 0 && src_w > 0 && max_images > 0, SSB_ERR_INVALID, "bad sizes");
  SSB_CHECK(static_cast<long long>(dst_h) * dst_w % 4 == 0, SSB_ERR_INVALID,
            "destination pixel count must be a multiple of 4");
  device_ = device;
  dh_ = dst_h, dw_ = dst_w, sh_ = src_h, sw_ = src_w, cap_ = max_images;
  SSB_CUDA_CHECK(cudaSetDevice(device));
  SSB_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  const size_t n = static_cast<size_t>(dst_h) * dst_w;
  std::vector<uint32_t> hxy(n);
  std::vector<uint16_t> hfr(n);
  convert_remap_maps(map_x, map_y, n, hxy.data(), hfr.data());
  SSB_CUDA_CHECK(cudaMalloc(&xy_, n * 4));
  SSB_CUDA_CHECK(cudaMalloc(&frac_, n * 2));
  SSB_CUDA_CHECK(cudaMemcpy(xy_, hxy.data(), n * 4, cudaMemcpyHostToDevice));
  SSB_CUDA_CHECK(cudaMemcpy(frac_, hfr.data(), n * 2, cudaMemcpyHostToDevice));
  const size_t sbytes = static_cast<size_t>(src_h) * src_w * max_images, dbytes = n * max_images;
  SSB_CUDA_CHECK(cudaMalloc(&src_, sbytes));
  SSB_CUDA_CHECK(cudaMalloc(&dst_, dbytes));
  SSB_CUDA_CHECK(cudaMallocHost(&src_host_, sbytes));
  SSB_CUDA_CHECK(cudaMallocHost(&dst_host_, dbytes));
  return SSB_OK;
}

int Rectifier::remap_device(const uint8_t* src_dev, int count, uint8_t* dst_dev, cudaStream_t stream,
                            size_t src_pitch, size_t dst_pitch) {
  SSB_CHECK(src_dev && dst_dev && count >= 1, SSB_ERR_INVALID, "bad arguments");
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  const int npx = dh_ * dw_;
  if (src_pitch == 0) src_pitch = static_cast<size_t>(sh_) * sw_;
  if (dst_pitch == 0) dst_pitch = static_cast<size_t>(npx);
  SSB_CHECK(dst_pitch % 4 == 0, SSB_ERR_INVALID, "destination image pitch must be a multiple of 4 bytes");
  dim3 grid((npx / 4 + 255) / 256, count);
  SSB_CUDA_CHECK(launch_kernel(remap_linear_u8_kernel, dim3(grid), dim3(256), 0, stream, 1, src_dev, sh_, sw_, src_pitch, xy_, frac_, npx, dst_dev, dst_pitch));
  count_launch();
  prof_mark(stream, "fe.remap");
  return SSB_OK;
}

int Rectifier::remap(const uint8_t* const* images, int count, int row_stride, uint8_t* const* out) {
  SSB_CHECK(images && out && count >= 1 && count <= cap_, SSB_ERR_INVALID, "count %d exceeds capacity %d", count, cap_);
  SSB_CHECK(row_stride >= sw_, SSB_ERR_INVALID, "row_stride smaller than a row");
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  const size_t sb = static_cast<size_t>(sh_) * sw_, db = static_cast<size_t>(dh_) * dw_;
  for (int i = 0; i < count; ++i) {
    SSB_CHECK(images[i] && out[i], SSB_ERR_INVALID, "image %d is null", i);
    for (int y = 0; y < sh_; ++y)
      std::memcpy(src_host_ + i * sb + static_cast<size_t>(y) * sw_, images[i] + static_cast<size_t>(y) * row_stride, sw_);
  }
  SSB_CUDA_CHECK(cudaMemcpyAsync(src_, src_host_, sb * count, cudaMemcpyHostToDevice, stream_));
  SSB_RETURN_IF(remap_device(src_, count, dst_, stream_));
  SSB_CUDA_CHECK(cudaMemcpyAsync(dst_host_, dst_, db * count, cudaMemcpyDeviceToHost, stream_));
  SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
  for (int i = 0; i < count; ++i) std::memcpy(out[i], dst_host_ + i * db, db);
  return SSB_OK;
}

// =================================================================================================
// RGB-D keypoint post-process
// =================================================================================================

// One thread per keypoint, fp64 with explicit round-to-nearest operations so that no multiply-add is
// contracted: cv::undistortPoints runs as plain SSE2 doubles, and the final float must match bit for bit.
__global__ void rgbd_post_kernel(const float* __restrict__ xy, int n, const uint8_t* __restrict__ depth,
                                 int depth_type, int dh, int dw, RgbdParams p, float* __restrict__ oxy,
                                 double* __restrict__ stereo, uint8_t* __restrict__ has_depth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float rx = xy[2 * i], ry = xy[2 * i + 1];
  float ux = rx, uy = ry;
  if (p.has_dist) {
    const double u = rx, v = ry;
    const double ifx = __ddiv_rn(1.0, p.fx), ify = __ddiv_rn(1.0, p.fy);
    double x = __dmul_rn(__dsub_rn(u, p.cx), ifx), y = __dmul_rn(__dsub_rn(v, p.cy), ify);
    const double x0 = x, y0 = y;
    const double* k = p.k;
    for (int it = 0; it < 5; ++it) {
      const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
      const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
      const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
      const double icdist = __ddiv_rn(num, den);
      if (icdist < 0) {
        x = x0;
        y = y0;
        break;
      }
      // deltaX = 2*k2*x*y + k3*(r2 + 2*x*x) + k8*r2 + k9*r2*r2   (left to right)
      const double twoxx = __dmul_rn(__dmul_rn(2.0, x), x), twoyy = __dmul_rn(__dmul_rn(2.0, y), y);
      double dx = __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y);
      dx = __dadd_rn(dx, __dmul_rn(k[3], __dadd_rn(r2, twoxx)));
      dx = __dadd_rn(dx, __dmul_rn(k[8], r2));
      dx = __dadd_rn(dx, __dmul_rn(__dmul_rn(k[9], r2), r2));
      double dy = __dmul_rn(k[2], __dadd_rn(r2, twoyy));
      dy = __dadd_rn(dy, __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y));
      dy = __dadd_rn(dy, __dmul_rn(k[10], r2));
      dy = __dadd_rn(dy, __dmul_rn(__dmul_rn(k[11], r2), r2));
      x = __dmul_rn(__dsub_rn(x0, dx), icdist);
      y = __dmul_rn(__dsub_rn(y0, dy), icdist);
    }
    // P = K, R = I:  xx = fx*x + 0*y + cx,  ww = 1 / (0*x + 0*y + 1)
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(p.fx, x), __dmul_rn(0.0, y)), p.cx);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(p.fy, y)), p.cy);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
    ux = static_cast<float>(__dmul_rn(xx, ww));
    uy = static_cast<float>(__dmul_rn(yy, ww));
  }
  oxy[2 * i] = ux;
  oxy[2 * i + 1] = uy;
  // sampleDepth at lround(raw) (RgbdFrontEnd.cc:12-20,46-47)
  const long du = lround(static_cast<double>(rx)), dv = lround(static_cast<double>(ry));
  double z = 0.0;
  if (du >= 0 && dv >= 0 && du < dw && dv < dh) {
    const size_t o = static_cast<size_t>(dv) * dw + du;
    if (depth_type == 0) z = __ddiv_rn(static_cast<double>(reinterpret_cast<const uint16_t*>(depth)[o]), p.depth_factor);
    else z = __ddiv_rn(static_cast<double>(reinterpret_cast<const float*>(depth)[o]), p.depth_factor);
  }
  const double ul = ux, vv = uy;
  double ur = nan("");
  uint8_t hd = 0;
  if (z > 0.0 && z < p.max_depth) {
    ur = __dsub_rn(ul, __ddiv_rn(p.bf, z));
    hd = 1;
  }
  stereo[3 * i] = ul;
  stereo[3 * i + 1] = ur;
  stereo[3 * i + 2] = vv;
  has_depth[i] = hd;
}

RgbdPost::~RgbdPost() {
  cudaSetDevice(device_);
  void* bufs[] = {xy_, oxy_, stereo_, has_, depth_};
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (host_) cudaFreeHost(host_);
  if (stream_) cudaStreamDestroy(stream_);
}

int RgbdPost::init(int max_keypoints, int max_h, int max_w, int device) {
  SSB_CHECK(max_keypoints > 0 && max_h > 0 && max_w > 0, SSB_ERR_INVALID, "bad sizes");
  device_ = device, kmax_ = max_keypoints, hmax_ = max_h, wmax_ = max_w;
  SSB_CUDA_CHECK(cudaSetDevice(device));
  SSB_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  const size_t K = max_keypoints, dbytes = static_cast<size_t>(max_h) * max_w * 4;
  SSB_CUDA_CHECK(cudaMalloc(&xy_, K * 8));
  SSB_CUDA_CHECK(cudaMalloc(&oxy_, K * 8));
  SSB_CUDA_CHECK(cudaMalloc(&stereo_, K * 24));
  SSB_CUDA_CHECK(cudaMalloc(&has_, K));
  SSB_CUDA_CHECK(cudaMalloc(&depth_, dbytes));
  host_bytes_ = K * 8 + dbytes + K * 8 + K * 24 + K + 64;
  SSB_CUDA_CHECK(cudaMallocHost(&host_, host_bytes_));
  return SSB_OK;
}

int RgbdPost::process(const float* xy, int n, const void* depth, int depth_type, int dh, int dw, int row_stride,
                      const RgbdParams& p, float* out_xy, double* out_stereo, uint8_t* out_has_depth) {
  SSB_CHECK(n >= 0 && n <= kmax_, SSB_ERR_INVALID, "n %d exceeds max_keypoints %d", n, kmax_);
  if (n == 0) return SSB_OK;
  SSB_CHECK(xy && depth && out_xy && out_stereo && out_has_depth, SSB_ERR_INVALID, "null argument");
  SSB_CHECK(depth_type == 0 || depth_type == 1, SSB_ERR_INVALID, "depth_type must be 0 (u16) or 1 (f32)");
  SSB_CHECK(dh > 0 && dw > 0 && dh <= hmax_ && dw <= wmax_, SSB_ERR_INVALID, "depth image %dx%d exceeds %dx%d", dw, dh,
            wmax_, hmax_);
  const size_t esz = depth_type == 0 ? 2 : 4, rowb = static_cast<size_t>(dw) * esz;
  SSB_CHECK(static_cast<size_t>(row_stride) >= rowb, SSB_ERR_INVALID, "row_stride smaller than a row");
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  const size_t K = kmax_;
  uint8_t* h_xy = host_;
  uint8_t* h_depth = h_xy + K * 8;
  uint8_t* h_oxy = h_depth + static_cast<size_t>(hmax_) * wmax_ * 4;
  uint8_t* h_st = h_oxy + K * 8;
  uint8_t* h_has = h_st + K * 24;
  std::memcpy(h_xy, xy, static_cast<size_t>(n) * 8);
  for (int y = 0; y < dh; ++y)
    std::memcpy(h_depth + y * rowb, static_cast<const uint8_t*>(depth) + static_cast<size_t>(y) * row_stride, rowb);
  SSB_CUDA_CHECK(cudaMemcpyAsync(xy_, h_xy, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, stream_));
  SSB_CUDA_CHECK(cudaMemcpyAsync(depth_, h_depth, rowb * dh, cudaMemcpyHostToDevice, stream_));
  rgbd_post_kernel<<<(n + 127) / 128, 128, 0, stream_>>>(xy_, n, depth_, depth_type, dh, dw, p, oxy_, stereo_, has_);
  SSB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  SSB_CUDA_CHECK(cudaMemcpyAsync(h_oxy, oxy_, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, stream_));
  SSB_CUDA_CHECK(cudaMemcpyAsync(h_st, stereo_, static_cast<size_t>(n) * 24, cudaMemcpyDeviceToHost, stream_));
  SSB_CUDA_CHECK(cudaMemcpyAsync(h_has, has_, static_cast<size_t>(n), cudaMemcpyDeviceToHost, stream_));
  SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
  std::memcpy(out_xy, h_oxy, static_cast<size_t>(n) * 8);
  std::memcpy(out_stereo, h_st, static_cast<size_t>(n) * 24);
  std::memcpy(out_has_depth, h_has, static_cast<size_t>(n));
  return SSB_OK;
}

}  // namespace ssb
