// C-ABI shim that RUNS the host logic of the reference's two TensorRT wrapper classes, compiled from the sources where
// they lie.  (1) SuperPoint::select_and_gather (/root/reference/src/SuperPoint.cc:681-750): border / threshold scan of the score map, the
// std::sort on (score, (h, w)) pairs, top-K, keypoint scaling, nearest-cell indices, DescriptorPool::make and - with a
// GPU - the reference's gather kernel on a descriptor grid.  The rest of SuperPoint.cc is TensorRT plumbing; it
// compiles against the declaration-level stand-ins in oracle/stubs_trt/ and is never called (the stand-in members
// below fail like a missing engine).  The function is private: this file alone is compiled with -fno-access-control.
// (2) LightGlue::prepare_inputs / store_keypoints / normalize_keypoints (src/LightGlue.cc:163-172, 228-283, row a10):
// what the engine is fed - keypoints normalised as (p - size/2) / (max(w, h)/2) in float, descriptors converted to the
// binding's type - and LightGlue::postprocess_outputs (:326-363, row a12): matches0 / mscores0 -> cv::DMatch.  The
// tensors are plain host buffers set up by this shim the way allocate_buffers would name and type them.
// (3) EigenPlaces::preprocess (src/EigenPlaces.cc:123-143) with cv::resize served by cv2 through a hook.
// Built by oracle/Makefile into oracle/_ref/libref_nethost.so.  TEST INFRASTRUCTURE.
//
// Without a GPU the pool's cudaMalloc fails, DescriptorPool::make hands back a null slot pointer and the function
// returns false at its "pool exhausted" check - AFTER the keypoints were written, which is what the CPU test reads.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include <cstring>

#include "EigenPlaces.h"
#include "LightGlue.h"
#include "SuperPoint.h"

// ---- "no TensorRT here" -------------------------------------------------------------------------------------------
static nvinfer1::Dims g_tensor_shape;   // what the stand-in context reports (set per call for "matches0")
namespace nvinfer1 {
bool IExecutionContext::setInputShape(const char*, const Dims&) { return false; }
Dims IExecutionContext::getTensorShape(const char*) const { return g_tensor_shape; }
bool IExecutionContext::setTensorAddress(const char*, void*) { return false; }
bool IExecutionContext::enqueueV3(cudaStream_t) { return false; }
IExecutionContext* ICudaEngine::createExecutionContext() { return nullptr; }
int32_t ICudaEngine::getNbIOTensors() const { return 0; }
const char* ICudaEngine::getIOTensorName(int32_t) const { return ""; }
DataType ICudaEngine::getTensorDataType(const char*) const { return DataType::kFLOAT; }
TensorIOMode ICudaEngine::getTensorIOMode(const char*) const { return TensorIOMode::kNONE; }
Dims ICudaEngine::getTensorShape(const char*) const { return Dims(); }
ICudaEngine* IRuntime::deserializeCudaEngine(const void*, std::size_t) { return nullptr; }
IRuntime* createInferRuntime(ILogger&) { return nullptr; }
}  // namespace nvinfer1
namespace cv {
static void not_on_this_path(const char* what) {
  std::fprintf(stderr, "oracle/ref_nethost_shim: %s is not part of the tested path\n", what);
  std::abort();
}
// cv::resize is served by the real OpenCV: the test installs a callback into cv2.resize (INTER_LINEAR, the default)
using ResizeFn = void (*)(const unsigned char* src, int src_h, int src_w, int channels, int src_step, unsigned char* dst,
                          int dst_h, int dst_w);
static ResizeFn g_resize = nullptr;
void resize(const Mat& src, Mat& dst, Size dsize) {
  if (!g_resize || src.depth() != CV_8U) not_on_this_path("cv::resize without a hook / on non-u8 data");
  Mat out(dsize.height, dsize.width, src.type());
  g_resize(src.data, src.rows, src.cols, src.channels(), static_cast<int>(src.step[0]), out.data, out.rows, out.cols);
  dst = out;
}
// the two channel reorderings of EigenPlaces::preprocess are exact copies; BGR2GRAY (SuperPoint's host path) is not served
void cvtColor(const Mat& src, Mat& dst, int code) {
  if (src.depth() != CV_8U || (code != COLOR_GRAY2RGB && code != COLOR_BGR2RGB)) not_on_this_path("cv::cvtColor (this code)");
  Mat out(src.rows, src.cols, CV_8UC3);
  for (int y = 0; y < src.rows; ++y) {
    const unsigned char* s = src.ptr<unsigned char>(y);
    unsigned char* d = out.ptr<unsigned char>(y);
    for (int x = 0; x < src.cols; ++x) {
      if (code == COLOR_GRAY2RGB) d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = s[x];
      else d[3 * x] = s[3 * x + 2], d[3 * x + 1] = s[3 * x + 1], d[3 * x + 2] = s[3 * x];
    }
  }
  dst = out;
}
void normalize(const Mat&, Mat&, double, double, int) { not_on_this_path("cv::normalize"); }
}  // namespace cv

extern "C" {

// The object as SuperPoint::initialize + allocate_dynamic_buffers leave it for a (input_h x input_w) image, minus the
// engine: pool of `descriptor_pool_slots` slots, the two cell-index buffers, a stream.  (Members set directly.)
void* ref_sp_new(int max_keypoints, double keypoint_threshold, int remove_borders, int input_h, int input_w) {
  SuperPoint* sp = new SuperPoint("none.engine", max_keypoints, keypoint_threshold, remove_borders);
  sp->input_height_ = input_h;
  sp->input_width_ = input_w;
  sp->pool_ = std::make_unique<superslam::DescriptorPool>(SuperPoint::descriptor_pool_slots, max_keypoints,
                                                          SuperPoint::descriptor_dim);
  if (cudaMalloc(&sp->cell_h_dev_, sizeof(int) * max_keypoints) != cudaSuccess) sp->cell_h_dev_ = nullptr;
  if (cudaMalloc(&sp->cell_w_dev_, sizeof(int) * max_keypoints) != cudaSuccess) sp->cell_w_dev_ = nullptr;
  if (cudaStreamCreate(&sp->stream_) != cudaSuccess) sp->stream_ = nullptr;
  cudaGetLastError();
  return sp;
}
void ref_sp_delete(void* h) {
  SuperPoint* sp = static_cast<SuperPoint*>(h);
  if (sp->cell_h_dev_) cudaFree(sp->cell_h_dev_);
  if (sp->cell_w_dev_) cudaFree(sp->cell_w_dev_);
  sp->cell_h_dev_ = sp->cell_w_dev_ = nullptr;
  delete sp;
}

// select_and_gather on a host fp32 score map [score_h, score_w] and a DEVICE fp16 descriptor grid [256, grid_h,
// grid_w] (may be null without a GPU).  Outputs sized max_keypoints: xy, response, (size, angle); desc_out (host,
// fp16 bits [n, 256]) is filled when the gather ran.  *ok = the function's return value; returns the keypoint count.
int ref_sp_select_and_gather(void* h, const float* scores_host, int score_h, int score_w, const void* grid_device,
                             int grid_h, int grid_w, float* xy, float* response, float* size_angle,
                             unsigned short* desc_out, int* ok, int* desc_info) {
  SuperPoint* sp = static_cast<SuperPoint*>(h);
  std::vector<cv::KeyPoint> kps;
  superslam::DeviceDescriptors d;
  *ok = sp->select_and_gather(scores_host, /*scores_half=*/false, score_h, score_w, grid_device,
                              SuperPoint::descriptor_dim, grid_h, grid_w, kps, d)
            ? 1
            : 0;
  const int n = static_cast<int>(kps.size());
  for (int i = 0; i < n; ++i) {
    xy[2 * i] = kps[i].pt.x, xy[2 * i + 1] = kps[i].pt.y;
    response[i] = kps[i].response;
    size_angle[2 * i] = kps[i].size, size_angle[2 * i + 1] = kps[i].angle;
  }
  desc_info[0] = d.count, desc_info[1] = d.dim, desc_info[2] = d.slot, desc_info[3] = d.data != nullptr;
  if (*ok && n > 0 && desc_out && d.data)
    cudaMemcpy(desc_out, d.data, sizeof(unsigned short) * static_cast<size_t>(n) * d.dim, cudaMemcpyDeviceToHost);
  return n;
}

// ---- LightGlue host logic -------------------------------------------------------------------------------------------
// prepare_inputs: kpts as fp32 or fp16 (kpts_half), descriptors as fp32 or fp16 (desc_half) host "bindings".
// out_k0 / out_k1: n*2 elements, out_d0 / out_d1: n*256 elements of the chosen width (4 or 2 bytes).
int ref_lg_prepare_inputs(int image_width, int image_height, const float* xy0, int n0, const float* desc0,
                          const float* xy1, int n1, const float* desc1, int kpts_half, int desc_half, void* out_k0,
                          void* out_k1, void* out_d0, void* out_d1) {
  LightGlue lg("none.engine", image_width, image_height);
  std::vector<cv::KeyPoint> k0, k1;
  for (int i = 0; i < n0; ++i) k0.emplace_back(xy0[2 * i], xy0[2 * i + 1], 1.0f);
  for (int i = 0; i < n1; ++i) k1.emplace_back(xy1[2 * i], xy1[2 * i + 1], 1.0f);
  cv::Mat d0(n0, 256, CV_32F, const_cast<float*>(desc0)), d1(n1, 256, CV_32F, const_cast<float*>(desc1));
  const char* names[4] = {"kpts0", "kpts1", "desc0", "desc1"};
  void* bufs[4] = {out_k0, out_k1, out_d0, out_d1};
  for (int i = 0; i < 4; ++i) {
    LightGlue::TensorInfo t;
    t.name = names[i];
    t.hostPtr = bufs[i];
    t.dtype = (i < 2 ? kpts_half : desc_half) ? nvinfer1::DataType::kHALF : nvinfer1::DataType::kFLOAT;
    lg.input_tensors_.push_back(t);
  }
  const bool ok = lg.prepare_inputs(k0, d0, k1, d1);
  lg.input_tensors_.clear();   // the buffers are the caller's: keep free_buffers away from them
  return ok ? 1 : 0;
}

// normalize_keypoints (the older float-only helper, :163-172)
void ref_lg_normalize_keypoints(int image_width, int image_height, const float* xy, int n, float* out) {
  LightGlue lg("none.engine", image_width, image_height);
  std::vector<cv::KeyPoint> k;
  for (int i = 0; i < n; ++i) k.emplace_back(xy[2 * i], xy[2 * i + 1], 1.0f);
  lg.normalize_keypoints(k, out);
}

// postprocess_outputs on host "output bindings": matches0 int32 [1, n0]; mscores0 fp32, fp16 (scores_kind 1) or
// absent (scores_kind 2).  Returns the number of matches (or -1 when the function returns false).
int ref_lg_postprocess(const int* matches0, const void* mscores0, int scores_kind, int n0, int* query, int* train,
                       float* distance) {
  LightGlue lg("none.engine", 640, 480);
  lg.context_.reset(new nvinfer1::IExecutionContext);
  LightGlue::TensorInfo m, s;
  m.name = "matches0", m.hostPtr = const_cast<int*>(matches0), m.dtype = nvinfer1::DataType::kINT32;
  s.name = "mscores0", s.hostPtr = const_cast<void*>(mscores0);
  s.dtype = scores_kind == 1 ? nvinfer1::DataType::kHALF : nvinfer1::DataType::kFLOAT;
  lg.output_tensors_.push_back(m);
  if (scores_kind != 2) lg.output_tensors_.push_back(s);
  g_tensor_shape = nvinfer1::Dims();
  g_tensor_shape.nbDims = 2, g_tensor_shape.d[0] = 1, g_tensor_shape.d[1] = n0;
  MatchResult r;
  r.matches.emplace_back(7, 7, 7.0f);   // must be cleared by the function
  const bool ok = lg.postprocess_outputs(r);
  lg.output_tensors_.clear();
  if (!ok) return -1;
  for (size_t i = 0; i < r.matches.size(); ++i) {
    query[i] = r.matches[i].queryIdx, train[i] = r.matches[i].trainIdx, distance[i] = r.matches[i].distance;
  }
  return static_cast<int>(r.matches.size());
}

// ---- EigenPlaces host logic: EigenPlaces::preprocess (src/EigenPlaces.cc:123-143) ---------------------
// gray / BGR u8 image -> RGB, cv::resize to (input_w, input_h) [served by cv2 through `resize`], convertTo(CV_32F,
// 1/255), ImageNet normalisation, HWC -> CHW.  out: float [3][input_h][input_w].
void ref_ep_preprocess(const unsigned char* image, int height, int width, int channels, int row_stride, int input_w,
                       int input_h, cv::ResizeFn resize, float* out) {
  cv::g_resize = resize;
  EigenPlaces ep("none.engine", input_w, input_h);
  const cv::Mat img(height, width, CV_MAKETYPE(CV_8U, channels), const_cast<unsigned char*>(image), row_stride);
  ep.preprocess(img, out);
  cv::g_resize = nullptr;
}
}
