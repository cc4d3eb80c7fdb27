// C-ABI shim that RUNS the reference's own SuperPoint::select_and_gather (/root/reference/src/SuperPoint.cc:681-750,
// SURVEY §8 rows a7-a9), compiled from the source where it lies: border / threshold scan of the score map, the
// std::sort on (score, (h, w)) pairs, top-K, keypoint scaling, nearest-cell indices, DescriptorPool::make and - with a
// GPU - the reference's gather kernel on a descriptor grid.  The rest of SuperPoint.cc is TensorRT plumbing; it
// compiles against the declaration-level stand-ins in oracle/stubs_trt/ and is never called (the stand-in members
// below fail like a missing engine).  The function is private: this file alone is compiled with -fno-access-control.
// Built by oracle/Makefile into oracle/_ref/libref_superpoint.so.  TEST INFRASTRUCTURE.
//
// Without a GPU the pool's cudaMalloc fails, DescriptorPool::make hands back a null slot pointer and the function
// returns false at its "pool exhausted" check - AFTER the keypoints were written, which is what the CPU test reads.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "SuperPoint.h"

// ---- "no TensorRT here" -------------------------------------------------------------------------------------------
namespace nvinfer1 {
bool IExecutionContext::setInputShape(const char*, const Dims&) { return false; }
Dims IExecutionContext::getTensorShape(const char*) const { return Dims(); }
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
  std::fprintf(stderr, "oracle/ref_superpoint_shim: %s is not part of the tested path\n", what);
  std::abort();
}
void cvtColor(const Mat&, Mat&, int) { not_on_this_path("cv::cvtColor"); }
void resize(const Mat&, Mat&, Size) { not_on_this_path("cv::resize"); }
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
}
