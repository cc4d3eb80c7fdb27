// End-to-end run of the REFERENCE's own wrapper code on a machine with neither TensorRT nor a GPU:
//   /root/reference/src/SuperPoint.cc  LightGlue.cc  DescriptorPool.cc  StereoFrontEnd.cc   - compiled in place, unchanged,
// over (a) a functional stand-in for the handful of TensorRT calls they make, whose enqueueV3 hands the bound buffers to
// a callback (the test serves the SuperPoint / LightGlue graphs with the CPU oracle), (b) the eight CUDA runtime calls
// they make, implemented on host memory, and (c) launch_gather_descriptors forwarded to a callback (the reference's
// kernel itself is compared with the product on the GPU, tests/test_gpu_zz_ref_gather.py).
// Everything between the image and the StereoFrame - gray / 255 conversion, {2,1,H,W} packing, buffer (re)sizing, score
// and grid slice offsets, select_and_gather, DescriptorPool slots, keypoint normalisation, fp16 descriptor D2D copies,
// postprocess_outputs, the disparity / row filter - is the reference's code; tests/test_oracle_ref_e2e.py holds the
// oracle's composition of the restated pieces to it bit for bit.
// Built by oracle/Makefile (g++, no cudart) into oracle/_ref/libref_e2e.so.  TEST INFRASTRUCTURE.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "DescriptorGather.h"
#include "LightGlue.h"
#include "StereoFrontEnd.h"
#include "SuperPoint.h"

// ---- hooks ----------------------------------------------------------------------------------------------------------
using SpInferFn = void (*)(const float* image, int batch, int h, int w, float* scores, unsigned short* desc_f16);
using LgInferFn = void (*)(const float* k0, int n0, const unsigned short* d0, const float* k1, int n1,
                           const unsigned short* d1, int* matches0, float* mscores0);
using GatherFn = void (*)(const unsigned short* grid_chw, int channels, int gh, int gw, const int* cell_h, const int* cell_w,
                          int n, unsigned short* out);
static SpInferFn g_sp_infer = nullptr;
static LgInferFn g_lg_infer = nullptr;
static GatherFn g_gather = nullptr;
static long g_cuda_live = 0;   // outstanding cudaMalloc + cudaMallocHost blocks (leak check)

// ---- the CUDA runtime calls of the four files, on host memory ------------------------------------------------------------
extern "C" {
cudaError_t cudaMalloc(void** p, size_t n) {
  *p = std::malloc(n ? n : 1);
  ++g_cuda_live;
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFree(void* p) {
  if (p) --g_cuda_live;
  std::free(p);
  return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memcpy(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memcpy(d, s, n);   // "stream order" = program order
  return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s) {
  *s = reinterpret_cast<cudaStream_t>(new int(0));
  return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  delete reinterpret_cast<int*>(s);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
}

namespace superslam {
void launch_gather_descriptors(const void* grid_fp16, int channels, int grid_h, int grid_w, const int* cell_h,
                               const int* cell_w, int num_keypoints, void* out_fp16, cudaStream_t) {
  if (num_keypoints <= 0) return;
  g_gather(static_cast<const unsigned short*>(grid_fp16), channels, grid_h, grid_w, cell_h, cell_w, num_keypoints,
           static_cast<unsigned short*>(out_fp16));
}
}  // namespace superslam

// ---- TensorRT, functionally: two fixed "engines" with the bindings scripts/rebuild_engines.sh:88-120 gives them -------
namespace {
struct Binding {
  const char* name;
  nvinfer1::DataType dtype;
  bool input;
  nvinfer1::Dims dims;   // -1 = dynamic
};
nvinfer1::Dims dims_of(std::initializer_list<int64_t> v) {
  nvinfer1::Dims d;
  d.nbDims = static_cast<int32_t>(v.size());
  int i = 0;
  for (int64_t x : v) d.d[i++] = x;
  return d;
}
using DT = nvinfer1::DataType;
const std::vector<Binding> kSuperPoint = {{"input", DT::kFLOAT, true, dims_of({-1, 1, -1, -1})},
                                          {"scores", DT::kFLOAT, false, dims_of({-1, -1, -1})},
                                          {"descriptors", DT::kHALF, false, dims_of({-1, 256, -1, -1})}};
const std::vector<Binding> kLightGlue = {{"kpts0", DT::kFLOAT, true, dims_of({1, -1, 2})},   {"desc0", DT::kHALF, true, dims_of({1, -1, 256})},
                                         {"kpts1", DT::kFLOAT, true, dims_of({1, -1, 2})},   {"desc1", DT::kHALF, true, dims_of({1, -1, 256})},
                                         {"matches0", DT::kINT32, false, dims_of({1, -1})},  {"mscores0", DT::kFLOAT, false, dims_of({1, -1})}};
struct EngineState {
  const std::vector<Binding>* io;
};
struct ContextState {
  const std::vector<Binding>* io;
  std::map<std::string, nvinfer1::Dims> shape;
  std::map<std::string, void*> addr;
};
std::map<const void*, EngineState> g_engines;
std::map<const void*, ContextState> g_contexts;
const Binding* find(const std::vector<Binding>& io, const char* name) {
  for (const Binding& b : io)
    if (std::strcmp(b.name, name) == 0) return &b;
  return nullptr;
}
}  // namespace

namespace nvinfer1 {
IRuntime* createInferRuntime(ILogger&) { return new IRuntime; }
ICudaEngine* IRuntime::deserializeCudaEngine(const void* blob, std::size_t size) {
  const std::string tag(static_cast<const char*>(blob), size < 16 ? size : 16);
  const std::vector<Binding>* io = tag.rfind("superpoint", 0) == 0 ? &kSuperPoint : tag.rfind("lightglue", 0) == 0 ? &kLightGlue : nullptr;
  if (!io) return nullptr;   // "TensorRT version mismatch"
  ICudaEngine* e = new ICudaEngine;
  g_engines[e] = EngineState{io};
  return e;
}
IExecutionContext* ICudaEngine::createExecutionContext() {
  IExecutionContext* c = new IExecutionContext;
  g_contexts[c] = ContextState{g_engines.at(this).io, {}, {}};
  return c;
}
int32_t ICudaEngine::getNbIOTensors() const { return static_cast<int32_t>(g_engines.at(this).io->size()); }
const char* ICudaEngine::getIOTensorName(int32_t i) const { return (*g_engines.at(this).io)[i].name; }
DataType ICudaEngine::getTensorDataType(const char* n) const { return find(*g_engines.at(this).io, n)->dtype; }
TensorIOMode ICudaEngine::getTensorIOMode(const char* n) const {
  return find(*g_engines.at(this).io, n)->input ? TensorIOMode::kINPUT : TensorIOMode::kOUTPUT;
}
Dims ICudaEngine::getTensorShape(const char* n) const { return find(*g_engines.at(this).io, n)->dims; }
bool IExecutionContext::setInputShape(const char* n, const Dims& d) {
  ContextState& c = g_contexts.at(this);
  const Binding* b = find(*c.io, n);
  if (!b || !b->input || d.nbDims != b->dims.nbDims) return false;
  for (int i = 0; i < d.nbDims; ++i)
    if (d.d[i] <= 0 || (b->dims.d[i] > 0 && b->dims.d[i] != d.d[i])) return false;
  c.shape[n] = d;
  return true;
}
Dims IExecutionContext::getTensorShape(const char* n) const {
  const ContextState& c = g_contexts.at(this);
  const Binding* b = find(*c.io, n);
  if (!b) return Dims();
  if (b->input) {
    auto it = c.shape.find(n);
    return it == c.shape.end() ? b->dims : it->second;
  }
  if (c.io == &kSuperPoint) {
    auto it = c.shape.find("input");
    if (it == c.shape.end()) return b->dims;
    const int64_t B = it->second.d[0], hc = it->second.d[2] / 8, wc = it->second.d[3] / 8;
    return std::strcmp(n, "scores") == 0 ? dims_of({B, hc * 8, wc * 8}) : dims_of({B, 256, hc, wc});
  }
  auto it = c.shape.find("kpts0");
  return it == c.shape.end() ? b->dims : dims_of({1, it->second.d[1]});
}
bool IExecutionContext::setTensorAddress(const char* n, void* p) {
  ContextState& c = g_contexts.at(this);
  if (!find(*c.io, n) || p == nullptr) return false;
  c.addr[n] = p;
  return true;
}
bool IExecutionContext::enqueueV3(cudaStream_t) {
  ContextState& c = g_contexts.at(this);
  for (const Binding& b : *c.io)
    if (!c.addr.count(b.name) || (b.input && !c.shape.count(b.name))) return false;
  if (c.io == &kSuperPoint) {
    const Dims& d = c.shape.at("input");
    g_sp_infer(static_cast<const float*>(c.addr.at("input")), static_cast<int>(d.d[0]), static_cast<int>(d.d[2]),
               static_cast<int>(d.d[3]), static_cast<float*>(c.addr.at("scores")),
               static_cast<unsigned short*>(c.addr.at("descriptors")));
  } else {
    const int n0 = static_cast<int>(c.shape.at("kpts0").d[1]), n1 = static_cast<int>(c.shape.at("kpts1").d[1]);
    if (c.shape.at("desc0").d[1] != n0 || c.shape.at("desc1").d[1] != n1) return false;
    g_lg_infer(static_cast<const float*>(c.addr.at("kpts0")), n0, static_cast<const unsigned short*>(c.addr.at("desc0")),
               static_cast<const float*>(c.addr.at("kpts1")), n1, static_cast<const unsigned short*>(c.addr.at("desc1")),
               static_cast<int*>(c.addr.at("matches0")), static_cast<float*>(c.addr.at("mscores0")));
  }
  return true;
}
}  // namespace nvinfer1

namespace cv {   // preprocessing calls that a gray image at the engine's size never reaches
static void unreachable(const char* what) {
  std::fprintf(stderr, "oracle/ref_e2e_shim: %s is not part of the tested path\n", what);
  std::abort();
}
void cvtColor(const Mat&, Mat&, int) { unreachable("cv::cvtColor"); }
void resize(const Mat&, Mat&, Size) { unreachable("cv::resize"); }
void normalize(const Mat&, Mat&, double, double, int) { unreachable("cv::normalize"); }
}  // namespace cv

// ---- the object graph of src/SuperSLAM.cc:82-113 and the calls of one stereo frame ----------------------------------
namespace {
struct E2E {
  std::shared_ptr<SuperPoint> sp;
  std::shared_ptr<LightGlue> lg;
  std::unique_ptr<superslam::StereoFrontEnd> fe;
};
}  // namespace

extern "C" {
void ref_e2e_set_hooks(SpInferFn sp, LgInferFn lg, GatherFn gather) { g_sp_infer = sp, g_lg_infer = lg, g_gather = gather; }
long ref_e2e_live_allocations() { return g_cuda_live; }

// engine files: any readable files whose first bytes are "superpoint" / "lightglue".  status bits: 1 SuperPoint, 2 LightGlue
void* ref_e2e_create(const char* sp_engine, const char* lg_engine, int max_keypoints, double threshold, int remove_borders,
                     int lg_width, int lg_height, float min_disparity, int* status) {
  E2E* e = new E2E;
  e->sp = std::make_shared<SuperPoint>(sp_engine, max_keypoints, threshold, remove_borders);
  e->lg = std::make_shared<LightGlue>(lg_engine, lg_width, lg_height);
  *status = (e->sp->initialize() ? 1 : 0) | (e->lg->initialize() ? 2 : 0);
  e->fe.reset(new superslam::StereoFrontEnd(e->sp.get(), e->lg.get(), gtsam::Cal3_S2Stereo(500, 500, 0, 320, 240, 0.1),
                                            min_disparity));
  return e;
}
void ref_e2e_destroy(void* h) { delete static_cast<E2E*>(h); }

// One stereo frame: StereoFrontEnd::process, then the same extractor / matcher calls once more to expose what the frame
// does not keep (right keypoints, descriptors of both sides, the match list).  Arrays sized `cap` (descriptors cap x 256).
int ref_e2e_process(void* h, const unsigned char* left, const unsigned char* right, int height, int width, int row_stride,
                    int cap, int* counts /* n_left, n_right, n_matches */, float* xy_l, float* resp_l, unsigned short* desc_l,
                    float* xy_r, float* resp_r, unsigned short* desc_r, double* stereo, char* has_depth, int* query, int* train,
                    float* distance) {
  E2E* e = static_cast<E2E*>(h);
  const cv::Mat l(height, width, CV_8UC1, const_cast<unsigned char*>(left), row_stride);
  const cv::Mat r(height, width, CV_8UC1, const_cast<unsigned char*>(right), row_stride);
  const superslam::StereoFrame f = e->fe->process(l, r, 0.0);
  superslam::IFeatureExtractor* ext = e->sp.get();
  superslam::IFeatureMatcher* mat = e->lg.get();
  const std::pair<superslam::Features, superslam::Features> lr = ext->extract_stereo(l, r);
  const MatchResult m = mat->match(lr.first.keypoints, lr.first.descriptors, lr.second.keypoints, lr.second.descriptors);
  const int nl = static_cast<int>(f.keypoints_left.size()), nr = static_cast<int>(lr.second.keypoints.size());
  counts[0] = nl, counts[1] = nr, counts[2] = static_cast<int>(m.matches.size());
  if (nl > cap || nr > cap || static_cast<int>(lr.first.keypoints.size()) != nl) return -1;
  for (int i = 0; i < nl; ++i) {
    xy_l[2 * i] = f.keypoints_left[i].pt.x, xy_l[2 * i + 1] = f.keypoints_left[i].pt.y;
    resp_l[i] = f.keypoints_left[i].response;
    stereo[3 * i] = f.stereo[i].uL(), stereo[3 * i + 1] = f.stereo[i].uR(), stereo[3 * i + 2] = f.stereo[i].v();
    has_depth[i] = f.has_depth[i];
  }
  for (int i = 0; i < nr; ++i) {
    xy_r[2 * i] = lr.second.keypoints[i].pt.x, xy_r[2 * i + 1] = lr.second.keypoints[i].pt.y;
    resp_r[i] = lr.second.keypoints[i].response;
  }
  if (!f.descriptors_left.empty()) std::memcpy(desc_l, f.descriptors_left.data, sizeof(unsigned short) * 256 * nl);
  if (!lr.second.descriptors.empty()) std::memcpy(desc_r, lr.second.descriptors.data, sizeof(unsigned short) * 256 * nr);
  for (size_t i = 0; i < m.matches.size(); ++i)
    query[i] = m.matches[i].queryIdx, train[i] = m.matches[i].trainIdx, distance[i] = m.matches[i].distance;
  return nl;
}

// StereoFrontEnd::process alone (the timed call of bench.py's CPU legs); returns the number of left keypoints and, through
// n_depth, how many of them received a stereo depth.
int ref_e2e_process_only(void* h, const unsigned char* left, const unsigned char* right, int height, int width, int row_stride,
                         int* n_depth) {
  E2E* e = static_cast<E2E*>(h);
  const cv::Mat l(height, width, CV_8UC1, const_cast<unsigned char*>(left), row_stride);
  const cv::Mat r(height, width, CV_8UC1, const_cast<unsigned char*>(right), row_stride);
  const superslam::StereoFrame f = e->fe->process(l, r, 0.0);
  int nd = 0;
  for (char c : f.has_depth) nd += c != 0;
  if (n_depth) *n_depth = nd;
  return static_cast<int>(f.keypoints_left.size());
}

// The mono / loop-closure call shapes: IFeatureExtractor::extract on two images (infer_device: batch-1 dynamic shape),
// IFeatureMatcher::descriptors_to_host on both results (src/LightGlue.cc:460-475), then the HOST-descriptor match
// (src/LightGlue.cc:285-324: prepare_inputs converts the CV_32F rows to the fp16 binding).  desc_host_*: cap x 256 floats.
int ref_e2e_mono_and_host_match(void* h, const unsigned char* img0, const unsigned char* img1, int height, int width,
                                int row_stride, int cap, int* counts, float* xy0, float* resp0, float* desc_host0, float* xy1,
                                float* resp1, float* desc_host1, int* query, int* train, float* distance) {
  E2E* e = static_cast<E2E*>(h);
  superslam::IFeatureExtractor* ext = e->sp.get();
  superslam::IFeatureMatcher* mat = e->lg.get();
  const cv::Mat a(height, width, CV_8UC1, const_cast<unsigned char*>(img0), row_stride);
  const cv::Mat b(height, width, CV_8UC1, const_cast<unsigned char*>(img1), row_stride);
  const superslam::Features f0 = ext->extract(a), f1 = ext->extract(b);
  const cv::Mat d0 = mat->descriptors_to_host(f0.descriptors), d1 = mat->descriptors_to_host(f1.descriptors);
  const MatchResult m = mat->match(f0.keypoints, d0, f1.keypoints, d1);
  const int n0 = static_cast<int>(f0.keypoints.size()), n1 = static_cast<int>(f1.keypoints.size());
  counts[0] = n0, counts[1] = n1, counts[2] = static_cast<int>(m.matches.size());
  if (n0 > cap || n1 > cap || d0.rows != n0 || d1.rows != n1 || (n0 && (d0.cols != 256 || d0.type() != CV_32F))) return -1;
  for (int i = 0; i < n0; ++i) {
    xy0[2 * i] = f0.keypoints[i].pt.x, xy0[2 * i + 1] = f0.keypoints[i].pt.y, resp0[i] = f0.keypoints[i].response;
    std::memcpy(desc_host0 + 256 * i, d0.ptr<float>(i), 256 * sizeof(float));
  }
  for (int i = 0; i < n1; ++i) {
    xy1[2 * i] = f1.keypoints[i].pt.x, xy1[2 * i + 1] = f1.keypoints[i].pt.y, resp1[i] = f1.keypoints[i].response;
    std::memcpy(desc_host1 + 256 * i, d1.ptr<float>(i), 256 * sizeof(float));
  }
  for (size_t i = 0; i < m.matches.size(); ++i)
    query[i] = m.matches[i].queryIdx, train[i] = m.matches[i].trainIdx, distance[i] = m.matches[i].distance;
  return n0;
}
}
