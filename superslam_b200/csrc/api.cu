// C-ABI of the library (include/superslam_b200.h): opaque handles over ssb::SuperPoint / ssb::LightGlue
// plus the chained frame-pair front end.  No exception and no C++ type crosses this boundary.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <vector>

#include "../../include/superslam_b200.h"
#include "eigenplaces.cuh"
#include "imgproc.cuh"
#include "lightglue.cuh"
#include "superpoint.cuh"

struct ssb_superpoint {
  ssb::SuperPoint impl;
};
struct ssb_lightglue {
  ssb::LightGlue impl;
};
struct ssb_eigenplaces {
  ssb::EigenPlaces impl;
};
struct ssb_rectifier {
  ssb::Rectifier impl;
};
struct ssb_rgbd {
  ssb::RgbdPost impl;
};

namespace ssb {

// StereoFrontEnd::process post-filter (src/StereoFrontEnd.cc:22-47) on the device: default
// (uL, NaN, v) / has_depth 0; a match is kept iff uL-uR >= min_disparity and |vL-vR| <= 2.
__global__ void stereo_postfilter_kernel(const float* __restrict__ kp_xy, int K, const int* __restrict__ cnt,
                                         const int32_t* __restrict__ matches0, int kp, float min_disp,
                                         float* __restrict__ ur, uint8_t* __restrict__ has_depth) {
  pdl_wait();
  pdl_launch_dependents();
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const size_t o = static_cast<size_t>(pair) * K + i;
  float u = nanf("");
  uint8_t hd = 0;
  const int n0 = cnt[2 * pair], n1 = cnt[2 * pair + 1];
  if (i < n0) {
    const int j = matches0[static_cast<size_t>(pair) * kp + i];
    if (j >= 0 && j < n1) {
      const float* l = kp_xy + (static_cast<size_t>(2 * pair) * K + i) * 2;
      const float* r = kp_xy + (static_cast<size_t>(2 * pair + 1) * K + j) * 2;
      if (l[0] - r[0] >= min_disp && fabsf(l[1] - r[1]) <= 2.0f) {
        u = r[0];
        hd = 1;
      }
    }
  }
  ur[o] = u;
  has_depth[o] = hd;
}

// ---- tracking chain (src/VoEstimator.cc:240-246): last keyframe <-> current left image -------------
// LightGlue::run takes its pairs as images (2p, 2p+1) of one keypoint array: even rows = the retained keyframe of
// stream p, odd rows = the left image SuperPoint has just extracted for stream p.
__global__ void tracking_assemble_kernel(const float* __restrict__ kf_xy, const int* __restrict__ kf_count,
                                         const float* __restrict__ kp_xy, const int* __restrict__ kp_count, int K,
                                         float* __restrict__ trk_xy, int* __restrict__ trk_count) {
  pdl_wait();
  pdl_launch_dependents();
  const int z = blockIdx.y, p = z >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool frame = (z & 1) != 0;
  const float2* src = reinterpret_cast<const float2*>(frame ? kp_xy + static_cast<size_t>(2 * p) * K * 2
                                                            : kf_xy + static_cast<size_t>(p) * K * 2);
  if (i < K) reinterpret_cast<float2*>(trk_xy)[static_cast<size_t>(z) * K + i] = src[i];
  if (i == 0) trk_count[z] = frame ? kp_count[2 * p] : kf_count[p];
}
// The part of VoEstimator::track's match loop that needs no pose (src/VoEstimator.cc:250-260): a tracking match
// (keyframe feature i -> frame feature j) is usable iff both ends carry stereo depth.
__global__ void tracking_postfilter_kernel(const int32_t* __restrict__ matches0, int kp, const int* __restrict__ trk_count,
                                           const uint8_t* __restrict__ kf_hd, const uint8_t* __restrict__ hd, int K,
                                           uint8_t* __restrict__ ok) {
  pdl_wait();
  pdl_launch_dependents();
  const int p = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  uint8_t r = 0;
  if (i < trk_count[2 * p]) {
    const int j = matches0[static_cast<size_t>(p) * kp + i];
    if (j >= 0 && j < trk_count[2 * p + 1]) r = kf_hd[static_cast<size_t>(p) * K + i] & hd[static_cast<size_t>(p) * K + j];
  }
  ok[static_cast<size_t>(p) * K + i] = r;
}
// "last_keyframe_ = frame" (src/VoEstimator.cc:327) for the streams whose mask byte is set: keypoints, count, depth
// flags and the fp16 descriptor rows of the current left image move into the stream's keyframe slot, device to device.
// grid (K / 8, pairs), block 256: one warp per descriptor row (512 bytes = 32 lanes x 16 bytes).
__global__ void __launch_bounds__(256)
promote_keyframe_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ kp_xy,
                        const int* __restrict__ kp_count, const uint8_t* __restrict__ hd,
                        void* const* __restrict__ slot_ptrs, int K, float* __restrict__ kf_xy,
                        int* __restrict__ kf_count, uint8_t* __restrict__ kf_hd, __half* __restrict__ kf_desc,
                        size_t kf_desc_stride) {
  pdl_wait();
  pdl_launch_dependents();
  const int p = blockIdx.y;
  if (mask[p] == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= K) return;
  const int n = kp_count[2 * p];
  const uint4* src = reinterpret_cast<const uint4*>(static_cast<const __half*>(slot_ptrs[2 * p]) + static_cast<size_t>(row) * 256);
  uint4* dst = reinterpret_cast<uint4*>(kf_desc + p * kf_desc_stride + static_cast<size_t>(row) * 256);
  dst[lane] = row < n ? src[lane] : make_uint4(0u, 0u, 0u, 0u);
  if (lane == 0) {
    const size_t o = static_cast<size_t>(p) * K + row;
    reinterpret_cast<float2*>(kf_xy)[o] = reinterpret_cast<const float2*>(kp_xy)[static_cast<size_t>(2 * p) * K + row];
    kf_hd[o] = row < n ? hd[o] : 0;
    if (row == 0) kf_count[p] = n;
  }
}

class FrontEnd {
 public:
  ~FrontEnd() {
    cudaSetDevice(device_);
    for (auto& g : graphs_)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& e : events_)
      if (e) cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b) {
      if (ev_up_[b]) cudaEventDestroy(ev_up_[b]);
      if (ev_used_[b]) cudaEventDestroy(ev_used_[b]);
      if (ev_done_[b]) cudaEventDestroy(ev_done_[b]);
      if (s_img_[b]) cudaFree(s_img_[b]);
      if (s_img_host_[b]) cudaFreeHost(s_img_host_[b]);
      if (s_host_[b]) cudaFreeHost(s_host_[b]);
    }
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    void* trk[] = {kf_xy_, kf_count_, kf_hd_, kf_desc_, trk_xy_, trk_count_, trk_desc_ptrs_, trk_ok_, promote_mask_};
    for (void* p : trk)
      if (p) cudaFree(p);
    if (rect_buf_) cudaFree(rect_buf_);
    if (ur_) cudaFree(ur_);
    if (hd_) cudaFree(hd_);
    if (img_dev_) cudaFree(img_dev_);
    if (slot_ptrs_) cudaFree(slot_ptrs_);
    if (host_) cudaFreeHost(host_);
    if (host_img_) cudaFreeHost(host_img_);
  }
  int init(const char* spw, const char* lgw, int K, double thr, int rb, int lw, int lh, float min_disp,
           int max_pairs, int device) {
    SSB_CHECK(max_pairs >= 1 && max_pairs <= 64, SSB_ERR_INVALID, "max_pairs out of range (1..64)");
    device_ = device;
    K_ = K;
    pairs_ = max_pairs;
    min_disp_ = min_disp;
    SSB_RETURN_IF(sp.impl.init(spw, K, thr, rb, 2 * max_pairs, device));
    auto w = std::make_shared<LgWeights>();
    SSB_RETURN_IF(w->load(lgw, device));
    SSB_RETURN_IF(lg.impl.init(w, lw, lh, K, max_pairs));
    stream_ = sp.impl.stream();
    // fixed descriptor slots: image i of a call always lands in slot i
    std::vector<void*>& ptrs = slot_host_;
    for (int i = 0; i < 2 * max_pairs; ++i) {
      const int s = sp.impl.pool().acquire();
      SSB_CHECK(s >= 0, SSB_ERR_EXHAUSTED, "front end could not reserve descriptor slots");
      ptrs.push_back(sp.impl.pool().slot_ptr(s));
    }
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&slot_ptrs_), ptrs.size() * sizeof(void*)));
    SSB_CUDA_CHECK(cudaMemcpy(slot_ptrs_, ptrs.data(), ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice));
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&ur_), static_cast<size_t>(max_pairs) * K * 4));
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&hd_), static_cast<size_t>(max_pairs) * K));
    host_bytes_ = result_bytes(max_pairs);
    SSB_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&host_), host_bytes_));
    for (auto& e : events_) SSB_CUDA_CHECK(cudaEventCreate(&e));
    return SSB_OK;
  }
  // pinned result block: count | xy | score | matches | mscores | ur | has_depth | (tracking, same call: keyframe
  // count | tracking matches | scores | usable flags)
  size_t result_bytes(int pairs) const {
    const size_t P = pairs, K = K_;
    return 2 * P * 4 + 2 * P * K * 8 + 2 * P * K * 4 + P * K * 4 + P * K * 4 + P * K * 4 + P * K + 64 +
           P * 4 + P * K * 4 + P * K * 4 + P * K + 64;
  }
  size_t tracking_offset(int pairs) const {   // of the tracking part inside a result block (16-byte aligned)
    const size_t P = pairs, K = K_;
    return (2 * P * 4 + 2 * P * K * 8 + 2 * P * K * 4 + P * K * 4 + P * K * 4 + P * K * 4 + P * K + 15) / 16 * 16;
  }

  // ---- tracking chain: SP x2 + LG (stereo) + LG (last keyframe <-> left) in one captured graph ------------------
  // Every pair slot p of a call is a stream with its own retained keyframe (count 0 until the first promotion: the
  // reference makes no tracking match for the first frame either, src/VoEstimator.cc:206-236).
  int enable_tracking(bool on) {
    SSB_CHECK(submitted_ == collected_, SSB_ERR_INVALID, "streamed steps are in flight: collect them first");
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
    drop_graphs();
    if (!on) {
      tracking_ = false;
      return SSB_OK;
    }
    if (kf_xy_ == nullptr) {
      const size_t P = pairs_, K = K_;
      SSB_RETURN_IF(lg_trk.impl.init(lg.impl.weights(), lg.impl.image_width(), lg.impl.image_height(), K_, pairs_));
      kf_desc_stride_ = static_cast<size_t>((K_ + 127) / 128 * 128) * 256;   // whole 128-row TMA boxes per stream
      auto alloc = [&](void** p, size_t bytes) -> int {
        SSB_CUDA_CHECK(cudaMalloc(p, bytes));
        SSB_CUDA_CHECK(cudaMemset(*p, 0, bytes));
        return SSB_OK;
      };
      SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&kf_xy_), P * K * 8));
      SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&kf_count_), P * 4));
      SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&kf_hd_), P * K));
      SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&kf_desc_), P * kf_desc_stride_ * 2 + 65536));
      SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&trk_xy_), 2 * P * K * 8));
      SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&trk_count_), 2 * P * 4));
      SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&trk_desc_ptrs_), 2 * P * sizeof(void*)));
      SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&trk_ok_), P * K));
      SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&promote_mask_), P));
      std::vector<void*> ptrs(2 * P);
      for (size_t p = 0; p < P; ++p) {
        ptrs[2 * p] = kf_desc_ + p * kf_desc_stride_;
        ptrs[2 * p + 1] = slot_host_[2 * p];   // the left image of pair p always lands in this slot
      }
      SSB_CUDA_CHECK(cudaMemcpy(trk_desc_ptrs_, ptrs.data(), ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice));
    }
    tracking_ = true;
    return SSB_OK;
  }
  int reset_tracking() {   // forget every keyframe
    SSB_CHECK(kf_count_ != nullptr, SSB_ERR_INVALID, "tracking was never enabled");
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    SSB_CUDA_CHECK(cudaMemsetAsync(kf_count_, 0, static_cast<size_t>(pairs_) * 4, stream_));
    return SSB_OK;
  }
  // mask[p] != 0: the left image of pair p of the LAST call becomes stream p's keyframe (nullptr: every stream).
  int promote_keyframes(const uint8_t* mask, int pairs) {
    SSB_CHECK(tracking_, SSB_ERR_INVALID, "tracking is not enabled");
    SSB_CHECK(pairs >= 1 && pairs <= pairs_, SSB_ERR_INVALID, "pairs %d exceeds capacity %d", pairs, pairs_);
    SSB_CHECK(submitted_ == collected_, SSB_ERR_INVALID,
              "a streamed step is in flight: its features would be promoted instead of the last collected ones");
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    uint8_t m[64];
    for (int p = 0; p < pairs; ++p) m[p] = mask == nullptr ? 1 : (mask[p] != 0);
    SSB_CUDA_CHECK(cudaMemcpyAsync(promote_mask_, m, pairs, cudaMemcpyHostToDevice, stream_));   // m: see the sync below
    SSB_CUDA_CHECK(launch_kernel(promote_keyframe_kernel, dim3(dim3((K_ + 7) / 8, pairs)), dim3(256), 0, stream_, 1, 
        promote_mask_, sp.impl.kp_xy(), sp.impl.kp_count(), hd_, slot_ptrs_, K_, kf_xy_, kf_count_, kf_hd_, kf_desc_,
        kf_desc_stride_));
    count_launch();
    SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
    return SSB_OK;
  }
  // Tracking outputs of the step whose results were delivered last (ssb_fe_process / fetch / collect).
  int tracking_results(int pairs, int* kf_count, int32_t* matches0, float* mscores0, uint8_t* ok) const {
    SSB_CHECK(tracking_ && last_block_ != nullptr && last_block_tracking_, SSB_ERR_INVALID,
              "no tracking results: enable tracking, then process / fetch / collect a step");
    SSB_CHECK(pairs == last_block_pairs_, SSB_ERR_INVALID, "the last delivered step had %d pairs, not %d",
              last_block_pairs_, pairs);
    const size_t P = pairs, K = K_;
    const uint8_t* t = last_block_ + tracking_offset(pairs);
    const int* h_kc = reinterpret_cast<const int*>(t);
    const int32_t* h_m = reinterpret_cast<const int32_t*>(t + (P * 4 + 15) / 16 * 16);
    const float* h_ms = reinterpret_cast<const float*>(h_m + P * K);
    const uint8_t* h_ok = reinterpret_cast<const uint8_t*>(h_ms + P * K);
    if (kf_count) std::memcpy(kf_count, h_kc, P * 4);
    if (matches0) std::memcpy(matches0, h_m, P * K * 4);
    if (mscores0) std::memcpy(mscores0, h_ms, P * K * 4);
    if (ok) std::memcpy(ok, h_ok, P * K);
    return SSB_OK;
  }
  bool tracking() const { return tracking_; }
  // extract-only mode: SuperPoint on 2 * pairs independent frames, no matching (the match / stereo outputs of a
  // call are then undefined)
  int set_extract_only(bool on) {
    SSB_CHECK(submitted_ == collected_, SSB_ERR_INVALID, "streamed steps are in flight: collect them first");
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
    drop_graphs();
    extract_only_ = on;
    return SSB_OK;
  }
  bool extract_only() const { return extract_only_; }
  // The whole pair pipeline (94 kernels, no host dependency thanks to device-side counts) is
  // captured into a CUDA graph the second time a (images, pairs, h, w) combination is seen and
  // replayed afterwards; SSB_NO_GRAPH=1 or an active event profiler falls back to eager launches.
  int enqueue_device(const uint8_t* images_dev, int pairs, int h, int w) {
    SSB_CHECK(pairs >= 1 && pairs <= pairs_, SSB_ERR_INVALID, "pairs %d exceeds capacity %d", pairs, pairs_);
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    // A captured graph holds SuperPoint's activation pointers and tensor maps: size the workspace first and drop
    // every graph if that moved it (a larger batch or another image size since the capture).
    {
      const int eh = rect_l_ ? rect_l_->dst_h() : h, ew = rect_l_ ? rect_l_->dst_w() : w;
      SSB_RETURN_IF(sp.impl.prepare(2 * pairs, eh, ew));
      if (sp.impl.shape_generation() != sp_gen_) {
        SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
        drop_graphs();
        sp_gen_ = sp.impl.shape_generation();
      }
    }
    static const bool no_graph = std::getenv("SSB_NO_GRAPH") != nullptr;
    if (no_graph || prof_enabled()) return enqueue_eager(images_dev, pairs, h, w);
    // graph cache keyed by (image buffer, pairs, h, w): the resident-input bench, process() and the two
    // buffers of the streaming path each get their own entry
    GraphEntry* e = nullptr;
    for (auto& g : graphs_)
      if (g.img == images_dev && g.pairs == pairs && g.h == h && g.w == w) e = &g;
    if (e != nullptr && e->exec != nullptr) {
      SSB_CUDA_CHECK(cudaGraphLaunch(e->exec, stream_));
      count_launch(e->kernels);
      return SSB_OK;
    }
    if (e == nullptr) {  // first sighting: run eagerly (allocations, function attributes), remember the key
      GraphEntry& g = graphs_[graph_next_++ % kGraphSlots];
      if (g.exec) cudaGraphExecDestroy(g.exec);
      g = GraphEntry{};
      g.img = images_dev, g.pairs = pairs, g.h = h, g.w = w;
      return enqueue_eager(images_dev, pairs, h, w);
    }
    if (e->failed) return enqueue_eager(images_dev, pairs, h, w);
    SSB_CUDA_CHECK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
    const long long before = launch_count();
    const int st = enqueue_eager(images_dev, pairs, h, w);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(stream_, &graph);
    if (st != SSB_OK || ce != cudaSuccess || graph == nullptr) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      e->failed = true;  // do not try again for this key
      if (st != SSB_OK) return st;
      set_last_error("CUDA graph capture failed: %s", cudaGetErrorString(ce));
      return SSB_ERR_CUDA;
    }
    e->kernels = static_cast<int>(launch_count() - before);
    const cudaError_t ie = cudaGraphInstantiate(&e->exec, graph, 0);
    cudaGraphDestroy(graph);
    SSB_CHECK(ie == cudaSuccess, SSB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
    SSB_CUDA_CHECK(cudaGraphLaunch(e->exec, stream_));
    return SSB_OK;
  }
  // Optional rectification in front of SuperPoint (EuRoC flow, examples/stereo/euroc.cc:176-181): image 2p goes
  // through the left camera's maps, image 2p+1 through the right camera's, device to device.
  int set_rectifiers(Rectifier* left, Rectifier* right) {
    SSB_CHECK((left == nullptr) == (right == nullptr), SSB_ERR_INVALID, "set both rectifiers or neither");
    SSB_CHECK(submitted_ == collected_, SSB_ERR_INVALID, "streamed steps are in flight: collect them first");
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
    drop_graphs();   // captured graphs bake the old configuration in
    if (rect_buf_) cudaFree(rect_buf_);
    rect_buf_ = nullptr;
    rect_l_ = rect_r_ = nullptr;
    if (left == nullptr) return SSB_OK;
    SSB_CHECK(left->device() == device_ && right->device() == device_, SSB_ERR_INVALID, "rectifiers live on another device");
    SSB_CHECK(left->dst_h() == right->dst_h() && left->dst_w() == right->dst_w() && left->src_h() == right->src_h() &&
                  left->src_w() == right->src_w(),
              SSB_ERR_INVALID, "left and right rectifiers must agree on source and destination sizes");
    SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&rect_buf_),
                              static_cast<size_t>(2 * pairs_) * left->dst_h() * left->dst_w()));
    rect_l_ = left;
    rect_r_ = right;
    return SSB_OK;
  }
  bool has_rectifiers() const { return rect_l_ != nullptr; }
  void drop_graphs() {
    for (auto& g : graphs_) {
      if (g.exec) cudaGraphExecDestroy(g.exec);
      g = GraphEntry{};
    }
  }
  int enqueue_eager(const uint8_t* images_dev, int pairs, int h, int w) {
    PdlScope pdl(2 * pairs);   // programmatic dependent launch for latency-bound (small) calls only
    prof_begin(stream_);
    if (rect_l_ != nullptr) {
      SSB_CHECK(h == rect_l_->src_h() && w == rect_l_->src_w(), SSB_ERR_INVALID,
                "images are %dx%d but the rectifiers expect %dx%d", w, h, rect_l_->src_w(), rect_l_->src_h());
      const size_t sp_ = static_cast<size_t>(h) * w, dp = static_cast<size_t>(rect_l_->dst_h()) * rect_l_->dst_w();
      SSB_RETURN_IF(rect_l_->remap_device(images_dev, pairs, rect_buf_, stream_, 2 * sp_, 2 * dp));
      SSB_RETURN_IF(rect_r_->remap_device(images_dev + sp_, pairs, rect_buf_ + dp, stream_, 2 * sp_, 2 * dp));
      images_dev = rect_buf_;
      h = rect_l_->dst_h();
      w = rect_l_->dst_w();
    }
    SSB_RETURN_IF(sp.impl.run(images_dev, 2 * pairs, h, w, slot_ptrs_, stream_));
    if (extract_only_) return SSB_OK;   // mono frames (BASELINE config C1): the 2 * pairs images are independent
    SSB_RETURN_IF(lg.impl.run(pairs, sp.impl.kp_xy(), K_, sp.impl.kp_count(), slot_ptrs_, stream_));
    SSB_CUDA_CHECK(launch_kernel(stereo_postfilter_kernel, dim3(dim3((K_ + 255) / 256, pairs)), dim3(256), 0, stream_, 1, 
        sp.impl.kp_xy(), K_, sp.impl.kp_count(), lg.impl.matches_dev(), lg.impl.kp(), min_disp_, ur_, hd_));
    count_launch();
    prof_mark(stream_, "fe.postfilter");
    if (tracking_) {
      SSB_CUDA_CHECK(launch_kernel(tracking_assemble_kernel, dim3(dim3((K_ + 255) / 256, 2 * pairs)), dim3(256), 0, stream_, 1, 
          kf_xy_, kf_count_, sp.impl.kp_xy(), sp.impl.kp_count(), K_, trk_xy_, trk_count_));
      count_launch();
      prof_mark(stream_, "fe.track_assemble");
      SSB_RETURN_IF(lg_trk.impl.run(pairs, trk_xy_, K_, trk_count_, trk_desc_ptrs_, stream_));
      SSB_CUDA_CHECK(launch_kernel(tracking_postfilter_kernel, dim3(dim3((K_ + 255) / 256, pairs)), dim3(256), 0, stream_, 1, 
          lg_trk.impl.matches_dev(), lg_trk.impl.kp(), trk_count_, kf_hd_, hd_, K_, trk_ok_));
      count_launch();
      prof_mark(stream_, "fe.track_postfilter");
    }
    return SSB_OK;
  }
  // Result block layout in pinned memory: count | xy | score | matches | mscores | ur | has_depth
  int enqueue_results(uint8_t* h, int pairs) {
    const size_t P = pairs, K = K_, KP = lg.impl.kp();
    int* h_cnt = reinterpret_cast<int*>(h);
    float* h_xy = reinterpret_cast<float*>(h + 2 * P * 4);
    float* h_sc = h_xy + 2 * P * K * 2;
    int32_t* h_m = reinterpret_cast<int32_t*>(h_sc + 2 * P * K);
    float* h_ms = reinterpret_cast<float*>(h_m + P * K);
    float* h_ur = h_ms + P * K;
    uint8_t* h_hd = reinterpret_cast<uint8_t*>(h_ur + P * K);
    SSB_CUDA_CHECK(cudaMemcpyAsync(h_cnt, sp.impl.kp_count(), 2 * P * 4, cudaMemcpyDeviceToHost, stream_));
    SSB_CUDA_CHECK(cudaMemcpyAsync(h_xy, sp.impl.kp_xy(), 2 * P * K * 8, cudaMemcpyDeviceToHost, stream_));
    SSB_CUDA_CHECK(cudaMemcpyAsync(h_sc, sp.impl.kp_score(), 2 * P * K * 4, cudaMemcpyDeviceToHost, stream_));
    SSB_CUDA_CHECK(cudaMemcpy2DAsync(h_m, K * 4, lg.impl.matches_dev(), KP * 4, K * 4, P, cudaMemcpyDeviceToHost, stream_));
    SSB_CUDA_CHECK(cudaMemcpy2DAsync(h_ms, K * 4, lg.impl.mscores_dev(), KP * 4, K * 4, P, cudaMemcpyDeviceToHost, stream_));
    SSB_CUDA_CHECK(cudaMemcpyAsync(h_ur, ur_, P * K * 4, cudaMemcpyDeviceToHost, stream_));
    SSB_CUDA_CHECK(cudaMemcpyAsync(h_hd, hd_, P * K, cudaMemcpyDeviceToHost, stream_));
    if (tracking_) {
      const size_t KT = lg_trk.impl.kp();
      uint8_t* t = h + tracking_offset(pairs);
      int32_t* t_m = reinterpret_cast<int32_t*>(t + (P * 4 + 15) / 16 * 16);
      float* t_ms = reinterpret_cast<float*>(t_m + P * K);
      uint8_t* t_ok = reinterpret_cast<uint8_t*>(t_ms + P * K);
      SSB_CUDA_CHECK(cudaMemcpy2DAsync(t, 4, trk_count_, 8, 4, P, cudaMemcpyDeviceToHost, stream_));   // even entries
      SSB_CUDA_CHECK(cudaMemcpy2DAsync(t_m, K * 4, lg_trk.impl.matches_dev(), KT * 4, K * 4, P, cudaMemcpyDeviceToHost, stream_));
      SSB_CUDA_CHECK(cudaMemcpy2DAsync(t_ms, K * 4, lg_trk.impl.mscores_dev(), KT * 4, K * 4, P, cudaMemcpyDeviceToHost, stream_));
      SSB_CUDA_CHECK(cudaMemcpyAsync(t_ok, trk_ok_, P * K, cudaMemcpyDeviceToHost, stream_));
    }
    return SSB_OK;
  }
  void copy_out(const uint8_t* h, int pairs, int* count, float* xy, float* score, int32_t* matches0,
                float* mscores0, float* ur, uint8_t* hd) const {
    const size_t P = pairs, K = K_;
    const int* h_cnt = reinterpret_cast<const int*>(h);
    const float* h_xy = reinterpret_cast<const float*>(h + 2 * P * 4);
    const float* h_sc = h_xy + 2 * P * K * 2;
    const int32_t* h_m = reinterpret_cast<const int32_t*>(h_sc + 2 * P * K);
    const float* h_ms = reinterpret_cast<const float*>(h_m + P * K);
    const float* h_ur = h_ms + P * K;
    const uint8_t* h_hd = reinterpret_cast<const uint8_t*>(h_ur + P * K);
    if (count) std::memcpy(count, h_cnt, 2 * P * 4);
    if (xy) std::memcpy(xy, h_xy, 2 * P * K * 8);
    if (score) std::memcpy(score, h_sc, 2 * P * K * 4);
    if (matches0) std::memcpy(matches0, h_m, P * K * 4);
    if (mscores0) std::memcpy(mscores0, h_ms, P * K * 4);
    if (ur) std::memcpy(ur, h_ur, P * K * 4);
    if (hd) std::memcpy(hd, h_hd, P * K);
  }
  int fetch(int pairs, int* count, float* xy, float* score, int32_t* matches0, float* mscores0, float* ur,
            uint8_t* hd) {
    SSB_CHECK(pairs >= 1 && pairs <= pairs_, SSB_ERR_INVALID, "pairs %d exceeds capacity %d", pairs, pairs_);
    SSB_CHECK(submitted_ == collected_, SSB_ERR_INVALID, "fetch while streamed steps are in flight: collect them first");
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    SSB_RETURN_IF(enqueue_results(host_, pairs));
    SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
    copy_out(host_, pairs, count, xy, score, matches0, mscores0, ur, hd);
    last_block_ = host_, last_block_pairs_ = pairs, last_block_tracking_ = tracking_;
    return SSB_OK;
  }

  // ---- streaming: submit(i+1) may be called before collect(i); the upload of step i+1 (copy stream, second
  // image buffer) then overlaps the kernels of step i, and results come back through per-slot pinned blocks.
  int submit(const uint8_t* const* images, int pairs, int h, int w, int row_stride) {
    SSB_CHECK(images != nullptr && row_stride >= w, SSB_ERR_INVALID, "bad image arguments");
    SSB_CHECK(pairs >= 1 && pairs <= pairs_, SSB_ERR_INVALID, "pairs %d exceeds capacity %d", pairs, pairs_);
    SSB_CHECK(submitted_ - collected_ < 2, SSB_ERR_INVALID, "two steps already in flight: collect one first");
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    const int b = static_cast<int>(submitted_ & 1);
    const size_t bytes = static_cast<size_t>(2 * pairs) * h * w;
    if (copy_stream_ == nullptr) {
      SSB_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        SSB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_up_[i], cudaEventDisableTiming));
        SSB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_used_[i], cudaEventDisableTiming));
        SSB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_done_[i], cudaEventDisableTiming));
        SSB_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&s_host_[i]), host_bytes_));
      }
    }
    if (bytes > s_img_bytes_[b]) {
      if (s_img_used_[b]) SSB_CUDA_CHECK(cudaEventSynchronize(ev_used_[b]));
      if (s_img_[b]) cudaFree(s_img_[b]);
      if (s_img_host_[b]) cudaFreeHost(s_img_host_[b]);
      s_img_[b] = nullptr, s_img_host_[b] = nullptr, s_img_bytes_[b] = 0;
      SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s_img_[b]), bytes));
      SSB_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&s_img_host_[b]), bytes));
      s_img_bytes_[b] = bytes;
    }
    // the kernels of step i-2 have finished reading this image buffer
    if (s_img_used_[b]) SSB_CUDA_CHECK(cudaStreamWaitEvent(copy_stream_, ev_used_[b], 0));
    bool pinned = true;
    for (int i = 0; i < 2 * pairs && pinned; ++i) {
      SSB_CHECK(images[i] != nullptr, SSB_ERR_INVALID, "image %d is null", i);
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, images[i]) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        pinned = false;
      }
    }
    const size_t img_bytes = static_cast<size_t>(h) * w;
    if (pinned) {
      for (int i = 0; i < 2 * pairs; ++i)
        SSB_CUDA_CHECK(cudaMemcpy2DAsync(s_img_[b] + i * img_bytes, w, images[i], row_stride, w, h,
                                         cudaMemcpyHostToDevice, copy_stream_));
    } else {
      if (s_img_used_[b]) SSB_CUDA_CHECK(cudaEventSynchronize(ev_up_[b]));   // staging of step i-2 has been read
      for (int i = 0; i < 2 * pairs; ++i) {
        SSB_CHECK(images[i] != nullptr, SSB_ERR_INVALID, "image %d is null", i);
        for (int y = 0; y < h; ++y)
          std::memcpy(s_img_host_[b] + i * img_bytes + static_cast<size_t>(y) * w, images[i] + static_cast<size_t>(y) * row_stride, w);
      }
      SSB_CUDA_CHECK(cudaMemcpyAsync(s_img_[b], s_img_host_[b], bytes, cudaMemcpyHostToDevice, copy_stream_));
    }
    SSB_CUDA_CHECK(cudaEventRecord(ev_up_[b], copy_stream_));
    SSB_CUDA_CHECK(cudaStreamWaitEvent(stream_, ev_up_[b], 0));
    SSB_RETURN_IF(enqueue_device(s_img_[b], pairs, h, w));
    SSB_CUDA_CHECK(cudaEventRecord(ev_used_[b], stream_));
    s_img_used_[b] = true;
    SSB_RETURN_IF(enqueue_results(s_host_[b], pairs));
    SSB_CUDA_CHECK(cudaEventRecord(ev_done_[b], stream_));
    s_pairs_[b] = pairs;
    ++submitted_;
    return SSB_OK;
  }
  int collect(int* pairs_out, int* count, float* xy, float* score, int32_t* matches0, float* mscores0, float* ur,
              uint8_t* hd) {
    SSB_CHECK(collected_ < submitted_, SSB_ERR_INVALID, "nothing in flight");
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    const int b = static_cast<int>(collected_ & 1);
    SSB_CUDA_CHECK(cudaEventSynchronize(ev_done_[b]));
    copy_out(s_host_[b], s_pairs_[b], count, xy, score, matches0, mscores0, ur, hd);
    last_block_ = s_host_[b], last_block_pairs_ = s_pairs_[b], last_block_tracking_ = tracking_;
    if (pairs_out) *pairs_out = s_pairs_[b];
    ++collected_;
    return SSB_OK;
  }
  int stage_images(const uint8_t* const* images, int count, int h, int w, int row_stride, bool own_copy,
                   uint8_t** out) {
    SSB_CHECK(images != nullptr && row_stride >= w, SSB_ERR_INVALID, "bad image arguments");
    SSB_CUDA_CHECK(cudaSetDevice(device_));
    const size_t bytes = static_cast<size_t>(count) * h * w;
    if (bytes > host_img_bytes_) {
      SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
      if (host_img_) cudaFreeHost(host_img_);
      if (img_dev_) cudaFree(img_dev_);
      host_img_ = nullptr, img_dev_ = nullptr;
      SSB_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&host_img_), bytes));
      SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&img_dev_), bytes));
      host_img_bytes_ = bytes;
    }
    uint8_t* dst = img_dev_;
    if (own_copy) {  // caller keeps it (bench: inputs resident in HBM); freed with the process
      SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&dst), bytes));
    }
    // Images that already live in pinned host memory go to the device directly (one strided async copy
    // each); pageable images are packed into the pinned staging buffer first.
    bool pinned = true;
    for (int i = 0; i < count && pinned; ++i) {
      SSB_CHECK(images[i] != nullptr, SSB_ERR_INVALID, "image %d is null", i);
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, images[i]) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        pinned = false;
      }
    }
    if (pinned) {
      for (int i = 0; i < count; ++i)
        SSB_CUDA_CHECK(cudaMemcpy2DAsync(dst + static_cast<size_t>(i) * h * w, w, images[i], row_stride, w, h,
                                         cudaMemcpyHostToDevice, stream_));
    } else {
      for (int i = 0; i < count; ++i) {
        SSB_CHECK(images[i] != nullptr, SSB_ERR_INVALID, "image %d is null", i);
        for (int y = 0; y < h; ++y)
          std::memcpy(host_img_ + (static_cast<size_t>(i) * h + y) * w, images[i] + static_cast<size_t>(y) * row_stride, w);
      }
      SSB_CUDA_CHECK(cudaMemcpyAsync(dst, host_img_, bytes, cudaMemcpyHostToDevice, stream_));
    }
    if (own_copy) SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
    *out = dst;
    return SSB_OK;
  }

  ssb_superpoint sp;
  ssb_lightglue lg;
  ssb_lightglue lg_trk;            // second context (shared weights) for the keyframe <-> frame match
  bool tracking_ = false;
  bool extract_only_ = false;
  float* kf_xy_ = nullptr;         // [pairs_][K][2] keyframe keypoints per stream
  int* kf_count_ = nullptr;        // [pairs_]
  uint8_t* kf_hd_ = nullptr;       // [pairs_][K] stereo-depth flags of the keyframe's features
  __half* kf_desc_ = nullptr;      // [pairs_][kf_desc_stride_] fp16 descriptor rows
  size_t kf_desc_stride_ = 0;
  float* trk_xy_ = nullptr;        // [2 pairs_][K][2]: (keyframe, frame) per stream, LightGlue::run's layout
  int* trk_count_ = nullptr;       // [2 pairs_]
  void** trk_desc_ptrs_ = nullptr; // [2 pairs_]
  uint8_t* trk_ok_ = nullptr;      // [pairs_][K]
  uint8_t* promote_mask_ = nullptr;
  std::vector<void*> slot_host_;   // host copy of slot_ptrs_
  const uint8_t* last_block_ = nullptr;   // pinned result block delivered last
  int last_block_pairs_ = 0;
  bool last_block_tracking_ = false;
  Rectifier* rect_l_ = nullptr;   // not owned
  Rectifier* rect_r_ = nullptr;
  uint8_t* rect_buf_ = nullptr;   // [2 * pairs_][dst_h][dst_w]
  struct GraphEntry {
    const uint8_t* img = nullptr;
    int pairs = 0, h = 0, w = 0, kernels = 0;
    bool failed = false;
    cudaGraphExec_t exec = nullptr;
  };
  static constexpr int kGraphSlots = 4;
  GraphEntry graphs_[kGraphSlots];
  unsigned graph_next_ = 0;
  unsigned long long sp_gen_ = 0;   // SuperPoint::shape_generation() the cached graphs were captured against
  cudaStream_t stream_ = nullptr;
  // streaming state (submit / collect)
  cudaStream_t copy_stream_ = nullptr;
  cudaEvent_t ev_up_[2] = {}, ev_used_[2] = {}, ev_done_[2] = {};
  uint8_t* s_img_[2] = {};
  uint8_t* s_img_host_[2] = {};
  size_t s_img_bytes_[2] = {};
  bool s_img_used_[2] = {};
  uint8_t* s_host_[2] = {};
  int s_pairs_[2] = {};
  unsigned long long submitted_ = 0, collected_ = 0;
  cudaEvent_t events_[16] = {};
  int device_ = 0, K_ = 0, pairs_ = 0;
  float min_disp_ = 1.0f;
  void** slot_ptrs_ = nullptr;
  float* ur_ = nullptr;
  uint8_t* hd_ = nullptr;
  uint8_t* host_ = nullptr;
  size_t host_bytes_ = 0;
  uint8_t* host_img_ = nullptr;
  uint8_t* img_dev_ = nullptr;
  size_t host_img_bytes_ = 0;
};

}  // namespace ssb

struct ssb_frontend {
  ssb::FrontEnd impl;
};

using namespace ssb;

#define SSB_API_BEGIN try {
#define SSB_API_END                                                  \
  }                                                                  \
  catch (const std::bad_alloc&) {                                    \
    ssb::set_last_error("out of host memory");                       \
    return SSB_ERR_INVALID;                                          \
  }                                                                  \
  catch (...) {                                                      \
    ssb::set_last_error("unexpected C++ exception");                 \
    return SSB_ERR_INVALID;                                          \
  }

extern "C" {

const char* ssb_last_error(void) { return ssb::last_error(); }
int ssb_version(void) { return 100; }
long long ssb_kernel_launch_count(void) { return ssb::launch_count(); }
void ssb_profile_enable(int on) { ssb::prof_enable(on != 0); }
void ssb_profile_collect(void) { ssb::prof_collect(); }
int ssb_profile_report(char* buf, size_t bytes) { return ssb::prof_report(buf, bytes); }

int ssb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    ssb::set_last_error("cudaGetDeviceCount failed (no driver / no device)");
    cudaGetLastError();
    return -SSB_ERR_NODEVICE;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

int ssb_sp_create(const char* weights_path, int max_keypoints, double keypoint_threshold, int remove_borders,
                  int num_slots, int device_id, ssb_superpoint** out) {
  SSB_API_BEGIN
  SSB_CHECK(out != nullptr, SSB_ERR_INVALID, "out is null");
  *out = nullptr;
  std::unique_ptr<ssb_superpoint> h(new ssb_superpoint);
  SSB_RETURN_IF(h->impl.init(weights_path, max_keypoints, keypoint_threshold, remove_borders, num_slots, device_id));
  *out = h.release();
  return SSB_OK;
  SSB_API_END
}
void ssb_sp_destroy(ssb_superpoint* sp) { delete sp; }

int ssb_sp_extract(ssb_superpoint* sp, const uint8_t* const* images, int batch, int height, int width,
                   int row_stride, int channels, float* const* xy, float* const* score, int* count,
                   void** desc_dev, int* slot) {
  SSB_API_BEGIN
  SSB_CHECK(sp != nullptr, SSB_ERR_INVALID, "sp is null");
  return sp->impl.extract(images, batch, height, width, row_stride, channels, xy, score, count, desc_dev, slot);
  SSB_API_END
}
int ssb_sp_slot_retain(ssb_superpoint* sp, int slot) {
  SSB_CHECK(sp != nullptr, SSB_ERR_INVALID, "sp is null");
  return sp->impl.pool().retain(slot);
}
int ssb_sp_slot_release(ssb_superpoint* sp, int slot) {
  SSB_CHECK(sp != nullptr, SSB_ERR_INVALID, "sp is null");
  return sp->impl.pool().release(slot);
}
int ssb_sp_slots_in_use(ssb_superpoint* sp) { return sp ? sp->impl.pool().in_use() : -1; }
int ssb_sp_max_keypoints(ssb_superpoint* sp) { return sp ? sp->impl.max_keypoints() : -1; }
int ssb_sp_debug_read(ssb_superpoint* sp, const char* what, void* dst, size_t bytes) {
  SSB_CHECK(sp != nullptr && dst != nullptr, SSB_ERR_INVALID, "null argument");
  return sp->impl.debug_read(what, dst, bytes);
}

int ssb_lg_create(const char* weights_path, int image_width, int image_height, int max_keypoints,
                  int device_id, ssb_lightglue** out) {
  SSB_API_BEGIN
  SSB_CHECK(out != nullptr, SSB_ERR_INVALID, "out is null");
  *out = nullptr;
  SSB_CUDA_CHECK(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  SSB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device_id));
  SSB_CHECK(prop.major == 10, SSB_ERR_NODEVICE, "device %d is sm_%d%d; this library needs sm_100", device_id,
            prop.major, prop.minor);
  auto w = std::make_shared<ssb::LgWeights>();
  SSB_RETURN_IF(w->load(weights_path, device_id));
  std::unique_ptr<ssb_lightglue> h(new ssb_lightglue);
  SSB_RETURN_IF(h->impl.init(w, image_width, image_height, max_keypoints, 1));
  *out = h.release();
  return SSB_OK;
  SSB_API_END
}
int ssb_lg_clone_context(ssb_lightglue* src, int image_width, int image_height, ssb_lightglue** out) {
  SSB_API_BEGIN
  SSB_CHECK(src != nullptr && out != nullptr, SSB_ERR_INVALID, "null argument");
  *out = nullptr;
  std::unique_ptr<ssb_lightglue> h(new ssb_lightglue);
  SSB_RETURN_IF(h->impl.init(src->impl.weights(), image_width, image_height, src->impl.max_keypoints(), 1));
  *out = h.release();
  return SSB_OK;
  SSB_API_END
}
void ssb_lg_destroy(ssb_lightglue* lg) { delete lg; }

int ssb_lg_match_device(ssb_lightglue* lg, const float* xy0, int n0, const void* desc0_dev, const float* xy1,
                        int n1, const void* desc1_dev, int32_t* matches0, float* mscores0) {
  SSB_API_BEGIN
  SSB_CHECK(lg != nullptr, SSB_ERR_INVALID, "lg is null");
  return lg->impl.match_device(xy0, n0, desc0_dev, xy1, n1, desc1_dev, matches0, mscores0);
  SSB_API_END
}
int ssb_lg_match_host(ssb_lightglue* lg, const float* xy0, int n0, const float* desc0_f32, const float* xy1,
                      int n1, const float* desc1_f32, int32_t* matches0, float* mscores0) {
  SSB_API_BEGIN
  SSB_CHECK(lg != nullptr, SSB_ERR_INVALID, "lg is null");
  return lg->impl.match_host(xy0, n0, desc0_f32, xy1, n1, desc1_f32, matches0, mscores0);
  SSB_API_END
}
int ssb_desc_to_host_f32(int device_id, const void* desc_dev_f16, int count, int dim, float* out) {
  SSB_API_BEGIN
  if (desc_dev_f16 == nullptr || count <= 0) return SSB_OK;  // empty handle -> empty Mat
  SSB_CHECK(out != nullptr && dim > 0, SSB_ERR_INVALID, "bad arguments");
  SSB_CUDA_CHECK(cudaSetDevice(device_id));
  const size_t n = static_cast<size_t>(count) * dim;
  std::vector<__half> tmp(n);
  SSB_CUDA_CHECK(cudaMemcpy(tmp.data(), desc_dev_f16, n * sizeof(__half), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) out[i] = __half2float(tmp[i]);
  return SSB_OK;
  SSB_API_END
}
int ssb_lg_debug_read(ssb_lightglue* lg, const char* what, void* dst, size_t bytes) {
  SSB_CHECK(lg != nullptr && dst != nullptr, SSB_ERR_INVALID, "null argument");
  return lg->impl.debug_read(what, dst, bytes);
}

int ssb_fe_create(const char* sp_weights, const char* lg_weights, int max_keypoints, double keypoint_threshold,
                  int remove_borders, int lg_image_width, int lg_image_height, float min_disparity,
                  int max_pairs, int device_id, ssb_frontend** out) {
  SSB_API_BEGIN
  SSB_CHECK(out != nullptr, SSB_ERR_INVALID, "out is null");
  *out = nullptr;
  std::unique_ptr<ssb_frontend> h(new ssb_frontend);
  SSB_RETURN_IF(h->impl.init(sp_weights, lg_weights, max_keypoints, keypoint_threshold, remove_borders,
                             lg_image_width, lg_image_height, min_disparity, max_pairs, device_id));
  *out = h.release();
  return SSB_OK;
  SSB_API_END
}
void ssb_fe_destroy(ssb_frontend* fe) { delete fe; }
int ssb_fe_set_rectifiers(ssb_frontend* fe, ssb_rectifier* left, ssb_rectifier* right) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  return fe->impl.set_rectifiers(left ? &left->impl : nullptr, right ? &right->impl : nullptr);
  SSB_API_END
}

int ssb_fe_process(ssb_frontend* fe, const uint8_t* const* images, int pairs, int height, int width,
                   int row_stride, int* count, float* xy, float* score, int32_t* matches0, float* mscores0,
                   float* stereo_ur, uint8_t* has_depth) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  uint8_t* dev = nullptr;
  SSB_RETURN_IF(fe->impl.stage_images(images, 2 * pairs, height, width, row_stride, false, &dev));
  SSB_RETURN_IF(fe->impl.enqueue_device(dev, pairs, height, width));
  return fe->impl.fetch(pairs, count, xy, score, matches0, mscores0, stereo_ur, has_depth);
  SSB_API_END
}
int ssb_fe_enqueue_device(ssb_frontend* fe, const uint8_t* images_dev, int pairs, int height, int width) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr && images_dev != nullptr, SSB_ERR_INVALID, "null argument");
  return fe->impl.enqueue_device(images_dev, pairs, height, width);
  SSB_API_END
}
int ssb_fe_fetch(ssb_frontend* fe, int pairs, int* count, float* xy, float* score, int32_t* matches0,
                 float* mscores0, float* stereo_ur, uint8_t* has_depth) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  return fe->impl.fetch(pairs, count, xy, score, matches0, mscores0, stereo_ur, has_depth);
  SSB_API_END
}
int ssb_fe_submit(ssb_frontend* fe, const uint8_t* const* images, int pairs, int height, int width,
                  int row_stride) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  return fe->impl.submit(images, pairs, height, width, row_stride);
  SSB_API_END
}
int ssb_fe_collect(ssb_frontend* fe, int* pairs, int* count, float* xy, float* score, int32_t* matches0,
                   float* mscores0, float* stereo_ur, uint8_t* has_depth) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  return fe->impl.collect(pairs, count, xy, score, matches0, mscores0, stereo_ur, has_depth);
  SSB_API_END
}
int ssb_fe_set_extract_only(ssb_frontend* fe, int on) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  return fe->impl.set_extract_only(on != 0);
  SSB_API_END
}
int ssb_fe_enable_tracking(ssb_frontend* fe, int enable) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  return fe->impl.enable_tracking(enable != 0);
  SSB_API_END
}
int ssb_fe_reset_tracking(ssb_frontend* fe) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  return fe->impl.reset_tracking();
  SSB_API_END
}
int ssb_fe_promote_keyframes(ssb_frontend* fe, const uint8_t* promote, int pairs) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  return fe->impl.promote_keyframes(promote, pairs);
  SSB_API_END
}
int ssb_fe_tracking_results(ssb_frontend* fe, int pairs, int* keyframe_count, int32_t* track_matches0,
                            float* track_mscores0, uint8_t* track_usable) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  return fe->impl.tracking_results(pairs, keyframe_count, track_matches0, track_mscores0, track_usable);
  SSB_API_END
}
int ssb_fe_sync(ssb_frontend* fe) {
  SSB_CHECK(fe != nullptr, SSB_ERR_INVALID, "fe is null");
  SSB_CUDA_CHECK(cudaSetDevice(fe->impl.device_));
  SSB_CUDA_CHECK(cudaStreamSynchronize(fe->impl.stream_));
  return SSB_OK;
}
int ssb_fe_event_record(ssb_frontend* fe, int index) {
  SSB_CHECK(fe != nullptr && index >= 0 && index < 16, SSB_ERR_INVALID, "bad event index");
  SSB_CUDA_CHECK(cudaSetDevice(fe->impl.device_));
  SSB_CUDA_CHECK(cudaEventRecord(fe->impl.events_[index], fe->impl.stream_));
  return SSB_OK;
}
int ssb_fe_event_elapsed_ms(ssb_frontend* fe, int start_index, int stop_index, float* ms) {
  SSB_CHECK(fe != nullptr && ms != nullptr && start_index >= 0 && start_index < 16 && stop_index >= 0 &&
                stop_index < 16,
            SSB_ERR_INVALID, "bad arguments");
  SSB_CUDA_CHECK(cudaSetDevice(fe->impl.device_));
  SSB_CUDA_CHECK(cudaEventSynchronize(fe->impl.events_[stop_index]));
  SSB_CUDA_CHECK(cudaEventElapsedTime(ms, fe->impl.events_[start_index], fe->impl.events_[stop_index]));
  return SSB_OK;
}
int ssb_fe_upload_images(ssb_frontend* fe, const uint8_t* const* images, int count, int height, int width,
                         int row_stride, uint8_t** images_dev_out) {
  SSB_API_BEGIN
  SSB_CHECK(fe != nullptr && images_dev_out != nullptr, SSB_ERR_INVALID, "null argument");
  return fe->impl.stage_images(images, count, height, width, row_stride, true, images_dev_out);
  SSB_API_END
}
int ssb_fe_kernel_launches_per_call(ssb_frontend* fe, int pairs) {
  (void)pairs;
  // SuperPoint: 10 convolution launches + nms, select, gather | LightGlue: prepare + 9 x (qkv, attention, ffn1,
  // ffn2) x 2 + final_proj, matchability, sim, sim^T, lse, arg-max, mutual | post-filter | (+ 2 remaps)
  const int lg_blocks = ssb::kLgLayers * 8 + (fe != nullptr && !fe->impl.lg.impl.weights()->fold_out ? ssb::kLgLayers * 2 : 0);
  const int lg_all = 1 + lg_blocks + 8;   // prepare | blocks | final_proj, matchability, sim, 2 sweeps + 2 merges, mutual
  if (fe != nullptr && fe->impl.extract_only()) return 13 + (fe->impl.has_rectifiers() ? 2 : 0);
  return 13 + lg_all + 1 + (fe != nullptr && fe->impl.has_rectifiers() ? 2 : 0) +
         (fe != nullptr && fe->impl.tracking() ? lg_all + 2 : 0);
}
static int require_sm100(int device_id) {
  SSB_CUDA_CHECK(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  SSB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device_id));
  SSB_CHECK(prop.major == 10, SSB_ERR_NODEVICE, "device %d is sm_%d%d; this library needs sm_100", device_id,
            prop.major, prop.minor);
  return SSB_OK;
}
int ssb_rect_create(const float* map_x, const float* map_y, int dst_height, int dst_width, int src_height,
                    int src_width, int max_images, int device_id, ssb_rectifier** out) {
  SSB_API_BEGIN
  SSB_CHECK(out != nullptr, SSB_ERR_INVALID, "out is null");
  *out = nullptr;
  SSB_RETURN_IF(require_sm100(device_id));
  std::unique_ptr<ssb_rectifier> h(new ssb_rectifier);
  SSB_RETURN_IF(h->impl.init(map_x, map_y, dst_height, dst_width, src_height, src_width,
                             max_images <= 0 ? 2 : max_images, device_id));
  *out = h.release();
  return SSB_OK;
  SSB_API_END
}
void ssb_rect_destroy(ssb_rectifier* r) { delete r; }
int ssb_rect_convert_maps(const float* map_x, const float* map_y, size_t count, uint32_t* xy, uint16_t* frac) {
  SSB_API_BEGIN
  SSB_CHECK(map_x && map_y && xy && frac, SSB_ERR_INVALID, "null argument");
  ssb::convert_remap_maps(map_x, map_y, count, xy, frac);
  return SSB_OK;
  SSB_API_END
}
int ssb_rect_remap(ssb_rectifier* r, const uint8_t* const* images, int count, int row_stride,
                   uint8_t* const* out) {
  SSB_API_BEGIN
  SSB_CHECK(r != nullptr, SSB_ERR_INVALID, "rectifier is null");
  return r->impl.remap(images, count, row_stride, out);
  SSB_API_END
}
int ssb_rect_remap_device(ssb_rectifier* r, const uint8_t* src_dev, int count, uint8_t* dst_dev) {
  SSB_API_BEGIN
  SSB_CHECK(r != nullptr, SSB_ERR_INVALID, "rectifier is null");
  SSB_RETURN_IF(r->impl.remap_device(src_dev, count, dst_dev, r->impl.stream()));
  SSB_CUDA_CHECK(cudaStreamSynchronize(r->impl.stream()));
  return SSB_OK;
  SSB_API_END
}
int ssb_rgbd_create(int max_keypoints, int max_height, int max_width, int device_id, ssb_rgbd** out) {
  SSB_API_BEGIN
  SSB_CHECK(out != nullptr, SSB_ERR_INVALID, "out is null");
  *out = nullptr;
  SSB_RETURN_IF(require_sm100(device_id));
  std::unique_ptr<ssb_rgbd> h(new ssb_rgbd);
  SSB_RETURN_IF(h->impl.init(max_keypoints, max_height, max_width, device_id));
  *out = h.release();
  return SSB_OK;
  SSB_API_END
}
void ssb_rgbd_destroy(ssb_rgbd* r) { delete r; }
int ssb_rgbd_process(ssb_rgbd* r, const float* xy, int n, const void* depth, int depth_type, int height,
                     int width, int row_stride, const double* camera, const double* dist, int n_dist,
                     double bf, double depth_factor, double max_depth, float* out_xy, double* out_stereo,
                     uint8_t* out_has_depth) {
  SSB_API_BEGIN
  SSB_CHECK(r != nullptr && camera != nullptr, SSB_ERR_INVALID, "null argument");
  SSB_CHECK(n_dist >= 0 && n_dist <= 14 && (dist != nullptr || n_dist == 0), SSB_ERR_INVALID, "bad dist");
  ssb::RgbdParams p;
  std::memset(&p, 0, sizeof(p));
  p.fx = camera[0], p.fy = camera[1], p.cx = camera[2], p.cy = camera[3];
  for (int i = 0; i < n_dist; ++i) {
    p.k[i] = dist[i];
    if (dist[i] != 0.0) p.has_dist = 1;   // cv::countNonZero(dist_coeffs_) > 0
  }
  p.bf = bf, p.depth_factor = depth_factor, p.max_depth = max_depth;
  return r->impl.process(xy, n, depth, depth_type, height, width, row_stride, p, out_xy, out_stereo, out_has_depth);
  SSB_API_END
}
int ssb_ep_create(const char* weights_path, int input_width, int input_height, int max_batch, int device_id,
                  ssb_eigenplaces** out) {
  SSB_API_BEGIN
  SSB_CHECK(out != nullptr, SSB_ERR_INVALID, "out is null");
  *out = nullptr;
  SSB_CUDA_CHECK(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  SSB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device_id));
  SSB_CHECK(prop.major == 10, SSB_ERR_NODEVICE, "device %d is sm_%d%d; this library needs sm_100", device_id,
            prop.major, prop.minor);
  std::unique_ptr<ssb_eigenplaces> h(new ssb_eigenplaces);
  SSB_RETURN_IF(h->impl.init(weights_path, input_width, input_height, max_batch <= 0 ? 1 : max_batch, device_id));
  *out = h.release();
  return SSB_OK;
  SSB_API_END
}
void ssb_ep_destroy(ssb_eigenplaces* ep) { delete ep; }
int ssb_ep_descriptor_dim(ssb_eigenplaces* ep) { return ep ? ssb::kEpDim : -1; }
int ssb_ep_compute(ssb_eigenplaces* ep, const uint8_t* const* images, int count, int height, int width,
                   int row_stride, int channels, float* descriptors) {
  SSB_API_BEGIN
  SSB_CHECK(ep != nullptr, SSB_ERR_INVALID, "ep is null");
  return ep->impl.compute(images, count, height, width, channels, row_stride, descriptors);
  SSB_API_END
}
int ssb_ep_add(ssb_eigenplaces* ep, uint64_t keyframe_id, const float* descriptor, int dim) {
  SSB_API_BEGIN
  SSB_CHECK(ep != nullptr, SSB_ERR_INVALID, "ep is null");
  return ep->impl.add(keyframe_id, descriptor, dim);
  SSB_API_END
}
int ssb_ep_query(ssb_eigenplaces* ep, const float* descriptor, int dim, uint64_t exclude_recent, int top_k,
                 float min_score, uint64_t* keyframe_ids, float* scores, int capacity, int* n_out) {
  SSB_API_BEGIN
  SSB_CHECK(ep != nullptr, SSB_ERR_INVALID, "ep is null");
  return ep->impl.query(descriptor, dim, exclude_recent, top_k, min_score, keyframe_ids, scores, capacity, n_out);
  SSB_API_END
}
int ssb_ep_index_size(ssb_eigenplaces* ep) { return ep ? ep->impl.index_size() : -1; }
int ssb_ep_debug_read(ssb_eigenplaces* ep, const char* what, void* dst, size_t bytes) {
  SSB_CHECK(ep != nullptr && dst != nullptr, SSB_ERR_INVALID, "null argument");
  return ep->impl.debug_read(what, dst, bytes);
}

ssb_superpoint* ssb_fe_superpoint(ssb_frontend* fe) { return fe ? &fe->impl.sp : nullptr; }
ssb_lightglue* ssb_fe_lightglue(ssb_frontend* fe) { return fe ? &fe->impl.lg : nullptr; }

}  // extern "C"
