// SuperPoint extractor runtime (B200-native).  Replaces the reference's TensorRT wrapper
// (/root/reference/src/SuperPoint.cc) + gather kernel (/root/reference/src/DescriptorGather.cu) +
// descriptor pool (/root/reference/include/DescriptorPool.h) behind the C-ABI in
// include/superslam_b200.h.  Everything from the u8 image to the sorted keypoint list and the
// fp16 descriptor rows stays on the device; keypoint counts are device-resident so LightGlue can
// be chained without a host round trip.
#pragma once

#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "weights.h"

namespace ssb {

constexpr int kDescDim = 256;
constexpr int kNmsRadius = 4;  // baked at export time in the reference (convert_superpoint_to_onnx.py:97)

// LIFO free list with the reference's semantics (DescriptorPool.h:25-44): acquire() == -1 when
// exhausted, release pushes back, in_use() counts live slots.  Adds a per-slot reference count so the
// C-ABI can hand out retain/release to the adapter's shared_ptr deleter.
class SlotPool {
 public:
  int init(int num_slots, int max_keypoints);
  ~SlotPool();
  int acquire();              // refcount := 1, or -1 if exhausted
  int retain(int slot);       // 0 ok
  int release(int slot);      // drops one reference; returns slot to the free list at zero
  int in_use();
  void* slot_ptr(int slot) const;
  int num_slots() const { return static_cast<int>(slots_.size()); }
  size_t slot_bytes() const { return slot_bytes_; }

 private:
  std::mutex mu_;
  std::vector<void*> slots_;
  std::vector<int> refs_;
  std::vector<int> free_;
  size_t slot_bytes_ = 0;
};

struct ConvLayer {
  int cin = 0, cout = 0, cout_pad = 0, taps = 0;
  __half* w = nullptr;   // [taps*taps][cout_pad][cin] fp16 (K-major rows for the B operand)
  float* bias = nullptr; // [cout_pad]
  CUtensorMap tmB;
  CUtensorMap tmB32;     // same matrix, 32-row boxes (one CTA of a pair)
  CUtensorMap tmB64;     // same matrix, 64-row boxes
  CUtensorMap tmB128;    // same matrix, 128-row boxes (conv_stream.cuh)
};

class SuperPoint {
 public:
  ~SuperPoint();
  int init(const char* weights_path, int max_keypoints, double threshold, int remove_borders,
           int num_slots, int device);

  // Run the whole extractor for `batch` same-size gray images already on the device
  // (u8, [batch][h][w] contiguous) on `stream`, no host synchronisation.  Results land in the
  // per-image device arrays below and in the descriptor slots `desc_out[i]` (fp16 [K][256]).
  int run(const uint8_t* images_dev, int batch, int h, int w, void* const* desc_out,
          cudaStream_t stream);

  // Reference-shaped synchronous call: host images in, host keypoints out, descriptors in pool slots.
  int extract(const uint8_t* const* images, int batch, int h, int w, int row_stride, int channels,
              float* const* xy, float* const* score, int* count, void** desc_dev, int* slot);

  int debug_read(const char* what, void* dst, size_t bytes);

  // Size the workspace for (batch, h, w) now.  shape_generation() changes whenever that frees and reallocates the
  // activation buffers (and re-encodes their tensor maps): anything that has baked the old pointers in - a captured
  // CUDA graph - must be dropped when it moves.
  int prepare(int batch, int h, int w) { return ensure_shape(batch, h, w); }
  unsigned long long shape_generation() const { return shape_gen_; }

  int max_keypoints() const { return max_kpts_; }
  SlotPool& pool() { return pool_; }
  cudaStream_t stream() const { return stream_; }
  // device-resident per-image outputs of the last run()
  const float* kp_xy() const { return kp_xy_; }        // [batch][K][2]
  const float* kp_score() const { return kp_score_; }  // [batch][K]
  const int* kp_count() const { return kp_count_; }    // [batch]
  int score_h() const { return hs_; }
  int score_w() const { return ws_; }
  int device() const { return device_; }
  uint8_t* staging_dev(size_t bytes);   // grow-only device staging for uploads
  uint8_t* staging_host(size_t bytes);  // grow-only pinned staging

 private:
  int ensure_shape(int batch, int h, int w);
  int load_layer(const WeightArchive& ar, const char* name, int cin, int cout, int taps, ConvLayer* L,
                 const char* name2 = nullptr);
  void free_shape();

  int device_ = 0;
  int max_kpts_ = 0, remove_borders_ = 0;
  double threshold_ = 0.0;
  cudaStream_t stream_ = nullptr;
  SlotPool pool_;

  // weights
  float* w1a_ = nullptr;  // conv1a fp32 [9][64]
  float* b1a_ = nullptr;
  ConvLayer l1b_, l2a_, l2b_, l3a_, l3b_, l4a_, l4b_, lpd_, lpb_, ldb_;

  // shape-dependent state
  int cap_batch_ = 0, h_ = 0, w_ = 0;
  unsigned long long shape_gen_ = 0;
  int h2_ = 0, w2_ = 0, h4_ = 0, w4_ = 0, hc_ = 0, wc_ = 0, hs_ = 0, ws_ = 0;
  uint8_t* img_ = nullptr;  // owned copy target for extract() [batch][h][w]
  __half *a1a_ = nullptr, *a1b_ = nullptr, *a2a_ = nullptr, *a2b_ = nullptr, *a3a_ = nullptr,
         *a3b_ = nullptr, *a4a_ = nullptr, *a4b_ = nullptr, *apd_ = nullptr, *grid_ = nullptr;
  float* scores_ = nullptr;              // [batch][hs][ws] softmax heat map (before NMS)
  unsigned long long* cand_ = nullptr;   // [batch][cand_cap] (score bits << 32 | h*ws + w)
  int* cand_count_ = nullptr;            // [batch]
  int cand_cap_ = 0;
  float* kp_xy_ = nullptr;
  float* kp_score_ = nullptr;
  int* kp_cell_ = nullptr;               // [batch][K] cell index (row*wc + col)
  int* kp_count_ = nullptr;
  CUtensorMap tm_a1a_, tm_a1b_, tm_a2a_, tm_a2b_, tm_a3a_, tm_a3b_, tm_a4a_, tm_a4b_, tm_apa_, tm_ada_;
  CUtensorMap tm_p1b_, tm_p2a_, tm_h2b_, tm_h3a_, tm_h3b_, tm_h4a_, tm_h4b_;  // (16+2) x (16+2) halo boxes
  CUtensorMap ts_a1b_, ts_a2a_, ts_a2b_, ts_a3a_, ts_a3b_, ts_a4a_, ts_a4b_, ts_apd_, ts_grid_;  // TMA-store maps

  void** desc_ptrs_dev_ = nullptr;  // [64] device table of per-image descriptor destinations
  uint8_t* stage_dev_ = nullptr;
  size_t stage_dev_bytes_ = 0;
  uint8_t* stage_host_ = nullptr;
  size_t stage_host_bytes_ = 0;
  float* out_host_ = nullptr;  // pinned: [cap_batch][K*3 + 1]
  size_t out_host_bytes_ = 0;
};

}  // namespace ssb
