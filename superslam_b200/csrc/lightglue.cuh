// LightGlue matcher runtime (B200-native).  Replaces the reference's TensorRT wrapper
// (/root/reference/src/LightGlue.cc, include/LightGlue.h) behind the C-ABI in
// include/superslam_b200.h.  The arithmetic is cvg/LightGlue's published model under the export
// settings of /root/reference/utils/convert_lightglue_to_onnx.py:61-90 (9 layers, 4 heads, d=256,
// no early exit / pruning, filter threshold 0.1); oracle/lightglue.py is the CPU restatement.
//
// Weights are immutable and shared between contexts (the reference shares one ICudaEngine between the
// tracking matcher and the loop-closure matcher: include/LightGlue.h:28-31, src/SuperSLAM.cc:129-133);
// each context owns its stream and workspace and is not re-entrant.
#pragma once

#include <memory>
#include <vector>

#include "common.cuh"
#include "weights.h"

namespace ssb {

constexpr int kLgLayers = 9;
constexpr int kLgHeads = 4;
constexpr int kLgDim = 256;
constexpr int kLgHeadDim = 64;

struct LgLinear {
  int n = 0, k = 0;
  __half* w = nullptr;   // [n][k] fp16, K-major rows (B operand)
  float* bias = nullptr; // [n]
  CUtensorMap tmB;
  CUtensorMap tmB128;    // same matrix, 128-row boxes
};

struct LgBlockFfn {
  LgLinear fc1, fc2;       // 512->512, 512->256
  float* ln_g = nullptr;   // LayerNorm(512) weight / bias
  float* ln_b = nullptr;
};

struct LgLayer {
  LgLinear qkv;      // 768 x 256, rows reordered to [q | k | v] x (head, dim); q rows pre-scaled by 1/8
  LgLinear out;      // out_proj
  LgBlockFfn sffn;
  LgLinear qkv_c;    // 512 x 256: [to_qk ; to_v]
  LgLinear to_out;
  LgBlockFfn cffn;
};

struct LgWeights {
  int device = 0;
  LgLayer layers[kLgLayers];
  LgLinear final_proj;
  float* match_w = nullptr;  // [256]
  float match_b = 0.f;
  float* wr = nullptr;       // posenc.Wr.weight [32][2]
  bool fold_out = true;      // out_proj / to_out folded into the right half of ffn.0 (LgWeights::load)
  std::vector<void*> owned;  // every device allocation, freed in the destructor
  ~LgWeights();
  int load(const char* path, int device);
};

class LightGlue {
 public:
  ~LightGlue();
  int init(std::shared_ptr<LgWeights> weights, int image_width, int image_height, int max_keypoints,
           int max_pairs);

  // Device path (LightGlue::match(kp, DeviceDescriptors, ...), src/LightGlue.cc:377-457):
  // host pixel keypoints, fp16 device descriptor rows; results host-visible on return.
  int match_device(const float* xy0, int n0, const void* desc0_dev, const float* xy1, int n1,
                   const void* desc1_dev, int32_t* matches0, float* mscores0);
  // Host path (src/LightGlue.cc:285-324): fp32 host descriptors are rounded to fp16 and uploaded.
  int match_host(const float* xy0, int n0, const float* desc0, const float* xy1, int n1,
                 const float* desc1, int32_t* matches0, float* mscores0);

  // Chained path: `pairs` pairs whose pixel keypoints / counts / descriptor rows are already on the
  // device (image 2p = left, 2p+1 = right); enqueues everything on `stream`, no host sync.
  int run(int pairs, const float* kp_xy_dev, int kp_stride, const int* kp_count_dev,
          void* const* desc_ptrs_dev, cudaStream_t stream);

  int max_keypoints() const { return kmax_; }
  int kp() const { return kp_; }
  cudaStream_t stream() const { return stream_; }
  const int32_t* matches_dev() const { return matches_; }   // [pairs][kp]
  const float* mscores_dev() const { return mscores_; }     // [pairs][kp]
  std::shared_ptr<LgWeights> weights() const { return w_; }
  int image_width() const { return img_w_; }
  int image_height() const { return img_h_; }
  int debug_read(const char* what, void* dst, size_t bytes);

 private:
  int alloc_workspace();
  int match_common(int n0, int n1, int32_t* matches0, float* mscores0);

  std::shared_ptr<LgWeights> w_;
  int device_ = 0, img_w_ = 0, img_h_ = 0, kmax_ = 0, kp_ = 0, pairs_ = 0;
  cudaStream_t stream_ = nullptr;

  // workspace (capacity: pairs_ pairs, kp_ = kmax_ rounded up to 128 rows per image)
  float* kp_xy_ = nullptr;      // [2P][kp][2] pixel keypoints (match_* entry points stage here)
  int* kp_count_ = nullptr;     // [2P]
  void** desc_ptrs_ = nullptr;  // [2P] device table
  __half* desc_stage_ = nullptr;  // [2][kp][256] staging for the host path
  float *cs_ = nullptr, *sn_ = nullptr;  // [2P][kp][32] cos / sin of the positional encoding
  float* x32_ = nullptr;        // residual stream, fp32 master, tile-transposed [2P][kp/128][256][128]
  __half* x16_ = nullptr;       // [2P][kp][256] fp16 copy (GEMM operand)
  __half *q_ = nullptr, *k_ = nullptr;  // [2P*4][kp][64]
  __half* v_ = nullptr;         // [2P*4][kp][64]
  float* s_ = nullptr;          // [P][kp][kp] assignment similarity sim (fp32)
  uint8_t* asg_part_ = nullptr; // per-tile partials of the two assignment sweeps (lightglue.cu assign_*_kernel)
  __half *ctx_ = nullptr, *msg_ = nullptr;  // [2P][kp][256]
  __half* h1_ = nullptr;        // [2P][kp][512]
  __half *mda_ = nullptr, *mdb_ = nullptr;  // [2P][kp][768] split-precision final projections
  float* lz_ = nullptr;         // [2P][kp] logsigmoid(matchability)
  float* lse_ = nullptr;        // [2P][kp] row / column log-sum-exp of sim
  float* max0_ = nullptr;       // [P][kp]
  int *arg0_ = nullptr, *arg1_ = nullptr;  // [P][kp]
  int32_t* matches_ = nullptr;
  float* mscores_ = nullptr;
  float* host_io_ = nullptr;    // pinned staging for match_* (xy in, matches/scores out)
  size_t host_io_bytes_ = 0;

  CUtensorMap tm_x16_, tm_msg_, tm_ctx_, tm_h1_, tm_q_a_, tm_q3_, tm_k3_, tm_v3_, tm_mda_a_,
      tm_mdb_b_;
  CUtensorMap ts_x16_, ts_msg_, ts_ctx_, ts_sim_, ts_h1_, ts_q_, ts_k_, ts_v_, ts_mda_, ts_mdb_;  // TMA-store maps
};

}  // namespace ssb
