// EigenPlaces global-descriptor runtime (B200-native): ResNet18 trunk + L2Norm / GeM / FC / L2Norm head and the
// cosine-similarity keyframe index.  Replaces the reference's TensorRT wrapper
// (/root/reference/src/EigenPlaces.cc, include/EigenPlaces.h) and CosineDescriptorIndex
// (/root/reference/src/PlaceRecognizer.cc:10-52) behind the C-ABI (ssb_ep_*).  Every convolution runs on
// the same tcgen05 implicit-GEMM core as SuperPoint's 1x1 layers and LightGlue's linears (umma_core.cuh):
// BatchNorm is folded into fp16 weights + fp32 bias at load time, the residual add and the ReLU live in the
// TMEM epilogue, stride-2 convolutions read their taps through TMA tensor maps with element strides 2.
#pragma once

#include <cstdint>
#include <vector>

#include "common.cuh"
#include "weights.h"

namespace ssb {

constexpr int kEpDim = 512;       // descriptor and final feature width
constexpr int kEpStemK = 192;     // 7*7*3 = 147 im2col columns padded to three 64-wide K chunks

struct EpConv {
  int cin = 0, cout = 0, taps = 0, stride = 1;
  int k_per_tap = 0;        // cin, or kEpStemK for the stem (a 1x1 product over the im2col matrix)
  __half* w = nullptr;      // [taps*taps][cout][k_per_tap] fp16, BatchNorm scale folded in
  float* bias = nullptr;    // [cout] folded BatchNorm shift
  CUtensorMap tmB;
};

struct EpStage {
  int C = 0, H = 0, W = 0;          // output geometry of the stage
  EpConv c1[2], c2[2], ds;          // two BasicBlocks; ds = 1x1 stride-2 projection of block 0 (stages 2-4)
  bool has_ds = false;
  __half *t = nullptr, *a = nullptr, *b = nullptr, *d = nullptr;   // conv1 out, block-0 out, block-1 out, ds out
  CUtensorMap ld_in, ld_t, ld_a;    // A-operand maps: stage input (strided for stages 2-4), t, a
  CUtensorMap st_t, st_a, st_b, st_d;
};

class EigenPlaces {
 public:
  ~EigenPlaces();
  // EigenPlaces(engine_file, input_width, input_height) + initialize() (include/EigenPlaces.h:26-30).
  int init(const char* weights_path, int in_w, int in_h, int max_batch, int device);
  // compute_global_descriptor for `count` same-size images (u8 gray or BGR); out: [count][512] fp32,
  // rows L2-normalised (src/EigenPlaces.cc:145-174).
  int compute(const uint8_t* const* images, int count, int h, int w, int channels, int row_stride, float* out);
  // CosineDescriptorIndex (src/PlaceRecognizer.cc:21-52); the database lives in device memory.
  int add(uint64_t keyframe_id, const float* desc, int dim);
  int query(const float* desc, int dim, uint64_t exclude_recent, int top_k, float min_score, uint64_t* ids,
            float* scores, int capacity, int* n_out);
  int index_size() const { return static_cast<int>(ids_.size()); }
  int debug_read(const char* what, void* dst, size_t bytes);
  int in_w() const { return in_w_; }
  int in_h() const { return in_h_; }

 private:
  int load_conv(const WeightArchive& ar, const std::string& conv, const std::string& bn, int cin, int cout,
                int taps, int stride, EpConv* L);
  int ensure_source(int h, int w, int channels);
  int run(int batch);
  int conv(const EpConv& L, const CUtensorMap& in, const CUtensorMap& out, const __half* residual, bool relu,
           int Ho, int Wo, int batch, const char* label);

  int device_ = 0, in_w_ = 0, in_h_ = 0, max_batch_ = 0;
  cudaStream_t stream_ = nullptr;
  std::vector<void*> owned_;   // every device allocation, freed by the destructor

  // weights
  EpConv stem_;
  EpStage stage_[4];
  float* fc_wt_ = nullptr;   // [512 in][512 out] fp32 (transposed: threads of a warp read consecutive outputs)
  float* fc_b_ = nullptr;
  float gem_p_ = 3.0f;

  // source-size dependent state (resize coefficient tables, staging)
  int src_h_ = 0, src_w_ = 0, src_c_ = 0, resize_mode_ = 0;
  uint8_t* src_dev_ = nullptr;
  uint8_t* src_host_ = nullptr;   // pinned
  size_t src_bytes_ = 0;
  int* tab_dev_ = nullptr;        // xofs[in_w] xa0[in_w] xa1[in_w] | yofs[in_h] yb0[in_h] yb1[in_h]
  // activations
  __half* x0_ = nullptr;     // [B][in_h][in_w][4] normalised RGB (+ one zero channel)
  __half* col_ = nullptr;    // [B][H2*W2][192] im2col of the 7x7 stride-2 stem
  __half* s0_ = nullptr;     // [B][H2][W2][64]
  __half* p0_ = nullptr;     // [B][H4][W4][64]
  CUtensorMap ld_col_, st_s0_;
  float* out_dev_ = nullptr;   // [B][512]
  float* out_host_ = nullptr;  // pinned

  // index
  std::vector<uint64_t> ids_;
  float* db_ = nullptr;        // [db_cap_][dim] fp32 rows, L2-normalised
  int db_cap_ = 0, db_dim_ = 0;
  float* q_dev_ = nullptr;
  float* sc_dev_ = nullptr;
  int sc_cap_ = 0;
  float* io_host_ = nullptr;   // pinned: query vector / new row, then scores
  size_t io_host_floats_ = 0;
};

}  // namespace ssb
