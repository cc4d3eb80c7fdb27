// SuperPoint on sm_100a: conv1a (SIMT, Cin=1) -> nine implicit-GEMM convolutions on tcgen05
// (umma_core.cuh) with bias/ReLU/2x2-maxpool, softmax + depth-to-space and channel-L2-norm fused in
// the TMEM epilogues -> 9x9 NMS + threshold + border + candidate compaction -> exact top-K with the
// reference's ordering -> nearest-cell descriptor gather with the reference's double normalisation.
//
// Parity notes (cite /root/reference):
//   graph            utils/convert_superpoint_to_onnx.py:51-90
//   preprocess       src/SuperPoint.cc:768-778   (u8 * (1/255) in fp32)
//   select           src/SuperPoint.cc:696-719   (strict > threshold as double, borders, sort order)
//   gather           src/DescriptorGather.cu:26-55 (fp32 256-wide tree sum, rsqrtf(+1e-12), fp16 out)
#include "superpoint.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>

#include "conv_stream.cuh"
#include "conv_pipe.cuh"
#include "umma_core.cuh"

namespace ssb {

// =================================================================================================
// SlotPool
// =================================================================================================
int SlotPool::init(int num_slots, int max_keypoints) {
  slot_bytes_ = static_cast<size_t>(max_keypoints) * kDescDim * sizeof(__half);
  // keep every slot 128-row padded so a TMA box never reads past the allocation
  size_t padded = static_cast<size_t>((max_keypoints + 127) / 128 * 128) * kDescDim * sizeof(__half);
  slots_.assign(num_slots, nullptr);
  refs_.assign(num_slots, 0);
  free_.clear();
  for (int i = num_slots - 1; i >= 0; --i) free_.push_back(i);
  for (int i = 0; i < num_slots; ++i) {
    SSB_CUDA_CHECK(cudaMalloc(&slots_[i], padded));
    SSB_CUDA_CHECK(cudaMemset(slots_[i], 0, padded));
  }
  return SSB_OK;
}
SlotPool::~SlotPool() {
  for (void* p : slots_)
    if (p) cudaFree(p);
}
int SlotPool::acquire() {
  std::lock_guard<std::mutex> g(mu_);
  if (free_.empty()) return -1;
  int s = free_.back();
  free_.pop_back();
  refs_[s] = 1;
  return s;
}
int SlotPool::retain(int slot) {
  std::lock_guard<std::mutex> g(mu_);
  if (slot < 0 || slot >= static_cast<int>(refs_.size()) || refs_[slot] <= 0) return SSB_ERR_INVALID;
  ++refs_[slot];
  return SSB_OK;
}
int SlotPool::release(int slot) {
  std::lock_guard<std::mutex> g(mu_);
  if (slot < 0 || slot >= static_cast<int>(refs_.size()) || refs_[slot] <= 0) return SSB_ERR_INVALID;
  if (--refs_[slot] == 0) free_.push_back(slot);
  return SSB_OK;
}
int SlotPool::in_use() {
  std::lock_guard<std::mutex> g(mu_);
  return static_cast<int>(slots_.size() - free_.size());
}
void* SlotPool::slot_ptr(int slot) const {
  if (slot < 0 || slot >= static_cast<int>(slots_.size())) return nullptr;
  return slots_[slot];
}

// =================================================================================================
// kernels
// =================================================================================================

// cv::cvtColor(BGR2GRAY) on u8 as OpenCV 4.x computes it: 15-bit fixed point,
// (B*3735 + G*19235 + R*9798 + 2^14) >> 15  (imgproc/src/color_rgb.simd.hpp: BY15, GY15, RY15, gray_shift = 15).
// oracle/imgproc.py::bgr_to_gray_u8 is pinned against cv2 on all 2^24 colours.  (The 14-bit coefficients
// 1868 / 9617 / 4899 of OpenCV 2.x/3.x, used here at first, differ by one grey level on 0.24 % of random pixels.)
__global__ void bgr_to_gray_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, size_t n) {
  pdl_wait();
  pdl_launch_dependents();
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint8_t* p = bgr + 3 * i;
  gray[i] = static_cast<uint8_t>((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
}

// ---- TMEM epilogues -------------------------------------------------------------------------------

// bias + ReLU (+ 2x2/2 max-pool, floor) -> fp16 NHWC, written with one TMA store per warp and 64-channel
// group (registers -> swizzled staging -> bulk store; out-of-image pixels are clipped by the TMA unit).
// Lane -> pixel mapping: 16x8 tiles of umma_core put two tile rows of 16 pixels in a warp
// (row_xor = 16); the halo kernels (conv_pipe, conv_stream) put four rows of 8 pixels in a warp
// (row_xor = 8).  Either way the 2x2 pool is two warp shuffles.
struct EpiConvRelu {
  const float* bias;
  CUtensorMap tm_out;  // 4-D (C, Wo, Ho, B); box = the warp's pixel block x 64 channels
  int pool;
  int row_xor;
  static constexpr bool kSplit = true;
  __device__ void operator()(EpiCtx& c, bool has_acc) const {
    const int px0 = __shfl_sync(0xffffffffu, c.px, 0), py0 = __shfl_sync(0xffffffffu, c.py, 0);
    bool writer = true;
    int srow = c.lane;
    if (pool) {
      writer = (c.lane & (1 | row_xor)) == 0;
      srow = row_xor == 16 ? ((c.lane >> 1) & 7) : (((c.lane >> 4) & 1) * 4 + ((c.lane >> 1) & 3));
    }
    for (int g0 = c.col_begin; g0 < c.col_end; g0 += 64) {
      stage_begin(c);
#pragma unroll 1
      for (int hc = 0; hc < 2; ++hc) {
        const int col = g0 + hc * 32;
        float v[32];
        tmem_ld_32x32(c.tmem_row + col, v);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(bias + c.n0 + col);
        uint32_t h[16];   // bias + ReLU in fp32, then packed fp16 pairs
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          const float x0 = fmaxf((has_acc ? v[4 * j + 0] : 0.f) + bb.x, 0.f);
          const float x1 = fmaxf((has_acc ? v[4 * j + 1] : 0.f) + bb.y, 0.f);
          const float x2 = fmaxf((has_acc ? v[4 * j + 2] : 0.f) + bb.z, 0.f);
          const float x3 = fmaxf((has_acc ? v[4 * j + 3] : 0.f) + bb.w, 0.f);
          h[2 * j] = pack_half2(x0, x1);
          h[2 * j + 1] = pack_half2(x2, x3);
        }
        if (pool) {
          // fp16 rounding is monotonic, so max over the rounded values == rounding of the fp32 max:
          // pool on packed half2 (half the shuffles, HMNMX2 instead of two FMNMX)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __half2 a = *reinterpret_cast<__half2*>(&h[j]);
            uint32_t o1 = __shfl_xor_sync(0xffffffffu, h[j], 1);
            a = __hmax2(a, *reinterpret_cast<__half2*>(&o1));
            uint32_t cur = *reinterpret_cast<uint32_t*>(&a);
            uint32_t o2 = __shfl_xor_sync(0xffffffffu, cur, row_xor);
            a = __hmax2(a, *reinterpret_cast<__half2*>(&o2));
            h[j] = *reinterpret_cast<uint32_t*>(&a);
          }
        }
        if (writer) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            stage_put(c, srow, hc * 4 + j, make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
        }
      }
      stage_fence(c);
      if (c.lane == 0) {
        tma_store_4d(&tm_out, c.stage_cur, c.n0 + g0, pool ? (px0 >> 1) : px0, pool ? (py0 >> 1) : py0, c.z);
        bulk_commit();
      }
    }
  }
};

// convPb epilogue: 65 logits -> softmax (fp32) -> drop the dustbin -> depth-to-space into the
// full-resolution heat map: channel c of cell (i,j) is pixel (8i + c/8, 8j + c%8)
// (convert_superpoint_to_onnx.py:77-81).
struct EpiScores {
  const float* bias;  // [80], entries >= 65 unused
  float* scores;      // [B][hs][ws]
  int Hc, Wc, hs, ws;
  static constexpr bool kSplit = false;  // the 65-way softmax needs the whole row in one thread
  // One thread = one cell: 65 logits -> probabilities of the 64 pixels.  exp through ex2.approx on log2(e)-scaled
  // logits and ONE reciprocal per cell (the first version: 65 expf + 64 IEEE divisions, ~1500 instructions per cell
  // and 0.2 ms per 128 images for a 0.16 GFLOP product); both stay within an ulp or two of the fp32 softmax, and all
  // index work downstream (NMS equality, threshold, top-K) is done on THIS map.
  __device__ void operator()(EpiCtx& c, bool) const {
    float v[80];
    tmem_ld_32x32(c.tmem_row, v);
    tmem_ld_32x32(c.tmem_row + 32, v + 32);
    tmem_ld_32x16(c.tmem_row + 64, v + 64);
    tmem_ld_wait();
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 65; ++j) {
      v[j] = (v[j] + __ldg(bias + j)) * kLog2e;
      m = fmaxf(m, v[j]);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 65; ++j) {
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[j] - m));
      v[j] = e;
      sum += e;
    }
    const float inv = 1.0f / sum;
    if (c.py < Hc && c.px < Wc) {
      float* base = scores + (static_cast<size_t>(c.z) * hs + c.py * 8) * ws + c.px * 8;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float4* dst = reinterpret_cast<float4*>(base + static_cast<size_t>(r) * ws);
        dst[0] = make_float4(v[8 * r + 0] * inv, v[8 * r + 1] * inv, v[8 * r + 2] * inv, v[8 * r + 3] * inv);
        dst[1] = make_float4(v[8 * r + 4] * inv, v[8 * r + 5] * inv, v[8 * r + 6] * inv, v[8 * r + 7] * inv);
      }
    }
  }
};

// convDb epilogue: F.normalize(p=2, dim=channels, eps=1e-12) in fp32, fp16 grid stored cell-major
// [B][Hc*Wc][256] so the gather reads one contiguous 512-byte row per keypoint.
struct EpiDescNorm {
  const float* bias;
  CUtensorMap tm_out;  // 4-D (256, Wc, Hc, B), box (64, 16, 2, 1)
  static constexpr bool kSplit = true;
  __device__ void operator()(EpiCtx& c, bool) const {
    const int px0 = __shfl_sync(0xffffffffu, c.px, 0), py0 = __shfl_sync(0xffffffffu, c.py, 0);
    float ss = 0.f;
    for (int col = c.col_begin; col < c.col_end; col += 32) {
      float v[32];
      tmem_ld_32x32(c.tmem_row + col, v);
      tmem_ld_wait();
      const float4* b4 = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = __ldg(b4 + j);
        const float x0 = v[4 * j] + bb.x, x1 = v[4 * j + 1] + bb.y, x2 = v[4 * j + 2] + bb.z, x3 = v[4 * j + 3] + bb.w;
        ss = fmaf(x0, x0, ss);      // same summation order as before (ascending columns)
        ss = fmaf(x1, x1, ss);
        ss = fmaf(x2, x2, ss);
        ss = fmaf(x3, x3, ss);
      }
    }
    ss = epi_pair_sum(c, ss);  // the two column halves of the row live in different warps
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    for (int g0 = c.col_begin; g0 < c.col_end; g0 += 64) {
      stage_begin(c);
#pragma unroll 1
      for (int hc = 0; hc < 2; ++hc) {
        const int col = g0 + hc * 32;
        float v[32];
        tmem_ld_32x32(c.tmem_row + col, v);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ba = __ldg(b4 + 2 * j), bb = __ldg(b4 + 2 * j + 1);
          float x[8];
          x[0] = (v[8 * j + 0] + ba.x) * inv, x[1] = (v[8 * j + 1] + ba.y) * inv;
          x[2] = (v[8 * j + 2] + ba.z) * inv, x[3] = (v[8 * j + 3] + ba.w) * inv;
          x[4] = (v[8 * j + 4] + bb.x) * inv, x[5] = (v[8 * j + 5] + bb.y) * inv;
          x[6] = (v[8 * j + 6] + bb.z) * inv, x[7] = (v[8 * j + 7] + bb.w) * inv;
          uint4 o;
          o.x = pack_half2(x[0], x[1]);
          o.y = pack_half2(x[2], x[3]);
          o.z = pack_half2(x[4], x[5]);
          o.w = pack_half2(x[6], x[7]);
          stage_put(c, c.lane, hc * 4 + j, o);
        }
      }
      stage_fence(c);
      if (c.lane == 0) {
        tma_store_4d(&tm_out, c.stage_cur, g0, px0, py0, c.z);
        bulk_commit();
      }
    }
  }
};

// ---- NMS + threshold + border + compaction ------------------------------------------------------
// 9x9 max filter (separable, -inf outside the map == max_pool2d padding), keep s == max
// (convert_superpoint_to_onnx.py:82-87), then the host scan of SuperPoint.cc:697-701 moved on device:
// interior only, (double)score > threshold.  Survivors are appended as 64-bit keys
// (score bits << 32 | linear index): scores are positive so integer order == (score, row, col) order.
// Tile: 128 x 32 outputs per CTA (256 threads), input region 136 x 40 in shared memory.  Both passes work on
// groups of four outputs whose nine-wide windows share six inputs: 12 inputs -> 4 maxima with nine 3-input
// maxima, read as three 16-byte shared-memory loads (rows) or 12 conflict-free loads (columns).  ~15
// instructions per pixel (the first version: 9 + 9 shared-memory loads and 16 maxima per pixel).
constexpr int kNmsTw = 128, kNmsTh = 32;
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
// out[k] = max(v[k .. k+8]), k = 0..3
__device__ __forceinline__ void window_max4(const float* v, float* out) {
  const float c = max3(max3(v[3], v[4], v[5]), max3(v[6], v[7], v[8]), v[3]);   // shared by all four windows
  const float l12 = fmaxf(v[1], v[2]), r910 = fmaxf(v[9], v[10]);
  out[0] = max3(c, v[0], l12);
  out[1] = max3(c, l12, v[9]);
  out[2] = max3(c, v[2], r910);
  out[3] = max3(c, r910, v[11]);
}
__global__ void __launch_bounds__(256)
nms_candidates_kernel(const float* __restrict__ scores, int hs, int ws, int rb, double thr,
                      unsigned long long* __restrict__ cand, int cand_cap, int* __restrict__ cand_count) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int R = kNmsRadius;
  constexpr int IW = kNmsTw + 2 * R, IH = kNmsTh + 2 * R;   // 136 x 40
  __shared__ __align__(16) float in[IH][IW];
  __shared__ __align__(16) float hm[IH][kNmsTw];
  const int z = blockIdx.z;
  const int x0 = blockIdx.x * kNmsTw, y0 = blockIdx.y * kNmsTh;
  const float* s = scores + static_cast<size_t>(z) * hs * ws;
  const int tid = threadIdx.x;
  for (int i = tid; i < IH * IW; i += 256) {
    const int r = i / IW, c = i - r * IW;
    const int y = y0 + r - R, x = x0 + c - R;
    in[r][c] = (y >= 0 && y < hs && x >= 0 && x < ws) ? __ldg(s + static_cast<size_t>(y) * ws + x) : -INFINITY;
  }
  __syncthreads();
  // horizontal: IH rows x 32 groups of four outputs
  for (int i = tid; i < IH * (kNmsTw / 4); i += 256) {
    const int r = i >> 5, g = (i & 31) * 4;
    float v[12], o[4];
    const float4 a = *reinterpret_cast<const float4*>(&in[r][g]);
    const float4 b = *reinterpret_cast<const float4*>(&in[r][g + 4]);
    const float4 c = *reinterpret_cast<const float4*>(&in[r][g + 8]);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    v[8] = c.x, v[9] = c.y, v[10] = c.z, v[11] = c.w;
    window_max4(v, o);
    *reinterpret_cast<float4*>(&hm[r][g]) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  // vertical: thread = column x (lane = consecutive columns), 16 rows in four groups of four
  const int tx = tid & (kNmsTw - 1), ty0 = (tid >> 7) * 16;
  const int x = x0 + tx;
  const int lane = tid & 31;
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    const int r0 = ty0 + g * 4;
    float v[12], o[4];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = hm[r0 + k][tx];
    window_max4(v, o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int y = y0 + r0 + k;
      const float c = in[r0 + k + R][tx + R];
      const bool keep = y >= rb && y < hs - rb && x >= rb && x < ws - rb && c == o[k] &&
                        static_cast<double>(c) > thr;
      const unsigned ball = __ballot_sync(0xffffffffu, keep);
      if (ball != 0) {
        int base = 0;
        const int leader = __ffs(ball) - 1;
        if (lane == leader) base = atomicAdd(&cand_count[z], __popc(ball));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) {
          const int pos = base + __popc(ball & ((1u << lane) - 1u));
          if (pos < cand_cap) {
            cand[static_cast<size_t>(z) * cand_cap + pos] =
                (static_cast<unsigned long long>(__float_as_uint(c)) << 32) | static_cast<unsigned>(y * ws + x);
          }
        }
      }
    }
  }
}

// ---- exact top-K with the reference's order -----------------------------------------------------
// One CTA per image.  If there are more candidates than K, an 8-pass MSB radix select finds the K-th
// largest key (keys are unique), the >= set is compacted into shared memory, and a bitonic sort
// produces std::sort(..., std::greater<>()) order: score desc, row desc, col desc
// (SuperPoint.cc:703-704).  Emits keypoints exactly like SuperPoint.cc:708-718.
__global__ void __launch_bounds__(1024)
select_topk_kernel(const unsigned long long* __restrict__ cand, int cand_cap,
                   const int* __restrict__ cand_count, int K, int sort_cap, int ws, int hc, int wc,
                   float scale_x, float scale_y, float* __restrict__ kp_xy, float* __restrict__ kp_score,
                   int* __restrict__ kp_cell, int* __restrict__ kp_count) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ unsigned long long skeys[];  // [sort_cap]
  __shared__ unsigned hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining, s_fill;
  const int z = blockIdx.x;
  const int tid = threadIdx.x;
  const unsigned long long* keys = cand + static_cast<size_t>(z) * cand_cap;
  const int nc = min(cand_count[z], cand_cap);
  const int n = min(nc, K);

  if (nc > K) {
    if (tid == 0) {
      s_prefix = 0ull;
      s_remaining = K;
    }
    unsigned long long mask = 0ull;
    for (int byte = 7; byte >= 0; --byte) {
      for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int i = tid; i < nc; i += blockDim.x) {
        const unsigned long long k = keys[i];
        if ((k & mask) == prefix) atomicAdd(&hist[(k >> (8 * byte)) & 0xffu], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        int rem = s_remaining;
        int b = 255;
        for (; b > 0; --b) {
          const int hcount = static_cast<int>(hist[b]);
          if (hcount >= rem) break;
          rem -= hcount;
        }
        s_remaining = rem;
        s_prefix = prefix | (static_cast<unsigned long long>(b) << (8 * byte));
      }
      mask |= 0xffull << (8 * byte);
      __syncthreads();
    }
    const unsigned long long kth = s_prefix;
    if (tid == 0) s_fill = 0;
    __syncthreads();
    for (int i = tid; i < nc; i += blockDim.x) {
      const unsigned long long k = keys[i];
      if (k >= kth) {
        const int pos = atomicAdd(&s_fill, 1);
        if (pos < sort_cap) skeys[pos] = k;
      }
    }
    __syncthreads();
  } else {
    for (int i = tid; i < nc; i += blockDim.x) skeys[i] = keys[i];
  }
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = n + tid; i < P; i += blockDim.x) skeys[i] = 0ull;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = skeys[i], b = skeys[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            skeys[i] = b;
            skeys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  float* xy = kp_xy + static_cast<size_t>(z) * K * 2;
  float* sc = kp_score + static_cast<size_t>(z) * K;
  int* cell = kp_cell + static_cast<size_t>(z) * K;
  for (int i = tid; i < K; i += blockDim.x) {
    if (i < n) {
      const unsigned long long k = skeys[i];
      const unsigned idx = static_cast<unsigned>(k & 0xffffffffull);
      const int h = static_cast<int>(idx) / ws, w = static_cast<int>(idx) % ws;
      xy[2 * i + 0] = static_cast<float>(w) * scale_x;
      xy[2 * i + 1] = static_cast<float>(h) * scale_y;
      sc[i] = __uint_as_float(static_cast<unsigned>(k >> 32));
      cell[i] = min(h / 8, hc - 1) * wc + min(w / 8, wc - 1);
    } else {
      xy[2 * i + 0] = 0.f;
      xy[2 * i + 1] = 0.f;
      sc[i] = 0.f;
      cell[i] = 0;
    }
  }
  if (tid == 0) kp_count[z] = n;
}

// ---- descriptor gather + second normalisation -----------------------------------------------------
// One warp per keypoint: lane l loads channels 8l..8l+7 (one 16-byte load, 512 B per row coalesced).
// The fp32 sum of squares reproduces the reference's 256-thread shared-memory tree bit for bit:
// strides 128,64,32,16,8 pair channel c with c+stride == lanes l and l ^ (stride/8); strides 4,2,1
// pair registers inside the lane (DescriptorGather.cu:41-46).
__global__ void __launch_bounds__(256)
gather_normalize_kernel(const __half* __restrict__ grid, int cells, const int* __restrict__ kp_cell,
                        const int* __restrict__ kp_count, int K, void* const* __restrict__ desc_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int z = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kp = blockIdx.x * 8 + warp;
  if (kp >= K) return;
  __half* out = static_cast<__half*>(desc_out[z]);
  if (out == nullptr) return;
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(kp) * kDescDim) + lane;
  if (kp >= kp_count[z]) {
    *dst = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const int cell = kp_cell[static_cast<size_t>(z) * K + kp];
  const uint4 raw = *(reinterpret_cast<const uint4*>(
                          grid + (static_cast<size_t>(z) * cells + cell) * kDescDim) + lane);
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
  float v[8], p[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h2[j]);
    v[2 * j] = f.x;
    v[2 * j + 1] = f.y;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) p[j] = v[j] * v[j];
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = p[j] + __shfl_xor_sync(0xffffffffu, p[j], m);
  }
  p[0] += p[4];
  p[1] += p[5];
  p[2] += p[6];
  p[3] += p[7];
  p[0] += p[2];
  p[1] += p[3];
  p[0] += p[1];
  const float inv = rsqrtf(p[0] + 1e-12f);
  uint4 o;
  o.x = pack_half2(v[0] * inv, v[1] * inv);
  o.y = pack_half2(v[2] * inv, v[3] * inv);
  o.z = pack_half2(v[4] * inv, v[5] * inv);
  o.w = pack_half2(v[6] * inv, v[7] * inv);
  *dst = o;
}

// =================================================================================================
// host side
// =================================================================================================
SuperPoint::~SuperPoint() {
  cudaSetDevice(device_);
  free_shape();
  auto freel = [](ConvLayer& L) {
    if (L.w) cudaFree(L.w);
    if (L.bias) cudaFree(L.bias);
  };
  freel(l1b_), freel(l2a_), freel(l2b_), freel(l3a_), freel(l3b_), freel(l4a_), freel(l4b_);
  freel(lpd_), freel(lpb_), freel(ldb_);
  if (desc_ptrs_dev_) cudaFree(desc_ptrs_dev_);
  if (w1a_) cudaFree(w1a_);
  if (b1a_) cudaFree(b1a_);
  if (stage_dev_) cudaFree(stage_dev_);
  if (stage_host_) cudaFreeHost(stage_host_);
  if (out_host_) cudaFreeHost(out_host_);
  if (stream_) cudaStreamDestroy(stream_);
}

// Lay a conv weight (cout, cin, t, t) out as the K-major B operand: row = tap*cout_pad + co, cin
// contiguous.  name2 (optional) is a second conv over the same input, stacked along cout.
int SuperPoint::load_layer(const WeightArchive& ar, const char* name, int cin, int cout, int taps,
                           ConvLayer* L, const char* name2) {
  const std::string n1(name);
  const HostTensor* w1 = ar.get(n1 + ".weight", {cout, cin, taps, taps});
  const HostTensor* b1 = ar.get(n1 + ".bias", {cout});
  if (!w1 || !b1) return SSB_ERR_IO;
  const HostTensor *w2 = nullptr, *b2 = nullptr;
  if (name2) {
    const std::string n2(name2);
    w2 = ar.get(n2 + ".weight", {cout, cin, taps, taps});
    b2 = ar.get(n2 + ".bias", {cout});
    if (!w2 || !b2) return SSB_ERR_IO;
  }
  const int total = name2 ? 2 * cout : cout;
  const int pad = (total + 15) / 16 * 16;
  L->cin = cin;
  L->cout = total;
  L->cout_pad = pad;
  L->taps = taps;
  const int T = taps * taps;
  std::vector<__half> hw(static_cast<size_t>(T) * pad * cin, __float2half(0.f));
  std::vector<float> hb(pad, 0.f);
  for (int src = 0; src < (name2 ? 2 : 1); ++src) {
    const HostTensor* w = src ? w2 : w1;
    const HostTensor* b = src ? b2 : b1;
    for (int co = 0; co < cout; ++co) {
      hb[src * cout + co] = b->data[co];
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < T; ++t)
          hw[(static_cast<size_t>(t) * pad + src * cout + co) * cin + ci] =
              __float2half(w->data[(static_cast<size_t>(co) * cin + ci) * T + t]);
    }
  }
  SSB_CUDA_CHECK(cudaMalloc(&L->w, hw.size() * sizeof(__half)));
  SSB_CUDA_CHECK(cudaMalloc(&L->bias, hb.size() * sizeof(float)));
  SSB_CUDA_CHECK(cudaMemcpy(L->w, hw.data(), hw.size() * sizeof(__half), cudaMemcpyHostToDevice));
  SSB_CUDA_CHECK(cudaMemcpy(L->bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
  const int n_part = pad > 256 ? 256 : pad;
  uint64_t dims[3] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(T) * pad, 1};
  uint64_t strides[2] = {static_cast<uint64_t>(cin) * 2, static_cast<uint64_t>(T) * pad * cin * 2};
  uint32_t box[3] = {64, static_cast<uint32_t>(n_part), 1};
  SSB_RETURN_IF(encode_tmap_f16(&L->tmB, L->w, 3, dims, strides, box));
  uint32_t box64[3] = {64, 64, 1};   // one 64-output-channel slice of one tap (conv_pipe.cuh)
  SSB_RETURN_IF(encode_tmap_f16(&L->tmB64, L->w, 3, dims, strides, box64));
  uint32_t box32[3] = {64, 32, 1};   // half a slice: one CTA of a pair (conv_pipe.cuh, kPair)
  SSB_RETURN_IF(encode_tmap_f16(&L->tmB32, L->w, 3, dims, strides, box32));
  uint32_t box128[3] = {64, static_cast<uint32_t>(pad >= 128 ? 128 : pad), 1};   // conv_stream.cuh
  return encode_tmap_f16(&L->tmB128, L->w, 3, dims, strides, box128);
}

int SuperPoint::init(const char* weights_path, int max_keypoints, double threshold, int remove_borders,
                     int num_slots, int device) {
  SSB_CHECK(max_keypoints > 0 && max_keypoints <= 8192, SSB_ERR_INVALID,
            "max_keypoints %d out of range (1..8192)", max_keypoints);
  SSB_CHECK(remove_borders >= 0, SSB_ERR_INVALID, "remove_borders must be >= 0");
  device_ = device;
  max_kpts_ = max_keypoints;
  threshold_ = threshold;
  remove_borders_ = remove_borders;
  SSB_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  SSB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  SSB_CHECK(prop.major == 10, SSB_ERR_NODEVICE, "device %d is sm_%d%d; this library needs sm_100",
            device, prop.major, prop.minor);
  SSB_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  WeightArchive ar;
  SSB_RETURN_IF(load_archive(weights_path, &ar));
  const HostTensor* w1 = ar.get("conv1a.weight", {64, 1, 3, 3});
  const HostTensor* b1 = ar.get("conv1a.bias", {64});
  if (!w1 || !b1) return SSB_ERR_IO;
  std::vector<float> w1t(576);
  for (int co = 0; co < 64; ++co)
    for (int t = 0; t < 9; ++t) w1t[t * 64 + co] = w1->data[co * 9 + t];
  SSB_CUDA_CHECK(cudaMalloc(&w1a_, 576 * sizeof(float)));
  SSB_CUDA_CHECK(cudaMalloc(&b1a_, 64 * sizeof(float)));
  SSB_CUDA_CHECK(cudaMemcpy(w1a_, w1t.data(), 576 * sizeof(float), cudaMemcpyHostToDevice));
  SSB_CUDA_CHECK(cudaMemcpy(b1a_, b1->data.data(), 64 * sizeof(float), cudaMemcpyHostToDevice));
  SSB_RETURN_IF(load_layer(ar, "conv1b", 64, 64, 3, &l1b_));
  SSB_RETURN_IF(load_layer(ar, "conv2a", 64, 64, 3, &l2a_));
  SSB_RETURN_IF(load_layer(ar, "conv2b", 64, 64, 3, &l2b_));
  SSB_RETURN_IF(load_layer(ar, "conv3a", 64, 128, 3, &l3a_));
  SSB_RETURN_IF(load_layer(ar, "conv3b", 128, 128, 3, &l3b_));
  SSB_RETURN_IF(load_layer(ar, "conv4a", 128, 128, 3, &l4a_));
  SSB_RETURN_IF(load_layer(ar, "conv4b", 128, 128, 3, &l4b_));
  SSB_RETURN_IF(load_layer(ar, "convPa", 128, 256, 3, &lpd_, "convDa"));
  SSB_RETURN_IF(load_layer(ar, "convPb", 256, 65, 1, &lpb_));
  SSB_RETURN_IF(load_layer(ar, "convDb", 256, 256, 1, &ldb_));
  SSB_RETURN_IF(pool_.init(num_slots > 0 ? num_slots : 8, max_keypoints));
  SSB_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&desc_ptrs_dev_), 128 * sizeof(void*)));
  return SSB_OK;
}

void SuperPoint::free_shape() {
  void* bufs[] = {img_, a1a_, a1b_, a2a_, a2b_, a3a_, a3b_, a4a_, a4b_, apd_, grid_, scores_,
                  cand_, cand_count_, kp_xy_, kp_score_, kp_cell_, kp_count_};
  for (void* p : bufs)
    if (p) cudaFree(p);
  img_ = nullptr;
  a1a_ = a1b_ = a2a_ = a2b_ = a3a_ = a3b_ = a4a_ = a4b_ = apd_ = grid_ = nullptr;
  scores_ = nullptr;
  cand_ = nullptr;
  cand_count_ = kp_cell_ = kp_count_ = nullptr;
  kp_xy_ = kp_score_ = nullptr;
  cap_batch_ = 0;
}

static int make_act_tmap(CUtensorMap* tm, const __half* base, int C, int Cpitch, int W, int H, int B,
                         int box_w = 16, int box_h = 8);
static int make_halo_tmap(CUtensorMap* tm, const __half* base, int C, int W, int H, int B, int subtiles) {
  return make_act_tmap(tm, base, C, C, W, H, B, 8 * subtiles + 2, 18);
}
static int make_act_tmap(CUtensorMap* tm, const __half* base, int C, int Cpitch, int W, int H, int B, int box_w,
                         int box_h) {
  uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                      static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(Cpitch) * 2, static_cast<uint64_t>(W) * Cpitch * 2,
                         static_cast<uint64_t>(H) * W * Cpitch * 2};
  uint32_t box[4] = {64, static_cast<uint32_t>(box_w), static_cast<uint32_t>(box_h), 1};
  return encode_tmap_f16(tm, base, 4, dims, strides, box);
}

int SuperPoint::ensure_shape(int batch, int h, int w) {
  if (batch <= cap_batch_ && h == h_ && w == w_) return SSB_OK;
  SSB_CHECK(h >= 16 && w >= 16 && h <= 4096 && w <= 4096, SSB_ERR_INVALID,
            "image size %dx%d unsupported (16..4096)", w, h);
  SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
  const int nb = std::max(batch, (h == h_ && w == w_) ? cap_batch_ : 0);
  free_shape();
  ++shape_gen_;
  h_ = h, w_ = w;
  h2_ = h / 2, w2_ = w / 2, h4_ = h2_ / 2, w4_ = w2_ / 2, hc_ = h4_ / 2, wc_ = w4_ / 2;
  hs_ = hc_ * 8, ws_ = wc_ * 8;
  const size_t B = nb;
  auto alloc = [&](void** p, size_t bytes) -> int {
    // +64 KiB slack: TMA boxes and vector stores never straddle the end of an allocation
    SSB_CUDA_CHECK(cudaMalloc(p, bytes + 65536));
    SSB_CUDA_CHECK(cudaMemsetAsync(*p, 0, bytes + 65536, stream_));
    return SSB_OK;
  };
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&img_), B * h * w));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&a1b_), B * h2_ * w2_ * 64 * 2));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&a2a_), B * h2_ * w2_ * 64 * 2));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&a2b_), B * h4_ * w4_ * 64 * 2));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&a3a_), B * h4_ * w4_ * 128 * 2));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&a3b_), B * hc_ * wc_ * 128 * 2));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&a4a_), B * hc_ * wc_ * 128 * 2));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&a4b_), B * hc_ * wc_ * 128 * 2));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&apd_), B * hc_ * wc_ * 512 * 2));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&grid_), B * hc_ * wc_ * 256 * 2));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&scores_), B * hs_ * ws_ * 4));
  // a 9x9 NMS leaves at most one survivor per 5x5 block, plus plateau ties: generous bound
  cand_cap_ = std::max(4 * max_kpts_, (hs_ * ws_) / 16 + 1024);
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&cand_), B * cand_cap_ * 8));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&cand_count_), B * 4));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&kp_xy_), B * max_kpts_ * 2 * 4));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&kp_score_), B * max_kpts_ * 4));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&kp_cell_), B * max_kpts_ * 4));
  SSB_RETURN_IF(alloc(reinterpret_cast<void**>(&kp_count_), B * 4));
  SSB_RETURN_IF(make_act_tmap(&tm_a1b_, a1b_, 64, 64, w2_, h2_, nb));
  SSB_RETURN_IF(make_act_tmap(&tm_a2a_, a2a_, 64, 64, w2_, h2_, nb));
  SSB_RETURN_IF(make_act_tmap(&tm_a2b_, a2b_, 64, 64, w4_, h4_, nb));
  SSB_RETURN_IF(make_act_tmap(&tm_a3a_, a3a_, 128, 128, w4_, h4_, nb));
  SSB_RETURN_IF(make_act_tmap(&tm_a3b_, a3b_, 128, 128, wc_, hc_, nb));
  SSB_RETURN_IF(make_act_tmap(&tm_a4a_, a4a_, 128, 128, wc_, hc_, nb));
  SSB_RETURN_IF(make_act_tmap(&tm_a4b_, a4b_, 128, 128, wc_, hc_, nb));
  SSB_RETURN_IF(make_halo_tmap(&tm_p1b_, a1b_, 64, w2_, h2_, nb, 2));   // 18 x 18 halo boxes
  SSB_RETURN_IF(make_halo_tmap(&tm_p2a_, a2a_, 64, w2_, h2_, nb, 2));
  SSB_RETURN_IF(make_halo_tmap(&tm_h2b_, a2b_, 64, w4_, h4_, nb, 2));
  SSB_RETURN_IF(make_halo_tmap(&tm_h3a_, a3a_, 128, w4_, h4_, nb, 2));
  SSB_RETURN_IF(make_halo_tmap(&tm_h3b_, a3b_, 128, wc_, hc_, nb, 2));
  SSB_RETURN_IF(make_halo_tmap(&tm_h4a_, a4a_, 128, wc_, hc_, nb, 2));
  SSB_RETURN_IF(make_halo_tmap(&tm_h4b_, a4b_, 128, wc_, hc_, nb, 2));
  // store maps: one warp's pixel block x 64 channels (halo kernels: 4 rows x 8 px, pooled 2 x 4;
  // 16x8-tile kernels: 2 rows x 16 px)
  SSB_RETURN_IF(make_act_tmap(&ts_a1b_, a1b_, 64, 64, w2_, h2_, nb, 4, 2));
  SSB_RETURN_IF(make_act_tmap(&ts_a2a_, a2a_, 64, 64, w2_, h2_, nb, 8, 4));
  SSB_RETURN_IF(make_act_tmap(&ts_a2b_, a2b_, 64, 64, w4_, h4_, nb, 4, 2));
  SSB_RETURN_IF(make_act_tmap(&ts_a3a_, a3a_, 128, 128, w4_, h4_, nb, 8, 4));
  SSB_RETURN_IF(make_act_tmap(&ts_a3b_, a3b_, 128, 128, wc_, hc_, nb, 4, 2));
  SSB_RETURN_IF(make_act_tmap(&ts_a4a_, a4a_, 128, 128, wc_, hc_, nb, 8, 4));
  SSB_RETURN_IF(make_act_tmap(&ts_a4b_, a4b_, 128, 128, wc_, hc_, nb, 8, 4));
  SSB_RETURN_IF(make_act_tmap(&ts_apd_, apd_, 512, 512, wc_, hc_, nb, 8, 4));
  SSB_RETURN_IF(make_act_tmap(&ts_grid_, grid_, 256, 256, wc_, hc_, nb, 16, 2));
  SSB_RETURN_IF(make_act_tmap(&tm_apa_, apd_, 256, 512, wc_, hc_, nb));
  SSB_RETURN_IF(make_act_tmap(&tm_ada_, apd_ + 256, 256, 512, wc_, hc_, nb));
  const size_t need = B * (static_cast<size_t>(max_kpts_) * 3 + 1) * sizeof(float);
  if (need > out_host_bytes_) {
    if (out_host_) cudaFreeHost(out_host_);
    SSB_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&out_host_), need));
    out_host_bytes_ = need;
  }
  cap_batch_ = nb;
  SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
  return SSB_OK;
}

uint8_t* SuperPoint::staging_dev(size_t bytes) {
  if (bytes > stage_dev_bytes_) {
    if (stage_dev_) cudaFree(stage_dev_);
    stage_dev_ = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&stage_dev_), bytes) != cudaSuccess) return nullptr;
    stage_dev_bytes_ = bytes;
  }
  return stage_dev_;
}
uint8_t* SuperPoint::staging_host(size_t bytes) {
  if (bytes > stage_host_bytes_) {
    if (stage_host_) cudaFreeHost(stage_host_);
    stage_host_ = nullptr;
    if (cudaMallocHost(reinterpret_cast<void**>(&stage_host_), bytes) != cudaSuccess) return nullptr;
    stage_host_bytes_ = bytes;
  }
  return stage_host_;
}

static CoreParams conv_params(int taps, int cin, int cout_rows, int block_n, int W) {
  CoreParams p;
  std::memset(&p, 0, sizeof(p));
  p.taps_h = p.taps_w = taps;
  p.pad = taps / 2;
  p.kc0 = cin / 64;
  p.kc1 = 0;
  p.b_tap_rows = cout_rows;
  p.tile_w = 16;
  p.tile_h = 8;
  p.tiles_w = (W + 15) / 16;
  p.block_n = block_n;
  p.a_z_mul = 1;
  return p;
}

int SuperPoint::run(const uint8_t* images_dev, int batch, int h, int w, void* const* desc_out,
                    cudaStream_t stream) {
  SSB_CHECK(batch >= 1, SSB_ERR_INVALID, "batch must be >= 1");
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  SSB_RETURN_IF(ensure_shape(batch, h, w));
  const int B = batch;
  // candidate counters of the NMS: reset here, in front of the first kernel, so that no memset node sits between two
  // kernels of the chain (every kernel-to-kernel edge stays a programmatic one, common.cuh pdl_wait)
  SSB_CUDA_CHECK(cudaMemsetAsync(cand_count_, 0, B * sizeof(int), stream));
  // Cin = 128 layers: persistent halo-reuse kernel with streamed weights, 128 output channels per CTA
  auto sconv = [&](const char* label, const CUtensorMap& tmH, const CUtensorMap& tmS, const ConvLayer& L, int H,
                   int W, int pool) -> int {
    StreamParams p;
    std::memset(&p, 0, sizeof(p));
    p.n_slices = L.cout / 128;
    p.cout_rows = L.cout_pad;
    p.label = label;
    EpiConvRelu e{L.bias, tmS, pool, 8};
    return launch_conv_stream(tmH, L.tmB128, p, e, W, H, B, stream);
  };
  // Cin = 64 layers: persistent warp-specialised kernel, weights resident, halo + TMEM double-buffered.
  // conv1a (Cin = 1) is evaluated by the halo-producer warps of conv1b: its activation never touches HBM.
  auto pconv = [&](const char* label, const CUtensorMap& tmH, const CUtensorMap& tmS, const ConvLayer& L, int H,
                   int W, int pool, bool fuse1a) -> int {
    PipeParams p;
    std::memset(&p, 0, sizeof(p));
    p.n_slices = L.cout / 64;
    p.cout_rows = L.cout_pad;
    p.label = label;
    p.img = images_dev;
    p.w1a = w1a_;
    p.b1a = b1a_;
    p.img_h = h;
    p.img_w = w;
    EpiConvRelu e{L.bias, tmS, pool, 8};
    // Cout = 64: CTA pairs (cta_group::2), each CTA reads half of the weights per MMA (conv1a+1b 2.88 -> 2.73 ms,
    // conv2a / conv2b 0.64 -> 0.57 ms per 128 images).  SSB_SP_PAIR=0 selects the one-CTA kernels for A/B measurements.
    static const bool pair = [] { const char* e = std::getenv("SSB_SP_PAIR"); return e == nullptr || std::atoi(e) != 0; }();
    if (pair && p.n_slices == 2 && !fuse1a)   // conv3a: 128 output channels, 64 per CTA of a pair
      return launch_conv_pipe<EpiConvRelu, false, true, 128>(tmH, L.tmB64, p, e, W, H, B, stream);
    if (pair && p.n_slices == 1) {
      if (fuse1a) return launch_conv_pipe<EpiConvRelu, true, true>(L.tmB32, L.tmB32, p, e, W, H, B, stream);
      return launch_conv_pipe<EpiConvRelu, false, true>(tmH, L.tmB32, p, e, W, H, B, stream);
    }
    if (fuse1a) return launch_conv_pipe<EpiConvRelu, true>(L.tmB64, L.tmB64, p, e, W, H, B, stream);
    return launch_conv_pipe<EpiConvRelu, false>(tmH, L.tmB64, p, e, W, H, B, stream);
  };
  SSB_RETURN_IF(pconv("sp.conv1ab", tm_p1b_, ts_a1b_, l1b_, h, w, 1, true));
  SSB_RETURN_IF(pconv("sp.conv2a", tm_p1b_, ts_a2a_, l2a_, h2_, w2_, 0, false));
  SSB_RETURN_IF(pconv("sp.conv2b", tm_p2a_, ts_a2b_, l2b_, h2_, w2_, 1, false));
  SSB_RETURN_IF(pconv("sp.conv3a", tm_h2b_, ts_a3a_, l3a_, h4_, w4_, 0, false));
  SSB_RETURN_IF(sconv("sp.conv3b", tm_h3a_, ts_a3b_, l3b_, h4_, w4_, 1));
  SSB_RETURN_IF(sconv("sp.conv4a", tm_h3b_, ts_a4a_, l4a_, hc_, wc_, 0));
  SSB_RETURN_IF(sconv("sp.conv4b", tm_h4a_, ts_a4b_, l4b_, hc_, wc_, 0));
  SSB_RETURN_IF(sconv("sp.convPaDa", tm_h4b_, ts_apd_, lpd_, hc_, wc_, 0));
  {
    CoreParams p = conv_params(1, 256, lpb_.cout_pad, lpb_.cout_pad, wc_);
    p.label = "sp.convPb";
    EpiScores e{lpb_.bias, scores_, hc_, wc_, hs_, ws_};
    dim3 g(p.tiles_w * ((hc_ + 7) / 8), 1, B);
    SSB_RETURN_IF(launch_core(tm_apa_, tm_apa_, lpb_.tmB, p, e, g, stream));
  }
  {
    CoreParams p = conv_params(1, 256, ldb_.cout_pad, 256, wc_);
    p.label = "sp.convDb";
    EpiDescNorm e{ldb_.bias, ts_grid_};
    dim3 g(p.tiles_w * ((hc_ + 7) / 8), 1, B);
    SSB_RETURN_IF(launch_core(tm_ada_, tm_ada_, ldb_.tmB, p, e, g, stream));
  }
  {
    dim3 g((ws_ + kNmsTw - 1) / kNmsTw, (hs_ + kNmsTh - 1) / kNmsTh, B);
    SSB_CUDA_CHECK(launch_kernel(nms_candidates_kernel, dim3(g), dim3(256), 0, stream, 1, scores_, hs_, ws_, remove_borders_, threshold_,
                                                          cand_, cand_cap_, cand_count_));
    count_launch();
    prof_mark(stream, "sp.nms");
  }
  {
    int sort_cap = 1;
    while (sort_cap < max_kpts_) sort_cap <<= 1;
    const float sx = static_cast<float>(w) / ws_;  // SuperPoint.cc:708-709
    const float sy = static_cast<float>(h) / hs_;
    const size_t smem = static_cast<size_t>(sort_cap) * 8;
    auto configure = [&]() -> int {
      SSB_CUDA_CHECK(cudaFuncSetAttribute(select_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          8192 * 8));
      return SSB_OK;
    };
    SSB_DEVICE_CONFIG(&select_topk_kernel, 1, configure());
    SSB_CUDA_CHECK(launch_kernel(select_topk_kernel, dim3(B), dim3(1024), smem, stream, 1, cand_, cand_cap_, cand_count_, max_kpts_, sort_cap, ws_,
                                                  hc_, wc_, sx, sy, kp_xy_, kp_score_, kp_cell_, kp_count_));
    count_launch();
    prof_mark(stream, "sp.select");
  }
  if (desc_out != nullptr) {
    dim3 g((max_kpts_ + 7) / 8, B);
    SSB_CUDA_CHECK(launch_kernel(gather_normalize_kernel, dim3(g), dim3(256), 0, stream, 1, grid_, hc_ * wc_, kp_cell_, kp_count_, max_kpts_,
                                                   desc_out));
    count_launch();
    prof_mark(stream, "sp.gather");
  }
  return SSB_OK;
}

int SuperPoint::extract(const uint8_t* const* images, int batch, int h, int w, int row_stride,
                        int channels, float* const* xy, float* const* score, int* count,
                        void** desc_dev, int* slot) {
  SSB_CHECK(images != nullptr && batch >= 1 && batch <= 128, SSB_ERR_INVALID, "bad batch %d", batch);
  SSB_CHECK(channels == 1 || channels == 3, SSB_ERR_INVALID, "channels must be 1 or 3");
  SSB_CHECK(row_stride >= w * channels, SSB_ERR_INVALID, "row_stride smaller than a row");
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  SSB_RETURN_IF(ensure_shape(batch, h, w));
  for (int i = 0; i < batch; ++i) {
    if (slot) slot[i] = -1;
    if (desc_dev) desc_dev[i] = nullptr;
    if (count) count[i] = 0;
  }
  // upload: pinned staging -> device (u8, 1/4 of the reference's fp32 H2D, SuperPoint.cc:825-836)
  const size_t img_bytes = static_cast<size_t>(h) * w * channels;
  uint8_t* hs = staging_host(img_bytes * batch);
  SSB_CHECK(hs != nullptr, SSB_ERR_CUDA, "pinned staging allocation failed");
  for (int i = 0; i < batch; ++i) {
    SSB_CHECK(images[i] != nullptr, SSB_ERR_INVALID, "image %d is null", i);
    for (int y = 0; y < h; ++y)
      std::memcpy(hs + i * img_bytes + static_cast<size_t>(y) * w * channels,
                  images[i] + static_cast<size_t>(y) * row_stride, static_cast<size_t>(w) * channels);
  }
  if (channels == 1) {
    SSB_CUDA_CHECK(cudaMemcpyAsync(img_, hs, img_bytes * batch, cudaMemcpyHostToDevice, stream_));
  } else {
    uint8_t* ds = staging_dev(img_bytes * batch);
    SSB_CHECK(ds != nullptr, SSB_ERR_CUDA, "device staging allocation failed");
    SSB_CUDA_CHECK(cudaMemcpyAsync(ds, hs, img_bytes * batch, cudaMemcpyHostToDevice, stream_));
    const size_t n = static_cast<size_t>(batch) * h * w;
    SSB_CUDA_CHECK(launch_kernel(bgr_to_gray_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, stream_, 1, ds, img_, n));
    count_launch();
    prof_mark(stream_, "sp.bgr2gray");
  }
  // descriptor slots (DescriptorPool::make, SuperPoint.cc:721-727): exhaustion -> no descriptors
  std::vector<void*> ptrs(batch, nullptr);
  int status = SSB_OK;
  for (int i = 0; i < batch; ++i) {
    const int s = pool_.acquire();
    if (s < 0) {
      set_last_error("descriptor pool exhausted (no free slot)");
      status = SSB_ERR_EXHAUSTED;
      continue;
    }
    ptrs[i] = pool_.slot_ptr(s);
    if (slot) slot[i] = s;
    if (desc_dev) desc_dev[i] = ptrs[i];
  }
  void** ptrs_dev = desc_ptrs_dev_;
  SSB_CHECK(ptrs_dev != nullptr, SSB_ERR_CUDA, "pointer table allocation failed");
  SSB_CUDA_CHECK(cudaMemcpyAsync(ptrs_dev, ptrs.data(), batch * sizeof(void*), cudaMemcpyHostToDevice, stream_));
  PdlScope pdl(batch);
  int rs = run(img_, batch, h, w, ptrs_dev, stream_);
  if (rs != SSB_OK) return rs;
  const int K = max_kpts_;
  float* oh = out_host_;
  int* cnt_h = reinterpret_cast<int*>(oh + static_cast<size_t>(batch) * K * 3);
  SSB_CUDA_CHECK(cudaMemcpyAsync(oh, kp_xy_, static_cast<size_t>(batch) * K * 2 * 4, cudaMemcpyDeviceToHost, stream_));
  SSB_CUDA_CHECK(cudaMemcpyAsync(oh + static_cast<size_t>(batch) * K * 2, kp_score_,
                                 static_cast<size_t>(batch) * K * 4, cudaMemcpyDeviceToHost, stream_));
  SSB_CUDA_CHECK(cudaMemcpyAsync(cnt_h, kp_count_, batch * sizeof(int), cudaMemcpyDeviceToHost, stream_));
  SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
  for (int i = 0; i < batch; ++i) {
    const int n = cnt_h[i];
    if (count) count[i] = n;
    if (xy && xy[i]) std::memcpy(xy[i], oh + static_cast<size_t>(i) * K * 2, static_cast<size_t>(n) * 2 * 4);
    if (score && score[i])
      std::memcpy(score[i], oh + static_cast<size_t>(batch) * K * 2 + static_cast<size_t>(i) * K,
                  static_cast<size_t>(n) * 4);
  }
  return status;
}

int SuperPoint::debug_read(const char* what, void* dst, size_t bytes) {
  SSB_CUDA_CHECK(cudaSetDevice(device_));
  SSB_CUDA_CHECK(cudaStreamSynchronize(stream_));
  const std::string k(what ? what : "");
  const size_t B = cap_batch_;
  struct Ent {
    const char* name;
    const void* ptr;
    size_t bytes;
  } tab[] = {
      {"conv1b", a1b_, B * h2_ * w2_ * 64 * 2},
      {"conv2a", a2a_, B * h2_ * w2_ * 64 * 2},     {"conv2b", a2b_, B * h4_ * w4_ * 64 * 2},
      {"conv3a", a3a_, B * h4_ * w4_ * 128 * 2},    {"conv3b", a3b_, B * hc_ * wc_ * 128 * 2},
      {"conv4a", a4a_, B * hc_ * wc_ * 128 * 2},    {"conv4b", a4b_, B * hc_ * wc_ * 128 * 2},
      {"convPaDa", apd_, B * hc_ * wc_ * 512 * 2},  {"grid", grid_, B * hc_ * wc_ * 256 * 2},
      {"scores", scores_, B * hs_ * ws_ * 4},       {"cand_count", cand_count_, B * 4},
      {"kp_cell", kp_cell_, B * max_kpts_ * 4},
  };
  for (const Ent& e : tab) {
    if (k == e.name) {
      SSB_CHECK(e.ptr != nullptr, SSB_ERR_INVALID, "debug_read: '%s' not allocated yet", what);
      SSB_CHECK(bytes <= e.bytes, SSB_ERR_INVALID, "debug_read: '%s' holds %zu bytes, asked %zu", what,
                e.bytes, bytes);
      SSB_CUDA_CHECK(cudaMemcpy(dst, e.ptr, bytes, cudaMemcpyDeviceToHost));
      return SSB_OK;
    }
  }
  set_last_error("debug_read: unknown buffer '%s'", what);
  return SSB_ERR_INVALID;
}

}  // namespace ssb
