// 3x3 convolution as an implicit GEMM with *halo reuse*: the (16+2) x (8b+2) pixel input patch of a
// 16 x 8b output tile is brought into shared memory ONCE by TMA (128B swizzle, zero-filled padding) and
// all nine filter taps read it in place: the A-operand descriptor of each tcgen05.mma simply starts at
// pixel (kh, 8s+kw) of the patch, with the stride between 8-pixel row groups (SBO) equal to the patch
// pitch.  tcgen05 applies the 128B swizzle on absolute shared-memory address bits, so a start address
// shifted by whole 128-byte pixels needs no base-offset (verified on hardware: tools/umma_probe.cu,
// profiles/umma_probe_r01.txt).  Compared with re-loading a shifted box per tap (umma_core.cuh) this
// cuts A traffic from 9 x 128 B to ~1.27 x 128 B per output pixel; weights stream through a small
// TMA ring.  One MMA covers 16 rows x 8 pixels (M = 128); sub-tile s of the CTA owns TMEM columns
// [s*N, (s+1)*N).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue.
#pragma once

#include "common.cuh"
#include "umma_core.cuh"

namespace ssb {

constexpr int kHaloThreads = 192;

struct HaloParams {
  int slabs;       // Cin / 64
  int subtiles;    // b: output tile is 16 rows x 8b pixels
  int block_n;     // output channels per CTA
  int cout_rows;   // rows per tap in the weight matrix
  int tiles_w;     // tiles along W
  int stages;      // weight ring depth
  int tmem_cols;
  int slab_bytes;  // halo bytes per 64-channel slab, rounded up to 1024
  const char* label;
  // fused first layer (kFuse1a): the halo of conv1a's OUTPUT is computed in the kernel from the u8 image
  // (3x3, Cin = 1, fp32 math on fp32 weights, ReLU, fp16) and written straight into the swizzled A
  // layout, so conv1a's 39 MB / image activation never exists in HBM.
  const uint8_t* img;   // [B][img_h][img_w]
  const float* w1a;     // [9][64]
  const float* b1a;     // [64]
  int img_h, img_w;
};

template <class Epi, bool kFuse1a>
__global__ void __launch_bounds__(kHaloThreads)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const HaloParams p, const __grid_constant__ Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  // round the base up to 1 KiB with pointer arithmetic on the __shared__ array itself, so the compiler keeps
  // the shared address space (LDS/STS instead of generic LD/ST with 64-bit address math)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int pw = 8 * p.subtiles + 2;           // patch pitch in pixels
  const int wstage = p.block_n * 128;          // one tap x one slab of weights
  uint8_t* s_halo = smem;
  uint8_t* s_w = smem + p.slabs * p.slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + p.stages * wstage);
  uint64_t* halo_full = bars;                  // [2]
  uint64_t* w_full = bars + 2;                 // [stages]
  uint64_t* w_empty = w_full + p.stages;       // [stages]
  uint64_t* accum_bar = w_empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const int w0 = (blockIdx.x % p.tiles_w) * 8 * p.subtiles;
  const int h0 = (blockIdx.x / p.tiles_w) * 16;
  const int n0 = blockIdx.y * p.block_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    mbar_init(&halo_full[0], 1);
    mbar_init(&halo_full[1], 1);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_w = p.slabs * 9;

  if constexpr (kFuse1a) {
    // conv1a on the fly: image patch (18+2) x (pw+2) -> 18 x pw halo pixels x 64 channels.
    // Halo pixels outside the image are ZERO (they are conv1b's padding), inside pixels see conv1a's
    // own zero padding of the image.  Same tap order / fp32 FMA chain as a stand-alone conv1a.
    const int ppw = pw + 2;
    float* patch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
    float* wsm = patch + ((20 * ppw + 3) & ~3);
    float* bsm = wsm + 576;
    const uint8_t* im = p.img + static_cast<size_t>(z) * p.img_h * p.img_w;
    const float inv255 = 1.0f / 255.0f;  // cv::Mat::convertTo(CV_32F, 1.0/255.0)
    for (int i = threadIdx.x; i < 20 * ppw; i += kHaloThreads) {
      const int r = i / ppw, c = i % ppw;
      const int y = h0 - 2 + r, x = w0 - 2 + c;
      float v = 0.f;
      if (y >= 0 && y < p.img_h && x >= 0 && x < p.img_w) v = static_cast<float>(im[static_cast<size_t>(y) * p.img_w + x]) * inv255;
      patch[i] = v;
    }
    for (int i = threadIdx.x; i < 576; i += kHaloThreads) wsm[i] = p.w1a[i];
    if (threadIdx.x < 64) bsm[threadIdx.x] = p.b1a[threadIdx.x];
    __syncthreads();
    // thread t always owns channel group t % 8 (192 % 8 == 0): its 72 weights + 8 biases stay in registers
    const int g = threadIdx.x & 7;
    float wr[72], br[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) wr[t * 8 + j] = wsm[t * 64 + g * 8 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) br[j] = bsm[g * 8 + j];
    const int npx = 18 * pw;
    int px = threadIdx.x >> 3;             // 24 pixels per sweep
    int hy = px / pw, hx = px - hy * pw;
    for (; px < npx; px += kHaloThreads / 8) {
      const int y = h0 - 1 + hy, x = w0 - 1 + hx;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (y >= 0 && y < p.img_h && x >= 0 && x < p.img_w) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = br[j];
        const float* pp = patch + hy * ppw + hx;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float v = pp[(t / 3) * ppw + (t % 3)];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wr[t * 8 + j], acc[j]);
        }
        o.x = pack_half2(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f));
        o.y = pack_half2(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
        o.z = pack_half2(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f));
        o.w = pack_half2(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
      }
      *reinterpret_cast<uint4*>(s_halo + px * 128 + ((g ^ (px & 7)) << 4)) = o;
      hx += kHaloThreads / 8;
      while (hx >= pw) {
        hx -= pw;
        ++hy;
      }
    }
    fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
  }

  // warps 0 / 1: whole-warp uniform loops, one elected lane issues (operands stay in uniform registers)
  if (warp == 0) {
    const uint32_t halo_tx = static_cast<uint32_t>(18 * pw * 128);
    if (elect_one()) {
      for (int s = 0; s < (kFuse1a ? 0 : p.slabs); ++s) {
        mbar_arrive_expect_tx(&halo_full[s], halo_tx);
        tma_load_4d(s_halo + s * p.slab_bytes, &tmA, &halo_full[s], s * 64, w0 - 1, h0 - 1, z);
      }
    }
    __syncwarp();
    for (int it = 0; it < num_w; ++it) {
      const int slab = it / 9, tap = it % 9;
      const int st = it % p.stages;
      const uint32_t ph = static_cast<uint32_t>(it / p.stages) & 1u;
      mbar_wait(&w_empty[st], ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&w_full[st], static_cast<uint32_t>(wstage));
        tma_load_3d(s_w + st * wstage, &tmB, &w_full[st], slab * 64, tap * p.cout_rows + n0, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(static_cast<uint32_t>(p.block_n));
    const uint32_t sbo = static_cast<uint32_t>(pw * 128);
    const uint32_t w_base = smem_u32(s_w), halo_base = smem_u32(s_halo);
    for (int it = 0; it < num_w; ++it) {
      const int slab = it / 9, tap = it % 9;
      const int kh = tap / 3, kw = tap % 3;
      const int st = it % p.stages;
      if (!kFuse1a && tap == 0) {
        mbar_wait(&halo_full[slab], 0);
      }
      mbar_wait(&w_full[st], static_cast<uint32_t>(it / p.stages) & 1u);
      tc_fence_after();
      const uint64_t bdesc = make_smem_desc_k_sw128(w_base + st * wstage, 1024);
      const uint32_t a_tap = halo_base + slab * p.slab_bytes + static_cast<uint32_t>((kh * pw + kw) * 128);
      if (elect_one()) {
        for (int sub = 0; sub < p.subtiles; ++sub) {
          const uint64_t adesc = make_smem_desc_k_sw128(a_tap + sub * 8 * 128, sbo);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + sub * p.block_n, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&w_empty[st]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(accum_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    EpiCtx c;
    c.row = q * 32 + lane;
    c.lane = lane;
    c.z = z;
    c.n0 = n0;
    c.m_valid = 0x7fffffff;
    c.col_begin = 0;
    c.col_end = p.block_n;
    c.half = 0;
    c.xchg = nullptr;
    // every MMA has completed (accum_bar), so the halo buffer is free: reuse it as TMA-store staging
    c.stage = s_halo + (warp - 2) * 8192;   // 2 x 4 KiB per warp
    c.stage_cur = c.stage;
    c.stage_bufs = 2;
    c.stage_sel = 0;
    c.py = h0 + (c.row >> 3);
    for (int sub = 0; sub < p.subtiles; ++sub) {
      c.px = w0 + sub * 8 + (c.row & 7);
      c.tmem_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + sub * p.block_n;
      epi(c, true);
    }
    stage_drain(c);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
}

inline int halo_slab_bytes(int subtiles) { return (18 * (8 * subtiles + 2) * 128 + 1023) / 1024 * 1024; }

template <class Epi, bool kFuse1a = false>
int launch_conv_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, HaloParams p, const Epi& epi, int W, int H,
                     int batch, int n_tiles, cudaStream_t stream) {
  p.tiles_w = (W + 8 * p.subtiles - 1) / (8 * p.subtiles);
  const int tiles_h = (H + 15) / 16;
  p.slab_bytes = halo_slab_bytes(p.subtiles);
  p.tmem_cols = core_tmem_cols(p.subtiles * p.block_n);
  if (p.subtiles * p.block_n > 512 || p.slabs < 1 || p.slabs > 2 || p.block_n % 16 != 0) {
    set_last_error("launch_conv_halo: unsupported configuration");
    return SSB_ERR_INVALID;
  }
  if (p.stages <= 0) p.stages = 4;
  const int smem = p.slabs * p.slab_bytes + p.stages * p.block_n * 128 + 1024 + 256 +
                   (kFuse1a ? (20 * (8 * p.subtiles + 4) + 4 + 576 + 64) * 4 : 0);
  static int configured = 0;
  if (smem > configured) {
    SSB_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<Epi, kFuse1a>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SSB_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<Epi, kFuse1a>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    configured = smem;
  }
  conv_halo_kernel<Epi, kFuse1a><<<dim3(p.tiles_w * tiles_h, n_tiles, batch), kHaloThreads, smem, stream>>>(tmA, tmB, p, epi);
  SSB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  prof_mark(stream, p.label);
  return SSB_OK;
}

}  // namespace ssb
