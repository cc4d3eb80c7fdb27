// Persistent, warp-specialised 3x3 convolution for the Cin = 64 layers of the SuperPoint trunk
// (conv1a+conv1b fused, conv2a, conv2b, conv3a): 72 % of the network's FLOPs.
//
// One CTA per SM keeps the 9 x [64 x 64] fp16 weights of its 64-output-channel slice RESIDENT in shared
// memory (73 KB, loaded once by TMA) and walks 16 x 16 pixel tiles.  Per tile:
//   halo   (16+2) x (16+2) pixels x 64 ch, double-buffered: either one TMA box load (zero-filled padding)
//          or - fused first layer - computed in place by eight CUDA-core warps from the u8 image
//          (conv1a: 3x3, Cin = 1, fp32), so conv1a's activation never exists in HBM;
//   MMA    9 taps x 2 sub-tiles x 4 K-slices of tcgen05.mma (M = 128 = 16 rows x 8 px, N = 64); each tap
//          reads the halo in place through a descriptor shifted by whole pixels (see conv_halo.cuh);
//   TMEM   2 x (2 x 64) accumulator columns: the epilogue of tile t (bias, ReLU, optional 2x2 max-pool,
//          fp16, staged TMA store) overlaps the halo production and the MMAs of tile t+1.
// Warp roles: 0 = TMA (weights once, halo boxes), 1 = TMEM alloc + MMA issue, 2..5 = epilogue,
// 6..13 = conv1a producers (fused variant only).
#pragma once

#include "common.cuh"
#include "umma_core.cuh"

namespace ssb {

constexpr int kPipeHaloBytes = 41984;                 // 18*18*128 = 41472, rounded up to 1 KiB
constexpr int kPipeWeightBytes = 9 * 64 * 128;        // 73,728
constexpr int kPipeStagingBytes = 4 * 2 * 4096;      // double-buffered 4 KiB staging per epilogue warp (fused variant)
// TMA-fed variant: THREE halo buffers (a 41 KB box from HBM takes longer than one tile's MMAs, so two
// loads must be in flight) paid for with single-buffered store staging
constexpr int kPipeHaloBufsTma = 3;
constexpr int kPipePatchFloats = 20 * 20;
constexpr int kPipeThreads = 192;
constexpr int kPipeThreadsFused = 192 + 256;
constexpr int kPipeSmemBytes = kPipeWeightBytes + 2 * kPipeHaloBytes + kPipeStagingBytes + 2 * kPipePatchFloats * 4 +
                               (576 + 64) * 4 + 256 + 1024;
constexpr int kPipeSmemBytesTma = kPipeWeightBytes + kPipeHaloBufsTma * kPipeHaloBytes + kPipeStagingBytes / 2 +
                                  2 * kPipePatchFloats * 4 + (576 + 64) * 4 + 256 + 1024;
static_assert(kPipeSmemBytesTma <= 227 * 1024 && kPipeSmemBytes <= 227 * 1024, "conv_pipe shared-memory budget");

// Packed fp32x2 FMA (sm_100: FFMA2): two IEEE fp32 fused multiply-adds per instruction.
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack_f32x2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

struct PipeParams {
  int tiles_w, tiles_h, batch;   // 16x16-pixel tiles per image
  int n_slices;                  // Cout / 64; CTA c serves slice c % n_slices
  int cout_rows;                 // rows per tap in the weight matrix
  const char* label;
  const uint8_t* img;            // fused conv1a: [B][img_h][img_w] u8
  const float* w1a;              // [9][64]
  const float* b1a;              // [64]
  int img_h, img_w;
};

template <class Epi, bool kFuse1a>
__global__ void __launch_bounds__(kFuse1a ? kPipeThreadsFused : kPipeThreads, 1)
conv_pipe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const PipeParams p, const __grid_constant__ Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  // round the base up to 1 KiB with pointer arithmetic on the __shared__ array itself, so the compiler keeps
  // the shared address space (LDS/STS instead of generic LD/ST with 64-bit address math)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_w = smem;
  constexpr int kHaloBufs = kFuse1a ? 2 : kPipeHaloBufsTma;
  constexpr int kStageBufs = kFuse1a ? 2 : 1;
  uint8_t* s_halo = smem + kPipeWeightBytes;                  // [kHaloBufs]
  uint8_t* s_stage = s_halo + kHaloBufs * kPipeHaloBytes;     // 4 warps x kStageBufs x 4 KiB
  float* s_patch = reinterpret_cast<float*>(s_stage + 4 * kStageBufs * 4096);   // [2][400]
  float* s_w1a = s_patch + 2 * kPipePatchFloats;              // [576]
  float* s_b1a = s_w1a + 576;                                 // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b1a + 64);
  uint64_t* w_full = bars;
  uint64_t* halo_full = bars + 1;    // [3]
  uint64_t* halo_empty = bars + 4;   // [3]
  uint64_t* tmem_full = bars + 7;    // [2]
  uint64_t* tmem_empty = bars + 9;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.n_slices;
  const int first = blockIdx.x / p.n_slices;
  const int stride = gridDim.x / p.n_slices;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int total = tiles_per_img * p.batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    mbar_init(w_full, 1);
    for (int b = 0; b < kHaloBufs; ++b) {
      mbar_init(&halo_full[b], kFuse1a ? 8 : 1);   // one arrival per producer warp
      mbar_init(&halo_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 4);                // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  if (kFuse1a) {
    for (int i = threadIdx.x; i < 576; i += blockDim.x) s_w1a[i] = p.w1a[i];
    if (threadIdx.x < 64) s_b1a[threadIdx.x] = p.b1a[threadIdx.x];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, kPipeWeightBytes);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_3d(s_w + tap * 8192, &tmB, w_full, 0, tap * p.cout_rows + slice * 64, 0);
      if (!kFuse1a) {
        int seq = 0;
        for (int t = first; t < total; t += stride, ++seq) {
          const int hb = seq % kHaloBufs;
          const int z = t / tiles_per_img, r = t % tiles_per_img;
          const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
          mbar_wait(&halo_empty[hb], (static_cast<uint32_t>(seq / kHaloBufs) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&halo_full[hb], 18 * 18 * 128);
          tma_load_4d(s_halo + hb * kPipeHaloBytes, &tmA, &halo_full[hb], 0, w0 - 1, h0 - 1, z);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(64);
      mbar_wait(w_full, 0);
      int seq = 0;
      for (int t = first; t < total; t += stride, ++seq) {
        const int b = seq & 1;
        const int hb = seq % kHaloBufs;
        const uint32_t use = static_cast<uint32_t>(seq >> 1) & 1u;
        mbar_wait(&tmem_empty[b], use ^ 1u);
        mbar_wait(&halo_full[hb], static_cast<uint32_t>(seq / kHaloBufs) & 1u);
        tc_fence_after();
        const uint32_t hbase = smem_u32(s_halo + hb * kPipeHaloBytes);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int kh = tap / 3, kw = tap % 3;
          const uint64_t bdesc = make_smem_desc_k_sw128(smem_u32(s_w + tap * 8192), 1024);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint64_t adesc = make_smem_desc_k_sw128(hbase + ((kh * 18 + kw) + sub * 8) * 128, 18 * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + b * 128 + sub * 64, adesc + 2 * k, bdesc + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&halo_empty[hb]);
        umma_commit(&tmem_full[b]);
      }
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    int seq = 0;
    EpiCtx c;
    c.row = q * 32 + lane;
    c.lane = lane;
    c.n0 = slice * 64;
    c.m_valid = 0x7fffffff;
    c.col_begin = 0;
    c.col_end = 64;
    c.half = 0;
    c.xchg = nullptr;
    c.stage = s_stage + (warp - 2) * (kStageBufs * 4096);
    c.stage_cur = c.stage;
    c.stage_bufs = kStageBufs;
    c.stage_sel = 0;
    for (int t = first; t < total; t += stride, ++seq) {
      const int b = seq & 1;
      const int z = t / tiles_per_img, r = t % tiles_per_img;
      const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
      mbar_wait(&tmem_full[b], static_cast<uint32_t>(seq >> 1) & 1u);
      tc_fence_after();
      c.z = z;
      c.py = h0 + (c.row >> 3);
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        c.px = w0 + sub * 8 + (c.row & 7);
        c.tmem_row = tmem_base + b * 128 + sub * 64 + (static_cast<uint32_t>(q * 32) << 16);
        epi(c, true);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[b]);
    }
    stage_drain(c);
  } else if (kFuse1a) {
    // conv1a producers: 256 threads; thread owns channel group (tid & 7) with its weights in registers.
    const int tid = threadIdx.x - 192;
    const int g = tid & 7;
    // weights / bias as fp32x2 pairs for FFMA2 (same per-element IEEE fma as the scalar chain)
    unsigned long long wr[36], br[4];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        wr[tp * 4 + j] = pack_f32x2(s_w1a[tp * 64 + g * 8 + 2 * j], s_w1a[tp * 64 + g * 8 + 2 * j + 1]);
#pragma unroll
    for (int j = 0; j < 4; ++j) br[j] = pack_f32x2(s_b1a[g * 8 + 2 * j], s_b1a[g * 8 + 2 * j + 1]);
    const float inv255 = 1.0f / 255.0f;  // cv::Mat::convertTo(CV_32F, 1.0/255.0): value * float(1/255)
    // image patch (20 x 20 pixels around the tile) of tile t, two elements per thread, as fp32 * (1/255)
    auto load_patch = [&](int t, float* pre) {
      const int z = t / tiles_per_img, r = t % tiles_per_img;
      const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
      const uint8_t* im = p.img + static_cast<size_t>(z) * p.img_h * p.img_w;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = tid + k * 256;
        const int pr = i / 20, pc = i - pr * 20;
        const int y = h0 - 2 + pr, x = w0 - 2 + pc;
        float v = 0.f;
        if (i < kPipePatchFloats && y >= 0 && y < p.img_h && x >= 0 && x < p.img_w)
          v = static_cast<float>(__ldg(im + static_cast<size_t>(y) * p.img_w + x)) * inv255;
        pre[k] = v;
      }
    };
    auto store_patch = [&](float* patch, const float* pre) {
      patch[tid] = pre[0];
      if (tid + 256 < kPipePatchFloats) patch[tid + 256] = pre[1];
    };
    int seq = 0;
    {
      float pre[2];
      if (first < total) {
        load_patch(first, pre);
        store_patch(s_patch, pre);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
    }
    for (int t = first; t < total; t += stride, ++seq) {
      const int hb = seq & 1;
      const int r = t % tiles_per_img;
      const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
      const float* patch = s_patch + hb * kPipePatchFloats;
      // software pipeline: the global loads of the NEXT tile's patch are in flight while this tile's
      // halo is computed; they are published to the other patch buffer at the end of the iteration
      float pre[2];
      const int tn = t + stride;
      if (tn < total) load_patch(tn, pre);
      mbar_wait(&halo_empty[hb], (static_cast<uint32_t>(seq >> 1) & 1u) ^ 1u);
      uint8_t* halo = s_halo + hb * kPipeHaloBytes;
      int px = tid >> 3;   // 32 halo pixels per sweep
      int hy = px / 18, hx = px - hy * 18;
      for (; px < 18 * 18; px += 32) {
        const int y = h0 - 1 + hy, x = w0 - 1 + hx;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);   // outside the image: conv1b's zero padding
        if (y >= 0 && y < p.img_h && x >= 0 && x < p.img_w) {
          unsigned long long acc[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = br[j];
          const float* pp = patch + hy * 20 + hx;
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            const float v = pp[(tp / 3) * 20 + (tp % 3)];
            const unsigned long long vv = pack_f32x2(v, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = ffma2(vv, wr[tp * 4 + j], acc[j]);
          }
          const float2 a0 = unpack_f32x2(acc[0]), a1 = unpack_f32x2(acc[1]), a2 = unpack_f32x2(acc[2]),
                       a3 = unpack_f32x2(acc[3]);
          o.x = pack_half2(fmaxf(a0.x, 0.f), fmaxf(a0.y, 0.f));
          o.y = pack_half2(fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f));
          o.z = pack_half2(fmaxf(a2.x, 0.f), fmaxf(a2.y, 0.f));
          o.w = pack_half2(fmaxf(a3.x, 0.f), fmaxf(a3.y, 0.f));
        }
        *reinterpret_cast<uint4*>(halo + px * 128 + ((g ^ (px & 7)) << 4)) = o;
        hx += 32;
        while (hx >= 18) {
          hx -= 18;
          ++hy;
        }
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&halo_full[hb]);
      if (tn < total) store_patch(s_patch + (hb ^ 1) * kPipePatchFloats, pre);
      asm volatile("bar.sync 3, 256;" ::: "memory");   // patch[hb^1] published, patch[hb] no longer read
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

template <class Epi, bool kFuse1a>
int launch_conv_pipe(const CUtensorMap& tmA, const CUtensorMap& tmB, PipeParams p, const Epi& epi, int W, int H,
                     int batch, cudaStream_t stream) {
  p.tiles_w = (W + 15) / 16;
  p.tiles_h = (H + 15) / 16;
  p.batch = batch;
  constexpr int smem_bytes = kFuse1a ? kPipeSmemBytes : kPipeSmemBytesTma;
  static bool configured = false;
  if (!configured) {
    SSB_CUDA_CHECK(cudaFuncSetAttribute(conv_pipe_kernel<Epi, kFuse1a>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        smem_bytes));
    configured = true;
  }
  const long long total = static_cast<long long>(p.tiles_w) * p.tiles_h * batch;
  int ctas = device_sm_count() / p.n_slices * p.n_slices;
  if (total * p.n_slices < ctas) ctas = static_cast<int>(total) * p.n_slices;
  conv_pipe_kernel<Epi, kFuse1a><<<ctas, kFuse1a ? kPipeThreadsFused : kPipeThreads, smem_bytes, stream>>>(
      tmA, tmB, p, epi);
  SSB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  prof_mark(stream, p.label);
  return SSB_OK;
}

}  // namespace ssb
