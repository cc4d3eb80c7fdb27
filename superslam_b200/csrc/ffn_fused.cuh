// LightGlue's position-wise FFN as ONE kernel per block (cvg/LightGlue lightglue.py TransformerLayer / CrossBlock .ffn;
// oracle/lightglue.py _ffn):
//     x <- x + W2 . GELU(LayerNorm(W1 . [x | msg] + b1)) + b2
// One persistent CTA per SM walks 128-row tiles.  Per tile:
//   G1  acc1[128 x 512] = [x16 | ctx] . W1^T      16 K-slices x 2 N-halves of tcgen05.mma (M 128, N 256): all 512 TMEM columns
//   E1  LayerNorm(512) + exact GELU on the accumulator, fp16 result written back to tensor memory IN PLACE as the
//       A operand of the second product (two fp16 per 32-bit cell, lane = row): the 128 x 512 hidden activation - 134 MB
//       per launch at 64 pairs, written and re-read through HBM by the two-kernel version - never leaves the SM
//   G2  acc2[128 x 256] = h1 . W2^T               tcgen05.mma with the TMEM A operand, two N = 128 accumulators in the
//       columns E1 has freed
//   E2  + b2 + residual (fp32 master, tile-transposed) -> fp32 master and fp16 copy (staged TMA store)
// TMEM columns:   [0,256) acc1 half 0      [256,512) acc1 half 1            after G1
//                 [0,128) h1 k 0..255  [128,256) acc2 n 0..127  [256,384) h1 k 256..511  [384,512) acc2 n 128..255
// (a thread converts its row's columns in increasing order, so the packed fp16 values always land in columns it has
// already read; the two warps that share a lane quadrant own disjoint halves.)
// Weights stream from L2 through a two-stage ring (stage = one 64-wide K chunk: 16 KB of activations + 64 KB of W1,
// or two K chunks of W2); the producer runs ahead, so W2 arrives under E1 and the next tile's first chunks under E2.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogues
// (TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4).
#pragma once

#include "attention.cuh"   // tmem_st_32x16_u32, umma_f16_ts, ffma2 / fadd2
#include "common.cuh"
#include "umma_core.cuh"

namespace ssb {

constexpr int kFfnThreads = 320;
constexpr int kFfnStageBytes = 16384 + 65536;
constexpr int kFfnStages = 2;
constexpr int kFfnStagingBytes = 8 * 2 * 4096;
constexpr int kFfnSmemBytes = kFfnStages * kFfnStageBytes + kFfnStagingBytes + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*xchg*/;

struct FfnParams {
  const float* b1;     // [512] (with the folded out_proj bias)
  const float* ln_g;   // [512]
  const float* ln_b;   // [512]
  const float* b2;     // [256]
  float* x32;          // fp32 master of the residual stream, tile-transposed [z][row/128][col][row%128]
  const int* cnt;      // per-image keypoint counts (device)
  int kp;              // padded rows per image (multiple of 128)
  int tiles_per_img, images;
  const char* label;   // host only
};

// (gelu_erf_h2_batch: lightglue.cu, defined before this header is included)

// tmem_chunks_pipelined (umma_core.cuh) for 2 * kPairs chunks as a ROLLED loop of chunk pairs: fully unrolled over the
// eight chunks of a 256-column half, ptxas hoists every chunk's parameter loads (bias / gamma / beta: 24 float4 per
// chunk) to the top and spills ~1.6 KB per thread.
template <int kPairs, class F>
__device__ __forceinline__ void tmem_chunk_pairs_pipelined(uint32_t t0, F&& f) {
  float va[32], vb[32];
  tmem_ld_32x32(t0, va);
#pragma unroll 1
  for (int i = 0; i < kPairs; ++i) {
    tmem_ld_wait();
    tmem_ld_32x32(t0 + (2 * i + 1) * 32, vb);
    f(2 * i, va);
    tmem_ld_wait();
    if (i + 1 < kPairs) tmem_ld_32x32(t0 + (2 * i + 2) * 32, va);
    f(2 * i + 1, vb);
  }
}

__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmOut, const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;
  uint8_t* staging = ring + kFfnStages * kFfnStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kFfnStagingBytes);
  uint64_t* full_bar = bars;            // [2]
  uint64_t* empty_bar = bars + 2;       // [2]
  uint64_t* acc1_full = bars + 4;       // MMA -> epilogue: G1 has retired
  uint64_t* h1_full = bars + 5;         // epilogue -> MMA: h1 is in tensor memory (8 warps)
  uint64_t* acc2_full = bars + 6;       // MMA -> epilogue: G2 has retired
  uint64_t* tmem_free = bars + 7;       // epilogue -> MMA: acc2 has been read, the next G1 may overwrite (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* xchg = reinterpret_cast<float*>(staging + kFfnStagingBytes + 256);   // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.tiles_per_img * p.images;
  const int stride = static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < kFfnStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(h1_full, 8);
    mbar_init(acc2_full, 1);
    mbar_init(tmem_free, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                // the previous kernel's results (PDL: everything above ran under its tail)
  pdl_launch_dependents();

  // tile -> (image z, first row); tiles that lie beyond the image's keypoint count are skipped by every role alike
  auto decode = [&](int tile, int& z, int& row0) -> bool {
    z = tile / p.tiles_per_img;
    row0 = (tile - z * p.tiles_per_img) * 128;
    return row0 < p.cnt[z];
  };

  if (warp == 0) {
    // ---- producer ----
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total; tile += stride) {
      int z, row0;
      if (!decode(tile, z, row0)) continue;
      for (int kc = 0; kc < 8; ++kc, ++it) {   // G1: activations chunk + the W1 rows of all 512 outputs
        const uint32_t s = it & 1u;
        mbar_wait(&empty_bar[s], ((it >> 1) & 1u) ^ 1u);
        uint8_t* st = ring + s * kFfnStageBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], kFfnStageBytes);
          tma_load_4d(st, kc < 4 ? &tmA0 : &tmA1, &full_bar[s], (kc & 3) * 64, row0, 0, z);
          tma_load_3d(st + 16384, &tmW1, &full_bar[s], kc * 64, 0, 0);
          tma_load_3d(st + 16384 + 32768, &tmW1, &full_bar[s], kc * 64, 256, 0);
        }
        __syncwarp();
      }
      for (int j = 0; j < 4; ++j, ++it) {      // G2: two 64-wide K chunks of W2 per stage
        const uint32_t s = it & 1u;
        mbar_wait(&empty_bar[s], ((it >> 1) & 1u) ^ 1u);
        uint8_t* st = ring + s * kFfnStageBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], 65536);
          tma_load_3d(st, &tmW2, &full_bar[s], (2 * j) * 64, 0, 0);
          tma_load_3d(st + 32768, &tmW2, &full_bar[s], (2 * j + 1) * 64, 0, 0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: whole warp in uniform control flow, one elected lane issues ----
    const uint32_t idesc1 = make_idesc_f16(256);
    const uint32_t idesc2 = make_idesc_f16(128);
    const uint32_t ring_base = smem_u32(ring);
    uint32_t it = 0, seq = 0;
    for (int tile = blockIdx.x; tile < total; tile += stride) {
      int z, row0;
      if (!decode(tile, z, row0)) continue;
      mbar_wait(tmem_free, (seq & 1u) ^ 1u);   // E2 of the previous tile has drained acc2
      tc_fence_after();
      for (int kc = 0; kc < 8; ++kc, ++it) {
        const uint32_t s = it & 1u;
        mbar_wait(&full_bar[s], (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t sa = ring_base + s * kFfnStageBytes;
        const uint64_t adesc = make_smem_desc_k_sw128(sa, 1024);
        if (elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t bdesc = make_smem_desc_k_sw128(sa + 16384 + h * 32768, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem + h * 256, adesc + 2 * k, bdesc + 2 * k, idesc1, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(acc1_full);
      __syncwarp();
      mbar_wait(h1_full, seq & 1u);            // LayerNorm + GELU done, h1 sits in tensor memory
      tc_fence_after();
      for (int j = 0; j < 4; ++j, ++it) {
        const uint32_t s = it & 1u;
        mbar_wait(&full_bar[s], (it >> 1) & 1u);
        tc_fence_after();
        const uint32_t sw = ring_base + s * kFfnStageBytes;
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int ks = (2 * j + c) * 4 + k;                       // K slice of 16 hidden units, 0..31
              const uint32_t a_tmem = tmem + (ks < 16 ? 8 * ks : 256 + 8 * (ks - 16));
#pragma unroll
              for (int h = 0; h < 2; ++h) {                             // output columns 128 h .. 128 h + 127
                const uint64_t bdesc = make_smem_desc_k_sw128(sw + c * 32768 + h * 16384, 1024);
                umma_f16_ts(tmem + 128 + h * 256, a_tmem, bdesc + 2 * k, idesc2, ks != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(acc2_full);
      __syncwarp();
      ++seq;
    }
  } else {
    // ---- epilogues ----
    const int ew = warp - 2, q = warp & 3, half = ew >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_acc1 = tmem + lane_off + half * 256;   // my 256 accumulator columns of G1
    const uint32_t t_h1 = t_acc1;                           // ... and the 128 cells their fp16 results go to
    const uint32_t t_acc2 = tmem + lane_off + 128 + half * 256;   // my 128 output columns of G2
    uint8_t* stage = staging + ew * 8192;
    int stage_sel = 0;
    uint32_t seq = 0;
    for (int tile = blockIdx.x; tile < total; tile += stride) {
      int z, row0;
      if (!decode(tile, z, row0)) continue;
      const bool valid = row0 + row < p.cnt[z];
      const int c0 = half * 256;   // first hidden unit of this thread
      float* xt = p.x32 + (static_cast<size_t>(z) * (p.kp >> 7) + (row0 >> 7)) * (256 * 128) + row +
                  static_cast<size_t>(half * 128) * 128;
      mbar_wait(acc1_full, seq & 1u);
      tc_fence_after();
      // ---- E1 pass 1: sum and sum of squares of (acc + b1) over my 256 columns
      float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
      tmem_chunk_pairs_pipelined<4>(t_acc1, [&](int i, float* v) {
        const float4* b4 = reinterpret_cast<const float4*>(p.b1 + c0 + i * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(b4 + j);
          const float2 x01 = fadd2(make_float2(v[4 * j], v[4 * j + 1]), make_float2(bb.x, bb.y));
          const float2 x23 = fadd2(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(bb.z, bb.w));
          sum2 = fadd2(sum2, fadd2(x01, x23));
          sq2 = ffma2(x01, x01, sq2);
          sq2 = ffma2(x23, x23, sq2);
        }
      });
      // all-reduce across the two column halves (both warps of the lane quadrant)
      xchg[half * 128 + row] = sum2.x + sum2.y;
      epi_pair_sync();
      const float sum = xchg[row] + xchg[128 + row];
      epi_pair_sync();
      xchg[half * 128 + row] = sq2.x + sq2.y;
      epi_pair_sync();
      const float sq = xchg[row] + xchg[128 + row];
      const float mean = sum * (1.0f / 512.0f);
      const float var = fmaxf(sq * (1.0f / 512.0f) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      const float nmr = -mean * rstd;
      const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(nmr, nmr);
      // ---- E1 pass 2: LayerNorm + GELU -> fp16 pairs -> tensor memory, in place (chunk i -> cells 16 i .. 16 i + 15)
      tmem_chunk_pairs_pipelined<4>(t_acc1, [&](int i, float* v) {
        const int col = c0 + i * 32;
        const float4* b4 = reinterpret_cast<const float4*>(p.b1 + col);
        const float4* g4 = reinterpret_cast<const float4*>(p.ln_g + col);
        const float4* be4 = reinterpret_cast<const float4*>(p.ln_b + col);
        uint32_t h[16];
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          __half2 y[8];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = g8 * 4 + jj;
            const float4 bb = __ldg(b4 + j), gg = __ldg(g4 + j), be = __ldg(be4 + j);
            const float2 y01 = ffma2(ffma2(fadd2(make_float2(v[4 * j], v[4 * j + 1]), make_float2(bb.x, bb.y)), rstd2, nmr2),
                                     make_float2(gg.x, gg.y), make_float2(be.x, be.y));
            const float2 y23 = ffma2(ffma2(fadd2(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(bb.z, bb.w)), rstd2, nmr2),
                                     make_float2(gg.z, gg.w), make_float2(be.z, be.w));
            y[2 * jj] = __floats2half2_rn(y01.x, y01.y);
            y[2 * jj + 1] = __floats2half2_rn(y23.x, y23.y);
          }
          gelu_erf_h2_batch(y);
#pragma unroll
          for (int t = 0; t < 8; ++t) h[g8 * 8 + t] = valid ? *reinterpret_cast<const uint32_t*>(&y[t]) : 0u;
        }
        tmem_st_32x16_u32(t_h1 + 16 * i, h);
      });
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(h1_full);
      // ---- E2: + b2 + residual -> fp32 master, fp16 copy.  The residual rows of the first chunk are requested
      // before the wait for G2 (their HBM round trip runs under its 64 MMAs).
      float ra[32];
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) ra[j] = xt[static_cast<size_t>(j) * 128];
      }
      mbar_wait(acc2_full, seq & 1u);
      tc_fence_after();
      {
        const int n0 = half * 128;   // first output column of this thread
        float va[32], rb[32];
        auto load_res = [&](int chunk, float* r) {
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = xt[static_cast<size_t>(chunk * 32 + j) * 128];
          }
        };
        auto chunk = [&](int i, float* v, const float* r) {
          const int col = n0 + i * 32;
          const int hc = i & 1;
          if (hc == 0) {   // next staging buffer: its last store (two blocks ago) has been read out
            if (lane == 0) bulk_wait_read1();
            __syncwarp();
          }
          uint8_t* buf = stage + stage_sel * 4096;
          const float4* b4 = reinterpret_cast<const float4*>(p.b2 + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(b4 + j);
            v[4 * j] += t.x, v[4 * j + 1] += t.y, v[4 * j + 2] += t.z, v[4 * j + 3] += t.w;
          }
          float* p0 = xt + static_cast<size_t>(i * 32) * 128;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = valid ? r[j] + v[j] : 0.f;
            p0[static_cast<size_t>(j) * 128] = x;
            v[j] = x;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_half2(v[8 * j + 0], v[8 * j + 1]);
            o.y = pack_half2(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_half2(v[8 * j + 4], v[8 * j + 5]);
            o.w = pack_half2(v[8 * j + 6], v[8 * j + 7]);
            const int ch = hc * 4 + j;
            *reinterpret_cast<uint4*>(buf + lane * 128 + ((ch ^ (lane & 7)) << 4)) = o;
          }
          if (hc == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmOut, buf, col - 32, row0 + q * 32, z);
              bulk_commit();
            }
            stage_sel ^= 1;
          }
        };
        tmem_ld_32x32(t_acc2, va);
        load_res(1, rb);
        tmem_ld_wait();
        chunk(0, va, ra);
        tmem_ld_32x32(t_acc2 + 32, va);
        load_res(2, ra);
        tmem_ld_wait();
        chunk(1, va, rb);
        tmem_ld_32x32(t_acc2 + 64, va);
        load_res(3, rb);
        tmem_ld_wait();
        chunk(2, va, ra);
        tmem_ld_32x32(t_acc2 + 96, va);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_free);   // acc2 is in registers: the next tile's G1 may start
        chunk(3, va, rb);
      }
      ++seq;
    }
    if (lane == 0) bulk_wait_all();   // outstanding TMA stores still read this CTA's shared memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

inline int launch_ffn_fused(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w1, const CUtensorMap& w2,
                            const CUtensorMap& out, FfnParams p, cudaStream_t stream) {
  auto configure = [&]() -> int {
    SSB_CUDA_CHECK(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmemBytes));
    return SSB_OK;
  };
  SSB_DEVICE_CONFIG(&ffn_fused_kernel, 1, configure());
  const int total = p.tiles_per_img * p.images;
  if (total <= 0) return SSB_OK;
  const int sms = device_sm_count();
  const int ctas = total < sms ? total : sms;
  SSB_CUDA_CHECK(launch_kernel(ffn_fused_kernel, dim3(ctas), dim3(kFfnThreads), kFfnSmemBytes, stream, 1, a0, a1, w1, w2,
                               out, p));
  count_launch();
  prof_mark(stream, p.label);
  return SSB_OK;
}

}  // namespace ssb
