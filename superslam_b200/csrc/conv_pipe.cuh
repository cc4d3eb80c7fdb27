// Persistent, warp-specialised 3x3 convolution for the Cin = 64 layers of the SuperPoint trunk
// (conv1a+conv1b fused, conv2a, conv2b, conv3a): 72 % of the network's FLOPs.
//
// One CTA per SM - by default one CTA PAIR per TPC (kPair below: tcgen05.mma.cta_group::2, each CTA keeps half of the
// output channels' weights) - keeps the 9 x [64 x 64] fp16 weights of its 64-output-channel slice RESIDENT in shared
// memory (73 KB, loaded once by TMA) and walks 16 x 16 pixel tiles.  Per tile:
//   halo   (16+2) x (16+2) pixels x 64 ch: either TMA box loads (zero-filled padding, two buffers) or -
//          fused first layer - produced on chip from the u8 image, so conv1a's activation never exists in
//          HBM: conv1a (3x3, Cin = 1) is itself a tensor-core product  [384 halo px x 16] x [16 x 64]  with
//          K = 9 taps + a ones column that carries the bias; the weights are split hi + lo (two fp16) so the
//          result is fp32-accurate.  Eight converter warps build the im2col operand (un-swizzled K-major
//          core matrices), read the product back from TMEM, apply ReLU / the image border and write the
//          fp16 halo in the 128B-swizzled layout the conv1b MMAs read;
//   MMA    9 taps x 2 sub-tiles x 4 K-slices of tcgen05.mma (M = 128 = 16 rows x 8 px, N = 64); each tap
//          reads the halo in place through a descriptor shifted by whole pixels (the 128B swizzle works on
//          absolute shared-memory address bits: tools/umma_probe.cu mode 0);
//   TMEM   2 x (2 x 64) accumulator columns: the epilogue of tile t (bias, ReLU, optional 2x2 max-pool,
//          fp16, staged TMA store) overlaps the halo production and the MMAs of tile t+1.
// Warp roles: 0 = TMA (weights once, halo boxes), 1 = TMEM alloc + MMA issue, 2..9 = epilogue,
// 10..17 = conv1a im2col builders / TMEM->halo converters (fused variant only).
// What bounds these kernels is the shared-memory port (profiles/README.md, round 2): an N = 64 MMA reads 6 KB of
// operands per 32 cycles of math = 48 cycles at 128 B/clk; halo, im2col and staging traffic shares the port.
#pragma once

#include "common.cuh"
#include "umma_core.cuh"

namespace ssb {

constexpr int kPipeHaloBytes = 41984;                 // 18*18*128 = 41472, rounded up to 1 KiB
constexpr int kPipeWeightBytes = 9 * 64 * 128;        // 73,728
constexpr int kPipeEpiWarps = 8;                     // two per TMEM lane quadrant: one per 128-pixel sub-tile
constexpr int kPipeStagingBytes = kPipeEpiWarps * 4096;   // one 4 KiB TMA-store staging buffer per epilogue warp
constexpr int kPipeHaloBufsTma = 2;
constexpr int kPipePatchElems = 20 * 20;               // u8 patch around a tile, kept as fp16 (value / 16, exact)
constexpr int kPipeA1Bytes = 384 * 32;                  // im2col operand of conv1a: 3 M-tiles x 128 rows x 16 fp16
constexpr int kPipeW1Bytes = 2 * 64 * 32;               // conv1a weights [64 x 16] fp16, hi and lo parts
constexpr float kPipeImgScale = 1.0f / 16.0f;           // A = u8 / 16 (exact), B = w * 16 / 255
constexpr int kPipeThreads = 64 + kPipeEpiWarps * 32;
constexpr int kPipeThreadsFused = kPipeThreads + 256;
constexpr int kPipeSmemBytes = kPipeWeightBytes + 2 * kPipeHaloBytes + kPipeStagingBytes + 2 * kPipeA1Bytes +
                               kPipeW1Bytes + 2048 /*patches*/ + 256 + 1024;
constexpr int kPipeSmemBytesTma = kPipeWeightBytes + kPipeHaloBufsTma * kPipeHaloBytes + 2 * kPipeStagingBytes +
                                  256 + 1024;
static_assert(kPipeSmemBytesTma <= 227 * 1024 && kPipeSmemBytes <= 227 * 1024, "conv_pipe shared-memory budget");

// Shared-memory descriptor of an UN-swizzled K-major operand made of dense 8-row x 16-byte core matrices:
// element (r, k) lives at (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2, i.e. LBO (K direction) = 128 B and
// SBO (next 8 rows) = 256 B for a K = 16 slice (checked on hardware: tools/umma_probe.cu mode 2).
__device__ __forceinline__ uint64_t make_smem_desc_k_plain16(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(128 >> 4) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ int plain16_offset(int r, int k) {
  return (r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2;
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

// kPair: CTA pairs (cta_group::2, see common.cuh).  The two CTAs of a cluster walk neighbouring tiles in lockstep; the
// leader's MMA warp issues M = 256 MMAs for both (each CTA: its own halo as A, 32 of the 64 output channels as B, its own
// 128 pixels x 64 channels in its tensor memory).  The N = 64 MMAs of these layers are bound by shared-memory operand
// reads (4 KB of A + 2 KB of B per 32-cycle MMA against 128 B/clk); a pair reads 4 + 1 KB per CTA.  Barriers the issuer
// waits on (halo_full, tmem_empty, a1_full, w_full) live in the leader and count both CTAs' arrivals; barriers the MMAs
// complete (halo_empty, tmem_full, c1_full) are signalled in both CTAs by one multicast commit.  Needs n_slices == 1.
// kN = 128 (pairs only, not fused): all 128 output channels of conv3a in one MMA stream - each CTA holds 64 of them, the
// N = 128 MMAs are math-bound (64 cycles against 48 of operand reads) and the halo is loaded once instead of once per slice.
template <class Epi, bool kFuse1a, bool kPair, int kN = 64>
__global__ void __launch_bounds__(kFuse1a ? kPipeThreadsFused : kPipeThreads, 1)
conv_pipe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const PipeParams p, const __grid_constant__ Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  // round the base up to 1 KiB with pointer arithmetic on the __shared__ array itself, so the compiler keeps
  // the shared address space (LDS/STS instead of generic LD/ST with 64-bit address math)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_w = smem;
  constexpr int kHaloBufs = kFuse1a ? 2 : kPipeHaloBufsTma;
  constexpr int kStageBufs = kFuse1a ? 1 : 2;   // the fused variant needs the room for the conv1a operands
  uint8_t* s_halo = smem + kPipeWeightBytes;                  // [kHaloBufs]
  uint8_t* s_stage = s_halo + kHaloBufs * kPipeHaloBytes;     // 8 warps x kStageBufs x 4 KiB
  uint8_t* s_a1 = s_stage + kStageBufs * kPipeStagingBytes;   // fused: [2] im2col operands
  uint8_t* s_w1 = s_a1 + (kFuse1a ? 2 * kPipeA1Bytes : 0);    // fused: conv1a weights hi | lo
  __half* s_patch = reinterpret_cast<__half*>(s_w1 + (kFuse1a ? kPipeW1Bytes : 0));   // fused: [2][400]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_patch) + (kFuse1a ? 2048 : 0));
  uint64_t* w_full = bars;
  uint64_t* halo_full = bars + 1;    // [3]
  uint64_t* halo_empty = bars + 4;   // [3]
  uint64_t* tmem_full = bars + 7;    // [2]
  uint64_t* tmem_empty = bars + 9;   // [2]
  uint64_t* a1_full = bars + 11;     // [2] fused: im2col operand built (8 warp arrivals)
  uint64_t* c1_full = bars + 13;     // fused: conv1a product in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  volatile uint32_t* wbase_slot = tmem_slot + 1;   // shared-memory address of the weights, re-read per tile (see the MMA warp)
  static_assert(kN == 64 || (kN == 128 && kPair && !kFuse1a), "128 output channels: CTA pairs, TMA-fed variant only");
  constexpr uint32_t kTmemCols = (kFuse1a || kN == 128) ? 512 : 256;   // accumulators 2 x (2 x kN), conv1a product 3 x 64
  constexpr uint32_t kC1Col = 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = kPair ? static_cast<int>(cluster_ctarank()) : 0;   // 0 = leader (issues the pair's MMAs)
  const int slice = kPair ? 0 : blockIdx.x % p.n_slices;
  // tile walk: `tl` runs over the leader's tiles (identical in both CTAs of a pair, so that they stay in lockstep); this
  // CTA's tile is tl + rank, clamped to the last one (an odd total: the peer recomputes that tile, same values)
  const int first = kPair ? static_cast<int>(blockIdx.x) - rank : blockIdx.x / p.n_slices;
  const int stride = kPair ? static_cast<int>(gridDim.x) : gridDim.x / p.n_slices;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int total = tiles_per_img * p.batch;
  auto my_tile = [&](int tl) { return kPair ? min(tl + rank, total - 1) : tl; };
  constexpr uint32_t kArrivals = kPair ? 2u : 1u;   // CTAs arriving on the leader's barriers
  constexpr int kWTapBytes = kPair ? kN * 64 : 8192;   // one tap of this CTA's weights: kN / 2 (pair) or 64 output channels x 64

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    mbar_init(w_full, 1);
    for (int b = 0; b < kHaloBufs; ++b) {
      mbar_init(&halo_full[b], kFuse1a ? 8 * kArrivals : 1);   // one arrival per producer warp (of both CTAs)
      mbar_init(&halo_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], kPipeEpiWarps * kArrivals);    // one arrival per epilogue warp
      mbar_init(&a1_full[b], 8 * kArrivals);
    }
    mbar_init(c1_full, 1);
    fence_mbar_init();
    *wbase_slot = smem_u32(s_w);
  }
  if (warp == 1) {
    if (kPair) {
      tmem_alloc2(tmem_slot, kTmemCols);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  if (kFuse1a) {
    // conv1a B operand [n = 64][k = 16]: k < 9 tap weights * 16/255, k = 9 the bias, rest 0; hi + lo fp16
    // (pair: this CTA's 32 output channels, rows 0..31)
    constexpr int kRows = kPair ? 32 : 64;
    for (int i = threadIdx.x; i < kRows * 16; i += blockDim.x) {
      const int r = i >> 4, k = i & 15;
      const int n = rank * kRows + r;
      const float v = k < 9 ? p.w1a[k * 64 + n] * (16.0f / 255.0f) : (k == 9 ? p.b1a[n] : 0.f);
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      *reinterpret_cast<__half*>(s_w1 + plain16_offset(r, k)) = hi;
      *reinterpret_cast<__half*>(s_w1 + 2048 + plain16_offset(r, k)) = lo;
    }
    // im2col rows 324..383 of both operands stay zero for the whole kernel
    for (int i = threadIdx.x; i < 2 * kPipeA1Bytes / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(s_a1)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // the leader's barriers exist before the peer arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                // the previous kernel's results (PDL: everything above ran under its tail)
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      if (kPair) {   // both CTAs' halves are counted on the leader's barrier
        if (rank == 0) mbar_arrive_expect_tx(w_full, 2 * 9 * kWTapBytes);
        for (int tap = 0; tap < 9; ++tap)
          tma_load_3d_pair(s_w + tap * kWTapBytes, &tmB, w_full, 0, tap * p.cout_rows + rank * (kN / 2), 0);
      } else {
        mbar_arrive_expect_tx(w_full, kPipeWeightBytes);
        for (int tap = 0; tap < 9; ++tap)
          tma_load_3d(s_w + tap * 8192, &tmB, w_full, 0, tap * p.cout_rows + slice * 64, 0);
      }
      if (!kFuse1a) {
        int seq = 0;
        for (int tl = first; tl < total; tl += stride, ++seq) {
          const int t = my_tile(tl);
          const int hb = seq % kHaloBufs;
          const int z = t / tiles_per_img, r = t % tiles_per_img;
          const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
          mbar_wait(&halo_empty[hb], (static_cast<uint32_t>(seq / kHaloBufs) & 1u) ^ 1u);
          if (kPair) {
            if (rank == 0) mbar_arrive_expect_tx(&halo_full[hb], 2 * 18 * 18 * 128);
            tma_load_4d_pair(s_halo + hb * kPipeHaloBytes, &tmA, &halo_full[hb], 0, w0 - 1, h0 - 1, z);
          } else {
            mbar_arrive_expect_tx(&halo_full[hb], 18 * 18 * 128);
            tma_load_4d(s_halo + hb * kPipeHaloBytes, &tmA, &halo_full[hb], 0, w0 - 1, h0 - 1, z);
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // The whole warp walks the tile loop and computes descriptors in warp-uniform control flow, so they live
    // in uniform registers and each tcgen05.mma is one instruction for the elected lane.  (Inside an
    // `if (lane == 0)` region the operands are vector registers and every MMA becomes an
    // ELECT / R2UR x3 / branch "waterfall" of ~100 cycles - three times the 32 cycles an N = 64 MMA takes.)
    const uint32_t idesc = kPair ? make_idesc2_f16(kN) : make_idesc_f16(64);
    const uint32_t w_base = smem_u32(s_w), halo_base = smem_u32(s_halo);
    const uint32_t a1_base = smem_u32(s_a1), w1_base = smem_u32(s_w1);
    auto wait = [&](uint64_t* bar, uint32_t parity) {   // barriers the peer arrives on need the cluster-scope acquire
      if (kPair) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity);
    };
    auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
      if (kPair) umma2_f16(d, a, b, idesc, acc); else umma_f16(d, a, b, idesc, acc);
    };
    auto commit = [&](uint64_t* bar) {
      if (kPair) umma2_commit(bar); else umma_commit(bar);
    };
    wait(w_full, 0);
    // fused: conv1a of tile `seq1` = 3 M-tiles x (hi + lo) MMAs with K = 16 into TMEM columns kC1Col..
    auto issue_conv1a = [&](int seq1) {
      wait(&a1_full[seq1 & 1], static_cast<uint32_t>(seq1 >> 1) & 1u);
      tc_fence_after();
      const uint32_t a1 = a1_base + (seq1 & 1) * kPipeA1Bytes;
      const uint64_t bhi = make_smem_desc_k_plain16(w1_base), blo = make_smem_desc_k_plain16(w1_base + 2048);
      if (elect_one()) {
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const uint64_t ad = make_smem_desc_k_plain16(a1 + m * 4096);
          mma(tmem_base + kC1Col + m * 64, ad, bhi, 0u);
          mma(tmem_base + kC1Col + m * 64, ad, blo, 1u);
        }
        commit(c1_full);
      }
      __syncwarp();
    };
    int seq = 0;
    if (kFuse1a && first < total) issue_conv1a(0);
    for (int t = first; t < total; t += stride, ++seq) {
      const int b = seq & 1;
      const int hb = seq % kHaloBufs;
      const uint32_t use = static_cast<uint32_t>(seq >> 1) & 1u;
      wait(&halo_full[hb], static_cast<uint32_t>(seq / kHaloBufs) & 1u);
      tc_fence_after();
      // fused: the converters are done with the conv1a product of this tile (halo_full), so the next
      // tile's conv1a goes first; its conversion then overlaps this tile's 72 conv1b MMAs
      if (kFuse1a && t + stride < total) issue_conv1a(seq + 1);
      wait(&tmem_empty[b], use ^ 1u);
      tc_fence_after();
      const uint32_t hbase = halo_base + hb * kPipeHaloBytes;
      const uint32_t d0 = tmem_base + b * (2 * kN);
      // The two base descriptors are formed HERE, in warp-uniform code, and every operand of the 72 MMAs is base +
      // compile-time constant (the start-address field counts 16-byte units; halo and weights sit below 256 KB, so the
      // sum never carries out of its 14 bits).  Formed inside the elected region they were vector-register values:
      // 88 R2UR + spills per tile, the issuing thread spent 70 % of its time in this loop and delivered an MMA every
      // ~45 cycles - no faster than the tensor pipe drains them (48), so every hiccup at a tile boundary cost tensor time.
      // The weights never move, so the compiler hoists their 36 descriptors out of the tile loop - into VECTOR registers
      // (there are not enough uniform ones), and every MMA then starts with two R2UR.  Reading the address back from
      // shared memory per tile makes it loop-variant: one R2UR per tile, the 36 descriptors are uniform adds.
      const uint64_t a_base = make_smem_desc_k_sw128(hbase, 18 * 128);
      const uint64_t b_base = make_smem_desc_k_sw128(*wbase_slot, 1024);
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int kh = tap / 3, kw = tap % 3;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma(d0 + sub * kN, a_base + static_cast<uint64_t>((((kh * 18 + kw) + sub * 8) * 128) >> 4) + 2 * k,
                  b_base + static_cast<uint64_t>((tap * kWTapBytes) >> 4) + 2 * k, (tap | k) != 0 ? 1u : 0u);
          }
        }
        commit(&halo_empty[hb]);
        commit(&tmem_full[b]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // pair: the peer's MMA warp has nothing to issue
  } else if (warp < 2 + kPipeEpiWarps) {
    // eight epilogue warps: warp w reads TMEM lanes 32*(w%4)..+31 (hardware rule) of sub-tile (w-2)/4.
    // (With four warps - one per scheduler - the dependent chain TMEM load -> bias/ReLU -> pooling
    // shuffles -> staging store ran at ~7 cycles per instruction and set the pace of the whole kernel.)
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    int seq = 0;
    EpiCtx c;
    c.row = q * 32 + lane;
    c.lane = lane;
    c.n0 = slice * 64;
    c.m_valid = 0x7fffffff;
    c.col_begin = 0;
    c.col_end = kN;
    c.half = 0;
    c.xchg = nullptr;
    c.stage = s_stage + (warp - 2) * (kStageBufs * 4096);
    c.stage_cur = c.stage;
    c.stage_bufs = kStageBufs;
    c.stage_sel = 0;
    for (int tl = first; tl < total; tl += stride, ++seq) {
      const int t = my_tile(tl);
      const int b = seq & 1;
      const int z = t / tiles_per_img, r = t % tiles_per_img;
      const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
      mbar_wait(&tmem_full[b], static_cast<uint32_t>(seq >> 1) & 1u);
      tc_fence_after();
      c.z = z;
      c.py = h0 + (c.row >> 3);
      c.px = w0 + sub * 8 + (c.row & 7);
      c.tmem_row = tmem_base + b * (2 * kN) + sub * kN + (static_cast<uint32_t>(q * 32) << 16);
      epi(c, true);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_remote(&tmem_empty[b], 0); else mbar_arrive(&tmem_empty[b]);
      }
    }
    stage_drain(c);
  } else if (kFuse1a) {
    // ---- conv1a: im2col builders + TMEM->halo converters (8 warps, 256 threads) ----
    const int tid = threadIdx.x - kPipeThreads;
    const int cw = warp - (2 + kPipeEpiWarps);
    const int q = warp & 3;            // TMEM lane quadrant this warp may read
    const int colhalf = cw >> 2;       // the two warps of a quadrant split the 64 channels
    // u8 patch (20 x 20 pixels around the tile) of tile t -> two fp16 (value / 16) per thread
    auto load_patch = [&](int t, __half* pre) {
      const int z = t / tiles_per_img, r = t % tiles_per_img;
      const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
      const uint8_t* im = p.img + static_cast<size_t>(z) * p.img_h * p.img_w;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int i = tid + k * 256;
        const int pr = i / 20, pc = i - pr * 20;
        const int y = h0 - 2 + pr, x = w0 - 2 + pc;
        float v = 0.f;   // conv1a's own zero padding
        if (i < kPipePatchElems && y >= 0 && y < p.img_h && x >= 0 && x < p.img_w)
          v = static_cast<float>(__ldg(im + static_cast<size_t>(y) * p.img_w + x)) * kPipeImgScale;
        pre[k] = __float2half_rn(v);
      }
    };
    auto store_patch = [&](__half* patch, const __half* pre) {
      patch[tid] = pre[0];
      if (tid + 256 < kPipePatchElems) patch[tid + 256] = pre[1];
    };
    // im2col: row p = halo pixel (hy, hx); k = 3*kh + kw -> patch[hy + kh][hx + kw]; k = 9 -> 1 (bias)
    auto build_a1 = [&](const __half* patch, uint8_t* a1) {
#pragma unroll 1
      for (int px = tid; px < 18 * 18; px += 256) {
        const int hy = px / 18, hx = px - hy * 18;
        const unsigned short* pp = reinterpret_cast<const unsigned short*>(patch) + hy * 20 + hx;
        uint32_t e[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) e[k] = pp[(k / 3) * 20 + (k % 3)];
        const uint4 lo = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
        const uint4 hi = make_uint4(e[8] | (0x3C00u << 16) /* fp16 1.0: bias column */, 0u, 0u, 0u);
        uint8_t* dst = a1 + (px >> 3) * 256 + (px & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = lo;
        *reinterpret_cast<uint4*>(dst + 128) = hi;
      }
    };
    // arrivals the issuer waits for go to the leader's barrier (pair) or this CTA's
    auto arrive = [&](uint64_t* bar) {
      if (kPair) mbar_arrive_remote(bar, 0); else mbar_arrive(bar);
    };
    int seq = 0;
    if (first < total) {
      __half pre[2];
      load_patch(my_tile(first), pre);
      store_patch(s_patch, pre);
      asm volatile("bar.sync 3, 256;" ::: "memory");
      build_a1(s_patch, s_a1);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive(&a1_full[0]);
      if (first + stride < total) {
        load_patch(my_tile(first + stride), pre);
        store_patch(s_patch + kPipePatchElems, pre);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
    }
    for (int tl = first; tl < total; tl += stride, ++seq) {
      const int t = my_tile(tl);
      const int hb = seq & 1;
      const int r = t % tiles_per_img;
      const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
      const int tn = tl + stride, tnn = tl + 2 * stride;   // the leader's next tiles: the loop bounds of both CTAs
      // software pipeline: the global loads of the patch two tiles ahead are in flight during this iteration
      __half pre[2];
      if (tnn < total) load_patch(my_tile(tnn), pre);
      // (1) im2col operand of the NEXT tile (its buffer was last read by conv1a of tile seq-1, whose
      //     completion this warp observed through c1_full one iteration ago)
      if (tn < total) {
        build_a1(s_patch + ((seq + 1) & 1) * kPipePatchElems, s_a1 + ((seq + 1) & 1) * kPipeA1Bytes);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive(&a1_full[(seq + 1) & 1]);
      }
      // (2) this tile's conv1a product -> ReLU -> fp16 halo (zero outside the image = conv1b's padding)
      mbar_wait(c1_full, static_cast<uint32_t>(seq) & 1u);
      mbar_wait(&halo_empty[hb], (static_cast<uint32_t>(seq >> 1) & 1u) ^ 1u);
      tc_fence_after();
      uint8_t* halo = s_halo + hb * kPipeHaloBytes;
#pragma unroll 1
      for (int m = 0; m < 3; ++m) {
        if (m * 128 + q * 32 >= 18 * 18) break;   // warp-uniform: rows beyond the halo
        const int px = m * 128 + q * 32 + lane;
        float v[32];
        tmem_ld_32x32(tmem_base + kC1Col + m * 64 + colhalf * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        if (px < 18 * 18) {
          const int hy = px / 18, hx = px - hy * 18;
          const int y = h0 - 1 + hy, x = w0 - 1 + hx;
          const bool inside = y >= 0 && y < p.img_h && x >= 0 && x < p.img_w;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (inside) {
              o.x = pack_half2(fmaxf(v[8 * j + 0], 0.f), fmaxf(v[8 * j + 1], 0.f));
              o.y = pack_half2(fmaxf(v[8 * j + 2], 0.f), fmaxf(v[8 * j + 3], 0.f));
              o.z = pack_half2(fmaxf(v[8 * j + 4], 0.f), fmaxf(v[8 * j + 5], 0.f));
              o.w = pack_half2(fmaxf(v[8 * j + 6], 0.f), fmaxf(v[8 * j + 7], 0.f));
            }
            *reinterpret_cast<uint4*>(halo + px * 128 + (((colhalf * 4 + j) ^ (px & 7)) << 4)) = o;
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) arrive(&halo_full[hb]);
      // (3) publish the prefetched patch (tile seq+2) into the buffer tile seq used (read in iteration seq-1)
      if (tnn < total) store_patch(s_patch + hb * kPipePatchElems, pre);
      asm volatile("bar.sync 3, 256;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) {
    cluster_sync_all();   // the peer's tensor memory and barriers stay alive until the leader's last MMA has retired
    if (warp == 1) tmem_dealloc2(tmem_base, kTmemCols);
  } else {
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
  }
}

// kPair: tmB must be the weight map with kN / 2-row boxes (one CTA's half of a tap); needs n_slices == kN / 64.
template <class Epi, bool kFuse1a, bool kPair = false, int kN = 64>
int launch_conv_pipe(const CUtensorMap& tmA, const CUtensorMap& tmB, PipeParams p, const Epi& epi, int W, int H,
                     int batch, cudaStream_t stream) {
  p.tiles_w = (W + 15) / 16;
  p.tiles_h = (H + 15) / 16;
  p.batch = batch;
  if (kPair) {
    if (p.n_slices != kN / 64) {
      set_last_error("launch_conv_pipe: CTA pairs serve all %d output channels of the layer", kN);
      return SSB_ERR_INVALID;
    }
    p.n_slices = 1;
  }
  constexpr int smem_bytes = kFuse1a ? kPipeSmemBytes : kPipeSmemBytesTma;
  auto configure = [&]() -> int {
    SSB_CUDA_CHECK(cudaFuncSetAttribute(conv_pipe_kernel<Epi, kFuse1a, kPair, kN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        smem_bytes));
    return SSB_OK;
  };
  SSB_DEVICE_CONFIG((&conv_pipe_kernel<Epi, kFuse1a, kPair, kN>), 1, configure());
  const long long total = static_cast<long long>(p.tiles_w) * p.tiles_h * batch;
  int ctas = device_sm_count() / p.n_slices * p.n_slices;
  if (total * p.n_slices < ctas) ctas = static_cast<int>(total) * p.n_slices;
  if (kPair) {
    // whole pairs, and no more than can be co-resident (cached per device like the function attribute)
    static char occupancy_key;
    const int sms = device_sm_count();   // (takes the registry lock itself: not inside begin / end)
    int* slot = device_config_begin(&occupancy_key);
    if (*slot == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(static_cast<unsigned>(sms / 2 * 2));
      cfg.blockDim = dim3(kFuse1a ? kPipeThreadsFused : kPipeThreads);
      cfg.dynamicSmemBytes = smem_bytes;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, conv_pipe_kernel<Epi, kFuse1a, kPair, kN>, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = sms / 2;
      }
      *slot = n;
      if (std::getenv("SSB_DEBUG") != nullptr)
        std::fprintf(stderr, "ssb: conv_pipe pair mode (%s): %d co-resident CTA pairs on %d SMs\n", p.label, n, sms);
    }
    const int max_pairs = *slot;
    device_config_end();
    const int want = (ctas + 1) / 2;
    ctas = 2 * (want < max_pairs ? want : max_pairs);
  }
  SSB_CUDA_CHECK(launch_kernel(conv_pipe_kernel<Epi, kFuse1a, kPair, kN>, dim3(ctas),
                               dim3(kFuse1a ? kPipeThreadsFused : kPipeThreads), smem_bytes, stream, kPair ? 2 : 1, tmA, tmB, p,
                               epi));
  count_launch();
  prof_mark(stream, p.label);
  return SSB_OK;
}

}  // namespace ssb
