// Persistent, warp-specialised 3x3 convolution for the Cin = 128 layers of SuperPoint (conv3b, conv4a,
// conv4b, convPa|convDa): the same halo-reuse scheme as conv_pipe.cuh - one (16+2) x (16+2) pixel halo
// per 16 x 16 output tile, all nine taps read it in place through UMMA descriptors shifted by whole
// pixels - but with 128 output channels per CTA and the weights STREAMED through a shared-memory ring
// (9 taps x 2 channel slabs x [128 x 64] fp16 = 288 KB per slice do not fit next to the halos).
//
//   halo     two 64-channel slabs per tile; THREE slab buffers rotate (slab loads are numbered
//            L = 2*tile + slab, buffer L % 3), so the load of a slab has half a tile of MMAs to land
//   weights  ring of 16 KB stages, one (slab, tap) slice each, by a dedicated TMA warp
//   MMA      per slab and tap: 2 sub-tiles x 4 K-slices of tcgen05.mma (M = 128, N = 128)
//   TMEM     2 x (2 x 128) accumulator columns: the epilogue of tile t overlaps the MMAs of tile t+1
// Warp roles (352 threads): 0 = weight TMA, 1 = TMEM alloc + MMA issue, 2..9 = epilogue (two per TMEM lane
// quadrant, one per sub-tile), 10 = halo TMA.  The two producers are separate warps because each blocks on
// a different consumer event (ring slot free / halo buffer free); one in-order thread doing both would
// stall the weight stream whenever it waits for a halo buffer.
#pragma once

#include "common.cuh"
#include "umma_core.cuh"

namespace ssb {

constexpr int kStreamHaloBytes = 41984;                 // 18*18*128 = 41472, rounded up to 1 KiB
constexpr int kStreamHaloBufs = 3;
constexpr int kStreamWStages = 4;
constexpr int kStreamWStageBytes = 128 * 128;           // [128 cout x 64 cin] fp16
constexpr int kStreamEpiWarps = 8;
constexpr int kStreamThreads = 64 + kStreamEpiWarps * 32 + 32;
constexpr int kStreamSmemBytes = kStreamHaloBufs * kStreamHaloBytes + kStreamWStages * kStreamWStageBytes +
                                 kStreamEpiWarps * 4096 + 256 + 1024;
static_assert(kStreamSmemBytes <= 227 * 1024, "conv_stream shared-memory budget");

struct StreamParams {
  int tiles_w, tiles_h, batch;   // 16x16-pixel tiles per image
  int n_slices;                  // Cout / 128; CTA c serves slice c % n_slices
  int cout_rows;                 // rows per tap in the weight matrix
  const char* label;
};

// tmA: 4-D (C = 128, W, H, B) box (64, 18, 18, 1).  tmB: 3-D (Cin, taps*cout_rows, 1) box (64, 128, 1).
template <class Epi>
__global__ void __launch_bounds__(kStreamThreads, 1)
conv_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const StreamParams p, const __grid_constant__ Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_halo = smem;                                             // [3]
  uint8_t* s_w = s_halo + kStreamHaloBufs * kStreamHaloBytes;         // [4]
  uint8_t* s_stage = s_w + kStreamWStages * kStreamWStageBytes;       // 8 x 4 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + kStreamEpiWarps * 4096);
  uint64_t* halo_full = bars;          // [3]
  uint64_t* halo_empty = bars + 3;     // [3]
  uint64_t* w_full = bars + 6;         // [4]
  uint64_t* w_empty = bars + 10;       // [4]
  uint64_t* tmem_full = bars + 14;     // [2]
  uint64_t* tmem_empty = bars + 16;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.n_slices;
  const int first = blockIdx.x / p.n_slices;
  const int stride = gridDim.x / p.n_slices;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int total = tiles_per_img * p.batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int b = 0; b < kStreamHaloBufs; ++b) {
      mbar_init(&halo_full[b], 1);
      mbar_init(&halo_empty[b], 1);
    }
    for (int s = 0; s < kStreamWStages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], kStreamEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                // the previous kernel's results (PDL: everything above ran under its tail)
  pdl_launch_dependents();

  if (warp == 0) {
    // weights: one (slab, tap) slice per ring stage, in the order the MMAs consume them
    int it = 0;
    for (int t = first; t < total; t += stride) {
      for (int sl = 0; sl < 2; ++sl) {
        for (int tap = 0; tap < 9; ++tap, ++it) {
          const int st = it % kStreamWStages;
          mbar_wait(&w_empty[st], (static_cast<uint32_t>(it / kStreamWStages) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&w_full[st], kStreamWStageBytes);
            tma_load_3d(s_w + st * kStreamWStageBytes, &tmB, &w_full[st], sl * 64, tap * p.cout_rows + slice * 128, 0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 10) {
    // halo slabs
    int ld = 0;
    for (int t = first; t < total; t += stride) {
      const int z = t / tiles_per_img, r = t % tiles_per_img;
      const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
      for (int sl = 0; sl < 2; ++sl, ++ld) {
        const int hb = ld % kStreamHaloBufs;
        mbar_wait(&halo_empty[hb], (static_cast<uint32_t>(ld / kStreamHaloBufs) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&halo_full[hb], 18 * 18 * 128);
          tma_load_4d(s_halo + hb * kStreamHaloBytes, &tmA, &halo_full[hb], sl * 64, w0 - 1, h0 - 1, z);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // whole warp in uniform control flow, one elected lane issues (descriptors stay in uniform registers)
    const uint32_t idesc = make_idesc_f16(128);
    const uint32_t w_base = smem_u32(s_w), halo_base = smem_u32(s_halo);
    int seq = 0, it = 0, ld = 0;
    for (int t = first; t < total; t += stride, ++seq) {
      const int b = seq & 1;
      mbar_wait(&tmem_empty[b], (static_cast<uint32_t>(seq >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + b * 256;
      for (int sl = 0; sl < 2; ++sl, ++ld) {
        const int hb = ld % kStreamHaloBufs;
        mbar_wait(&halo_full[hb], static_cast<uint32_t>(ld / kStreamHaloBufs) & 1u);
        const uint32_t hbase = halo_base + hb * kStreamHaloBytes;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, ++it) {
          const int st = it % kStreamWStages;
          mbar_wait(&w_full[st], static_cast<uint32_t>(it / kStreamWStages) & 1u);
          tc_fence_after();
          const int kh = tap / 3, kw = tap - kh * 3;
          const uint64_t bdesc = make_smem_desc_k_sw128(w_base + st * kStreamWStageBytes, 1024);
          const uint32_t a_tap = hbase + (kh * 18 + kw) * 128;
          if (elect_one()) {
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const uint64_t adesc = make_smem_desc_k_sw128(a_tap + sub * 8 * 128, 18 * 128);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(d0 + sub * 128, adesc + 2 * k, bdesc + 2 * k, idesc, (sl | tap | k) != 0 ? 1u : 0u);
            }
            umma_commit(&w_empty[st]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&halo_empty[hb]);
        __syncwarp();
      }
      if (elect_one()) umma_commit(&tmem_full[b]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    int seq = 0;
    EpiCtx c;
    c.row = q * 32 + lane;
    c.lane = lane;
    c.n0 = slice * 128;
    c.m_valid = 0x7fffffff;
    c.col_begin = 0;
    c.col_end = 128;
    c.half = 0;
    c.xchg = nullptr;
    c.stage = s_stage + (warp - 2) * 4096;
    c.stage_cur = c.stage;
    c.stage_bufs = 1;
    c.stage_sel = 0;
    c.seq = 0;
    c.peer_slots = nullptr;
    c.peer_bar = nullptr;
    for (int t = first; t < total; t += stride, ++seq) {
      const int b = seq & 1;
      const int z = t / tiles_per_img, r = t % tiles_per_img;
      const int w0 = (r % p.tiles_w) * 16, h0 = (r / p.tiles_w) * 16;
      mbar_wait(&tmem_full[b], static_cast<uint32_t>(seq >> 1) & 1u);
      tc_fence_after();
      c.z = z;
      c.py = h0 + (c.row >> 3);
      c.px = w0 + sub * 8 + (c.row & 7);
      c.tmem_row = tmem_base + b * 256 + sub * 128 + (static_cast<uint32_t>(q * 32) << 16);
      epi(c, true);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[b]);
    }
    stage_drain(c);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <class Epi>
int launch_conv_stream(const CUtensorMap& tmA, const CUtensorMap& tmB, StreamParams p, const Epi& epi, int W, int H,
                       int batch, cudaStream_t stream) {
  p.tiles_w = (W + 15) / 16;
  p.tiles_h = (H + 15) / 16;
  p.batch = batch;
  auto configure = [&]() -> int {
    SSB_CUDA_CHECK(cudaFuncSetAttribute(conv_stream_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kStreamSmemBytes));
    return SSB_OK;
  };
  SSB_DEVICE_CONFIG((&conv_stream_kernel<Epi>), 1, configure());
  const long long total = static_cast<long long>(p.tiles_w) * p.tiles_h * batch;
  if (total <= 0) return SSB_OK;
  int ctas = device_sm_count() / p.n_slices * p.n_slices;
  if (total * p.n_slices < ctas) ctas = static_cast<int>(total) * p.n_slices;
  SSB_CUDA_CHECK(launch_kernel(conv_stream_kernel<Epi>, dim3(ctas), dim3(kStreamThreads), kStreamSmemBytes, stream, 1, tmA,
                               tmB, p, epi));
  count_launch();
  prof_mark(stream, p.label);
  return SSB_OK;
}

}  // namespace ssb
