// The one tensor-core engine every dense contraction in this library goes through:
//   D[128 x BLOCK_N] (fp32, TMEM) = sum over k-steps  A_step[128 x 64] * B_step[BLOCK_N x 64]^T
// issued as tcgen05.mma (kind::f16, M=128) by one thread, operands staged by TMA into a multi-stage
// 128B-swizzled shared-memory ring, accumulator read back with tcgen05.ld by four epilogue warps.
//
// It is an *implicit GEMM*: the A operand is addressed through a 4-D tensor map (c, w, h, n), so a
// 3x3 convolution is nine shifted box loads (TMA zero-fills the out-of-bounds halo = conv padding)
// and a linear layer / attention product is the degenerate 1x1 case with (w = row, h = 0).  Up to two
// A sources are supported so that concat([x, msg]) never has to be materialised.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (warp w may only touch TMEM lanes 32*(w%4) .. +31).
#pragma once

#include "common.cuh"

namespace ssb {

constexpr int kCoreThreads = 192;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                  // fp16 elements per 128-byte swizzle row
constexpr int kATileBytes = kTileM * 128;    // 16 KiB

// A count that lives in device memory, looked up per blockIdx.z:
//   ptr[((z / div) ^ xr) * mul + add]   (ptr == nullptr -> unbounded)
struct DevCount {
  const int* ptr;
  int div, xr, mul, add;
  __device__ __forceinline__ int get(int z) const {
    return ptr == nullptr ? 0x7fffffff : ptr[((z / div) ^ xr) * mul + add];
  }
};
inline DevCount dev_count(const int* ptr, int div = 1, int xr = 0, int mul = 1, int add = 0) {
  DevCount d;
  d.ptr = ptr, d.div = div, d.xr = xr, d.mul = mul, d.add = add;
  return d;
}

struct CoreParams {
  // K loop: taps_h*taps_w filter taps, each contributing kc0 (+kc1) 64-wide channel chunks.
  int taps_h, taps_w, pad;
  int kc0, kc1;
  int b_tap_rows;      // row offset in B between consecutive taps
  // M tiling: blockIdx.x -> (tile_x, tile_y); the TMA box is tile_w x tile_h pixels (product 128).
  int tile_w, tile_h, tiles_w;
  // N tiling: n0 = blockIdx.y * block_n; the tile is issued as n_parts MMAs of n_part columns.
  int block_n, n_part, n_parts;
  int stages, tmem_cols;
  // batching over blockIdx.z
  int a_z_mul;         // A 4th coordinate = blockIdx.z * a_z_mul + a_z_add
  int a_z_add;
  int b_z_xor;         // B 3rd coordinate = (blockIdx.z ^ b_z_xor) * b_z_mul + b_z_add
  int b_z_mul;
  int b_z_add;
  // optional device-side extents (keypoint counts live on the device so a whole frame pair runs
  // without a host round trip): tiles whose first row >= m_valid or first column >= n_valid exit,
  // rows >= m_valid are masked by the epilogues, K chunks beyond ceil(k_valid / 64) are not issued.
  DevCount m_valid, n_valid, k_valid;
  const char* label;   // host-only: kernel name for the event profiler
};

struct EpiCtx {
  uint32_t tmem_row;   // TMEM address of this thread's accumulator row, column 0
  int row;             // 0..127 within the tile
  int lane;            // lane in warp (row & 31)
  int px, py;          // pixel (or row-index, 0) coordinates of this row: tile origin + offset
  int z;               // blockIdx.z
  int n0;              // first output column of this tile
  int m_valid;         // number of valid rows for this z (INT_MAX if unbounded)
};

__host__ __device__ inline int core_stage_bytes(int block_n) { return kATileBytes + block_n * 128; }

inline int core_smem_bytes(int block_n, int stages) {
  return stages * core_stage_bytes(block_n) + 1024 /*align slack*/ + 256 /*barriers*/;
}

template <class Epi>
__global__ void __launch_bounds__(kCoreThreads)
umma_core_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const CoreParams p, const Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int stage_bytes = core_stage_bytes(p.block_n);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* accum_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const int tile_x = blockIdx.x % p.tiles_w;
  const int tile_y = blockIdx.x / p.tiles_w;
  const int w0 = tile_x * p.tile_w;
  const int h0 = tile_y * p.tile_h;
  const int n0 = blockIdx.y * p.block_n;

  const int m_valid = p.m_valid.get(z);
  // whole tile beyond the valid rows / columns (uniform across the CTA, before any barrier exists)
  if (p.m_valid.ptr != nullptr && w0 >= m_valid) return;
  if (p.n_valid.ptr != nullptr && n0 >= p.n_valid.get(z)) return;
  int kc0 = p.kc0;
  if (p.k_valid.ptr != nullptr) kc0 = min(kc0, (p.k_valid.get(z) + kChunkK - 1) / kChunkK);
  const int kc = kc0 + p.kc1;
  const int num_k = p.taps_h * p.taps_w * kc;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
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

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = static_cast<uint32_t>(stage_bytes);
      const int az = z * p.a_z_mul + p.a_z_add;
      const int bz = (z ^ p.b_z_xor) * p.b_z_mul + p.b_z_add;
      int it = 0;
      for (int th = 0; th < p.taps_h; ++th) {
        for (int tw = 0; tw < p.taps_w; ++tw) {
          const int tap = th * p.taps_w + tw;
          for (int c = 0; c < kc; ++c, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = static_cast<uint32_t>(it / p.stages) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
            uint8_t* sa = smem + s * stage_bytes;
            uint8_t* sb = sa + kATileBytes;
            if (c < kc0) {
              tma_load_4d(sa, &tmA0, &full_bar[s], c * kChunkK, w0 + tw - p.pad, h0 + th - p.pad, az);
            } else {
              tma_load_4d(sa, &tmA1, &full_bar[s], (c - kc0) * kChunkK, w0 + tw - p.pad,
                          h0 + th - p.pad, az);
            }
            // B columns: chunk index within the tap, source-1 chunks follow the *nominal* source-0
            // chunk count so that weight matrices keep their layout when kc0 is clipped.
            const int bcol = (c < kc0 ? c : p.kc0 + (c - kc0)) * kChunkK;
            for (int part = 0; part < p.n_parts; ++part) {
              tma_load_3d(sb + part * p.n_part * 128, &tmB, &full_bar[s], bcol,
                          tap * p.b_tap_rows + n0 + part * p.n_part, bz);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(static_cast<uint32_t>(p.n_part));
      for (int it = 0; it < num_k; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = static_cast<uint32_t>(it / p.stages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint32_t sb = sa + kATileBytes;
        const uint64_t adesc = make_smem_desc_k_sw128(sa, 1024);
#pragma unroll 1
        for (int part = 0; part < p.n_parts; ++part) {
          const uint64_t bdesc = make_smem_desc_k_sw128(sb + part * p.n_part * 128, 1024);
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k) {
            // +32 bytes per K=16 slice inside the 128B swizzle row -> +2 in the (addr>>4) field
            umma_f16(tmem_base + part * p.n_part, adesc + 2 * k, bdesc + 2 * k, idesc,
                     (it | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(accum_bar);
    }
  } else {
    const int q = warp & 3;
    EpiCtx c;
    c.row = q * 32 + lane;
    c.lane = lane;
    c.px = w0 + (c.row % p.tile_w);
    c.py = h0 + (c.row / p.tile_w);
    c.z = z;
    c.n0 = n0;
    c.m_valid = m_valid;
    c.tmem_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if (num_k > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    epi(c, num_k > 0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
}

// ---- host launcher --------------------------------------------------------------------------------
inline int core_tmem_cols(int block_n) {
  int c = 32;
  while (c < block_n) c <<= 1;
  return c;
}

// Pick the deepest ring that still lets two CTAs share an SM when the tile is small.
inline int core_pick_stages(int block_n, int num_k) {
  const int sb = core_stage_bytes(block_n);
  int st = (108 * 1024) / sb;
  if (st > 4) st = 4;
  if (st < 2) st = 2;
  if (st > num_k && num_k >= 1) st = num_k < 2 ? 2 : num_k;
  return st;
}

template <class Epi>
int launch_core(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, CoreParams p,
                const Epi& epi, dim3 grid, cudaStream_t stream) {
  if (p.block_n % 16 != 0 || p.block_n > 512 || p.tile_w * p.tile_h != kTileM) {
    set_last_error("launch_core: unsupported tile (block_n=%d tile=%dx%d)", p.block_n, p.tile_w,
                   p.tile_h);
    return SSB_ERR_INVALID;
  }
  p.n_parts = (p.block_n + 255) / 256;
  p.n_part = p.block_n / p.n_parts;
  if (p.n_part % 16 != 0) {
    set_last_error("launch_core: n_part %d not a multiple of 16", p.n_part);
    return SSB_ERR_INVALID;
  }
  p.tmem_cols = core_tmem_cols(p.block_n);
  if (p.stages <= 0) p.stages = core_pick_stages(p.block_n, p.taps_h * p.taps_w * (p.kc0 + p.kc1));
  const int smem = core_smem_bytes(p.block_n, p.stages);
  static int configured_smem = 0;  // per template instantiation
  if (smem > configured_smem) {
    SSB_CUDA_CHECK(cudaFuncSetAttribute(umma_core_kernel<Epi>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  umma_core_kernel<Epi><<<grid, kCoreThreads, smem, stream>>>(a0, a1, b, p, epi);
  SSB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  prof_mark(stream, p.label);
  return SSB_OK;
}

}  // namespace ssb
