// The tensor-core engine behind every dense contraction that is not a halo-reuse convolution
// (conv_pipe.cuh, conv_stream.cuh) or attention (attention2.cuh):
//   D[128 x BLOCK_N] (fp32, TMEM) = sum over k-steps  A_step[128 x 64] * B_step[BLOCK_N x 64]^T
// issued as tcgen05.mma (kind::f16, M=128) by one thread, operands staged by TMA into a multi-stage
// 128B-swizzled shared-memory ring, accumulator read back with tcgen05.ld by eight epilogue warps.
//
// It is an *implicit GEMM*: the A operand is addressed through a 4-D tensor map (c, w, h, n), so a
// 3x3 convolution is nine shifted box loads (TMA zero-fills the out-of-bounds halo = conv padding)
// and a linear layer is the degenerate 1x1 case with (w = row, h = 0).  Up to two A sources are
// supported so that concat([x, msg]) never has to be materialised.
//
// Persistent: one CTA per SM (pair mode, CoreParams::b_rows: one CTA pair per TPC sharing a cta_group::2 MMA
// stream) walks the (x, y, z) tile space; the accumulator is double-buffered in
// TMEM (2 x BLOCK_N columns) so the epilogue of tile t overlaps the TMA loads and MMAs of tile t+1.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue; warp w may only touch TMEM lanes 32*(w%4)..+31, so two warps share each lane
// quadrant and split the accumulator columns between them.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"

namespace ssb {

constexpr int kCoreThreads = 320;
constexpr int kCoreEpiThreads = 256;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                  // fp16 elements per 128-byte swizzle row
constexpr int kATileBytes = kTileM * 128;    // 16 KiB

// A count that lives in device memory, looked up per batch index z:
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
  int a_stride;        // convolution stride (0 or 1 = dense; 2 needs an A tensor map with element strides 2)
  int kc0, kc1;
  int b_tap_rows;      // row offset in B between consecutive taps
  // M tiling: tile x -> (tile_x, tile_y); the TMA box is tile_w x tile_h pixels (product 128).
  int tile_w, tile_h, tiles_w;
  // N tiling: n0 = y * block_n; the tile is issued as n_parts MMAs of n_part columns.
  int block_n, n_part, n_parts;
  int stages, tmem_cols, tmem_bufs, buf_stride;
  int grid_x, grid_y, grid_z;   // logical tile space walked by the persistent CTAs
  // batching over z
  int a_z_mul;         // A 4th coordinate = z * a_z_mul + a_z_add
  int a_z_add;
  int b_z_xor;         // B 3rd coordinate = (z ^ b_z_xor) * b_z_mul + b_z_add
  int b_z_mul;
  int b_z_add;
  // optional device-side extents (keypoint counts live on the device so a whole frame pair runs
  // without a host round trip): tiles whose first row >= m_valid or first column >= n_valid are skipped,
  // rows >= m_valid are masked by the epilogues, K chunks beyond ceil(k_valid / 64) are not issued.
  DevCount m_valid, n_valid, k_valid;
  int stage_bufs;      // TMA-store staging depth per epilogue warp (2 unless shared memory is short)
  // Cluster mode: the kernel is launched as clusters of grid_y CTAs; the CTA with cluster rank r serves the
  // N tile y = r of the SAME sequence of M tiles as its peers, so an epilogue that needs whole-row
  // statistics (LayerNorm over all N tiles) can exchange per-row partials through distributed shared
  // memory (EpiCtx::peer_*).  Used with grid_y == 2.
  int cluster_y;
  // Pair mode (launch_core<Epi, true>): the two CTAs of a cluster take neighbouring M tiles (x even / odd) of the same
  // (y, z) and run ONE cta_group::2 MMA stream issued by the leader - each CTA stages its own 128 rows of A and HALF of
  // the B rows (b_rows = block_n / 2; the B tensor map must have boxes of that many rows).  A [128 x 64] x [256 x 64]
  // K chunk costs 48 KB of TMA writes plus 48 KB of operand reads against 512 cycles of math at 128 B/clk of shared
  // memory: the linears were shared-memory bound; a pair moves 32 + 32 KB per CTA.  Linear layers only (taps = 1,
  // tile_h = 1, an even number of M tiles, one N part).
  int b_rows;          // rows of B this CTA stages per K chunk (block_n, or block_n / 2 in pair mode)
  long long* trace;    // diagnostic (SSB_CORE_TRACE=<label>): CTA 0 time-stamps its first 32 tiles, [tile][8]
  const char* label;   // host-only: kernel name for the event profiler
};

struct EpiCtx {
  uint32_t tmem_row;   // TMEM address of this thread's accumulator row, column 0 of the tile
  int row;             // 0..127 within the tile
  int lane;            // lane in warp (row & 31)
  int px, py;          // pixel (or row-index, 0) coordinates of this row: tile origin + offset
  int z;               // batch index
  int n0;              // first output column of this tile
  int m_valid;         // number of valid rows for this z (INT_MAX if unbounded)
  int col_begin, col_end;  // accumulator columns this thread handles (the two warps of a lane quadrant
                           // split the tile's columns; functors with kSplit == false get all of them)
  int half;            // 0 or 1: which of the two warps sharing this lane quadrant
  float* xchg;         // shared scratch [2][128] floats for row reductions across the two halves
  uint8_t* stage;      // this warp's staging area for TMA stores: stage_bufs x 4 KiB, 1024-byte aligned
  uint8_t* stage_cur;  // buffer selected by the last stage_begin()
  int stage_bufs;      // 1 or 2 (double-buffered: the next block is staged while the last store drains)
  int stage_sel;
  // cluster mode (CoreParams::cluster_y): per-row exchange with the peer CTA
  int seq;             // running count of tiles this CTA has processed (both CTAs of a cluster agree)
  float* peer_slots;   // local [2 parity][128 rows][2] floats, written by the peer
  uint64_t* peer_bar;  // local [2 parity] mbarriers: one local arrival + 1 KiB of st.async bytes from the peer
};

// ---- staged output: registers -> 128B-swizzled shared memory -> one TMA store per 32-row x 64-column
// block.  A thread that owns a row writing 16-byte pieces straight to global memory touches 32
// different lines per warp instruction (measured ~1 TB/s); the bulk store writes full lines.
// All four calls are warp-collective.
__device__ __forceinline__ void stage_begin(EpiCtx& c) {   // pick a buffer whose last store has been read out
  if (c.lane == 0) {
    if (c.stage_bufs == 2) bulk_wait_read1(); else bulk_wait_read0();
  }
  __syncwarp();
  c.stage_cur = c.stage + (c.stage_bufs == 2 ? c.stage_sel * 4096 : 0);
  c.stage_sel ^= 1;
}
__device__ __forceinline__ void stage_put(const EpiCtx& c, int srow, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(c.stage_cur + srow * 128 + ((chunk ^ (srow & 7)) << 4)) = v;
}
__device__ __forceinline__ void stage_fence(const EpiCtx&) {
  fence_proxy_async_smem();
  __syncwarp();
}
__device__ __forceinline__ void stage_drain(const EpiCtx& c) {   // before the buffer / CTA goes away
  if (c.lane == 0) bulk_wait_all();
  __syncwarp();
}

// Walk kChunks consecutive 32-column accumulator chunks starting at TMEM address `t0` with the loads
// software-pipelined: the tcgen05.ld of chunk i+1 is in flight while f(i, v) works on chunk i.  (With
// "load, wait, compute" per chunk the two epilogue warps of a scheduler spent most of their time in the
// TMEM-load latency.)  f must consume v before it returns.
template <int kChunks, class F>
__device__ __forceinline__ void tmem_chunks_pipelined(uint32_t t0, F&& f) {
  float v[2][32];
  tmem_ld_32x32(t0, v[0]);
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    tmem_ld_wait();
    if (i + 1 < kChunks) tmem_ld_32x32(t0 + (i + 1) * 32, v[(i + 1) & 1]);
    f(i, v[i & 1]);
  }
}

// Optional epilogue hook: a functor that declares a nested type `Pre` and a member
//   void prefetch(const EpiCtx&, Pre&) const
// gets it called BEFORE the wait for the tile's accumulator, and receives the same object as a third argument
// of operator().  Operands that do not depend on the accumulator (rotary factors, residual rows) are then in
// flight underneath the tile's MMAs instead of exposing a global-memory round trip per 32-column chunk.
template <class E, class = void>
struct EpiPrefetch {
  static constexpr bool value = false;
  struct type {};
};
template <class E>
struct EpiPrefetch<E, std::void_t<typename E::Pre>> {
  static constexpr bool value = true;
  using type = typename E::Pre;
};

// Barrier among the 256 epilogue threads (both halves); every epilogue thread must call it.
__device__ __forceinline__ void epi_pair_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Cluster mode: all-reduce two per-row values (already reduced over this CTA's columns) with the peer CTA
// that holds the other N tile of the same rows.  Every epilogue thread calls it once per tile.
__device__ __forceinline__ void epi_cluster_sum2(const EpiCtx& c, float& a, float& b) {
  const int par = c.seq & 1;
  if (c.half == 0) {
    // this tile's phase of the local barrier completes when the peer's 128 rows x 8 bytes have landed
    if (c.row == 0) mbar_arrive_expect_tx(c.peer_bar + par, 128 * 8);
    const uint32_t peer = cluster_ctarank() ^ 1u;
    st_async_f32x2(cluster_map(smem_u32(c.peer_slots + (par * 128 + c.row) * 2), peer), a, b,
                   cluster_map(smem_u32(c.peer_bar + par), peer));
  }
  mbar_wait(c.peer_bar + par, static_cast<uint32_t>(c.seq >> 1) & 1u);
  const float2 r = *reinterpret_cast<const float2*>(c.peer_slots + (par * 128 + c.row) * 2);
  a += r.x;
  b += r.y;
}

// Row-wise all-reduce (sum) across the two column halves of a tile.
__device__ __forceinline__ float epi_pair_sum(const EpiCtx& c, float v) {
  c.xchg[c.half * 128 + c.row] = v;
  epi_pair_sync();
  const float t = c.xchg[c.row] + c.xchg[128 + c.row];
  epi_pair_sync();
  return t;
}

__host__ __device__ inline int core_stage_bytes(int b_rows) { return kATileBytes + b_rows * 128; }

constexpr int kCoreStagingBytes = 8 * 4096;   // one 4 KiB TMA-store staging buffer per epilogue warp (x stage_bufs)

constexpr int kCoreCountSlots = 256;   // device-side extents cached per batch index z (see the kernel prologue)
inline int core_smem_bytes(int block_n, int stages, int stage_bufs) {
  return stages * core_stage_bytes(block_n) + stage_bufs * kCoreStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*xchg*/ +
         2048 /*peer slots*/ + 3 * kCoreCountSlots * 4 /*extents*/;
}

template <class Epi, bool kPair = false>
__global__ void __launch_bounds__(kCoreThreads, 1)
umma_core_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const CoreParams p, const __grid_constant__ Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  // round the base up to 1 KiB with pointer arithmetic on the __shared__ array itself, so the compiler keeps
  // the shared address space (LDS/STS instead of generic LD/ST with 64-bit address math)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = core_stage_bytes(p.b_rows);
  uint8_t* ring = smem;
  uint8_t* staging = ring + p.stages * stage_bytes;    // 8 x 4 KiB, 1024-aligned (all sizes are multiples of 1 KiB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + p.stage_bufs * kCoreStagingBytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 3);
  float* xchg = reinterpret_cast<float*>(staging + p.stage_bufs * kCoreStagingBytes + 256);
  float* peer_slots = xchg + 256;                        // [2][128][2]
  uint64_t* peer_bar = reinterpret_cast<uint64_t*>(tmem_slot + 2);   // [2]
  int* s_ext = reinterpret_cast<int*>(peer_slots + 512);   // [3][kCoreCountSlots]: m_valid, n_valid, k_valid per z

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = kPair ? static_cast<int>(cluster_ctarank()) : 0;   // pair mode: 0 = leader (issues the MMAs)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], (kCoreEpiThreads / 32) * (kPair ? 2 : 1));   // one arrival per epilogue warp (of both CTAs)
      mbar_init(&peer_bar[b], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) {
      tmem_alloc2(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  const bool clustered = kPair || p.cluster_y != 0;
  if (clustered) cluster_sync_all();   // the peers' barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                // the previous kernel's results (PDL: everything above ran under its tail)
  pdl_launch_dependents();
  // The device-side extents (keypoint counts) of every batch index are read ONCE into shared memory.  Looked up in
  // global memory by decode(), each tile cost every role one to three dependent L2 round trips before it could
  // start: the tile trace showed 1 000 - 1 700 cycles between the end of one epilogue and the start of the next in
  // kernels whose epilogue sets the pace (qkv, ffn1).
  // (Only when a CTA walks several tiles: a latency-bound call - one tile per CTA at most - is better off with the one
  // dependent load than with a block-wide copy and barrier behind the dependency wait of every kernel.)
  const long long all_tiles = static_cast<long long>(p.grid_x) * p.grid_y * p.grid_z;
  const bool ext_cached = p.grid_z <= kCoreCountSlots && all_tiles > 2LL * static_cast<long long>(gridDim.x);
  if (ext_cached) {
    for (int i = threadIdx.x; i < p.grid_z; i += blockDim.x) {
      s_ext[i] = p.m_valid.get(i);
      s_ext[kCoreCountSlots + i] = p.n_valid.get(i);
      s_ext[2 * kCoreCountSlots + i] = p.k_valid.get(i);
    }
    __syncthreads();
  }
  // tile walk: all (x, y, z) tiles strided over the CTAs, or - cluster mode - a fixed y per CTA (= its cluster rank)
  // pair mode: the tile index runs over PAIRS of M tiles (both CTAs walk the same sequence); this CTA's tile is
  // x = 2 * (pair index) + rank, and a pair is skipped as a whole when the leader's tile lies outside the extents
  const bool fixed_y = p.cluster_y != 0;
  const int ny = kPair ? 2 : (fixed_y ? p.grid_y : 1);
  const int y_fixed = static_cast<int>(blockIdx.x) % ny;
  const int first = static_cast<int>(blockIdx.x) / ny;
  const int stride = static_cast<int>(gridDim.x) / ny;
  const int gx = kPair ? p.grid_x / 2 : p.grid_x;
  const int total = fixed_y ? gx * p.grid_z : gx * p.grid_y * p.grid_z;

  // The walk keeps (x, y, z) of the current tile and advances them by the decomposed stride: decoding the linear index
  // took five integer divisions by run-time values per tile and role (plus two for the pixel of a row) - some 800 cycles
  // in front of every epilogue, in kernels whose epilogue sets the pace.
  struct Walk {
    int tile, x, y, z;
  };
  const int gy_eff = fixed_y ? 1 : p.grid_y;
  const int walk_sx = stride % gx, walk_sy = (stride / gx) % gy_eff, walk_sz = (stride / gx) / gy_eff;
  auto walk_begin = [&]() {
    Walk w;
    w.tile = first;
    w.x = first % gx;
    w.y = (first / gx) % gy_eff;
    w.z = (first / gx) / gy_eff;
    return w;
  };
  auto walk_next = [&](Walk& w) {
    w.tile += stride;
    w.x += walk_sx;
    w.y += walk_sy;
    w.z += walk_sz;
    if (w.x >= gx) {
      w.x -= gx;
      ++w.y;
    }
    if (w.y >= gy_eff) {
      w.y -= gy_eff;
      ++w.z;
    }
  };
  const bool linear_rows = p.tiles_w == p.grid_x;   // linear layers: one row of M tiles per batch index

  // Decode a tile; returns false for tiles that lie entirely outside the device-side extents.
  auto decode = [&](const Walk& wk, int& z, int& w0, int& h0, int& n0, int& m_valid, int& kc0) -> bool {
    const int x = kPair ? 2 * wk.x + rank : wk.x;
    const int y = fixed_y ? y_fixed : wk.y;
    z = wk.z;
    w0 = linear_rows ? x * p.tile_w : (x % p.tiles_w) * p.tile_w;
    h0 = linear_rows ? 0 : (x / p.tiles_w) * p.tile_h;
    n0 = y * p.block_n;
    m_valid = ext_cached ? s_ext[z] : p.m_valid.get(z);
    if (p.m_valid.ptr != nullptr && (kPair ? w0 - rank * p.tile_w : w0) >= m_valid) return false;
    if (p.n_valid.ptr != nullptr && n0 >= (ext_cached ? s_ext[kCoreCountSlots + z] : p.n_valid.get(z))) return false;
    kc0 = p.kc0;
    if (p.k_valid.ptr != nullptr)
      kc0 = min(kc0, ((ext_cached ? s_ext[2 * kCoreCountSlots + z] : p.k_valid.get(z)) + kChunkK - 1) / kChunkK);
    return true;
  };

  if (warp == 0) {
    // producer: whole warp in uniform control flow, one elected lane issues the TMA loads (see warp 1)
    const uint32_t tx_bytes = static_cast<uint32_t>(kPair ? 2 * stage_bytes : stage_bytes);   // pair: both CTAs' boxes
    int it = 0;
    for (Walk wk = walk_begin(); wk.tile < total; walk_next(wk)) {
      int z, w0, h0, n0, m_valid, kc0;
      if (!decode(wk, z, w0, h0, n0, m_valid, kc0)) continue;
      const int kc = kc0 + p.kc1;
      const int az = z * p.a_z_mul + p.a_z_add;
      const int bz = (z ^ p.b_z_xor) * p.b_z_mul + p.b_z_add;
      if (p.a_stride > 1) {   // input coordinates of the tile origin
        w0 *= p.a_stride;
        h0 *= p.a_stride;
      }
      for (int th = 0; th < p.taps_h; ++th) {
        for (int tw = 0; tw < p.taps_w; ++tw) {
          const int tap = th * p.taps_w + tw;
          for (int c = 0; c < kc; ++c, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = static_cast<uint32_t>(it / p.stages) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* sa = ring + s * stage_bytes;
            uint8_t* sb = sa + kATileBytes;
            // B columns: source-1 chunks follow the *nominal* source-0 chunk count so that weight
            // matrices keep their layout when kc0 is clipped by k_valid.
            const int bcol = (c < kc0 ? c : p.kc0 + (c - kc0)) * kChunkK;
            if (elect_one()) {
              if (kPair) {   // the bytes of both CTAs are counted on the leader's barrier
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                tma_load_4d_pair(sa, c < kc0 ? &tmA0 : &tmA1, &full_bar[s], (c < kc0 ? c : c - kc0) * kChunkK, w0, h0, az);
                tma_load_3d_pair(sb, &tmB, &full_bar[s], bcol, n0 + rank * p.b_rows, bz);
              } else {
                mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                if (c < kc0) {
                  tma_load_4d(sa, &tmA0, &full_bar[s], c * kChunkK, w0 + tw - p.pad, h0 + th - p.pad, az);
                } else {
                  tma_load_4d(sa, &tmA1, &full_bar[s], (c - kc0) * kChunkK, w0 + tw - p.pad, h0 + th - p.pad, az);
                }
                for (int part = 0; part < p.n_parts; ++part) {
                  tma_load_3d(sb + part * p.n_part * 128, &tmB, &full_bar[s], bcol,
                              tap * p.b_tap_rows + n0 + part * p.n_part, bz);
                }
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // Whole-warp, warp-uniform loop; one elected lane issues.  Keeping descriptor arithmetic out of a
    // divergent `if (lane == 0)` region lets it live in uniform registers, so a tcgen05.mma is a single
    // instruction instead of an ELECT / R2UR / branch sequence of ~100 cycles.
    const uint32_t idesc = kPair ? make_idesc2_f16(static_cast<uint32_t>(p.n_part)) : make_idesc_f16(static_cast<uint32_t>(p.n_part));
    auto wait = [&](uint64_t* bar, uint32_t parity) {   // barriers the peer's agents complete: cluster-scope acquire
      if (kPair) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity);
    };
    const uint32_t ring_base = smem_u32(ring);
    int it = 0, seq = 0;
    for (Walk wk = walk_begin(); wk.tile < total; walk_next(wk)) {
      int z, w0, h0, n0, m_valid, kc0;
      if (!decode(wk, z, w0, h0, n0, m_valid, kc0)) continue;
      const int num_k = p.taps_h * p.taps_w * (kc0 + p.kc1);
      const int buf = p.tmem_bufs == 2 ? (seq & 1) : 0;
      const uint32_t use = static_cast<uint32_t>(p.tmem_bufs == 2 ? (seq >> 1) : seq);
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && seq < 32;
      if (tr) p.trace[seq * 8 + 0] = clock64();
      wait(&tmem_empty[buf], (use & 1u) ^ 1u);   // epilogue has drained this accumulator
      tc_fence_after();
      if (tr) p.trace[seq * 8 + 1] = clock64();
      const uint32_t d_tmem = tmem_base + buf * p.buf_stride;
      for (int kk = 0; kk < num_k; ++kk, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = static_cast<uint32_t>(it / p.stages) & 1u;
        wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = ring_base + s * stage_bytes;
        const uint32_t sb = sa + kATileBytes;
        const uint64_t adesc = make_smem_desc_k_sw128(sa, 1024);
        if (elect_one()) {
#pragma unroll 1
          for (int part = 0; part < p.n_parts; ++part) {
            const uint64_t bdesc = make_smem_desc_k_sw128(sb + part * p.n_part * 128, 1024);
#pragma unroll
            for (int k = 0; k < kChunkK / 16; ++k) {
              // +32 bytes per K=16 slice inside the 128B swizzle row -> +2 in the (addr>>4) field
              if (kPair) umma2_f16(d_tmem + part * p.n_part, adesc + 2 * k, bdesc + 2 * k, idesc, (kk | k) != 0 ? 1u : 0u);
              else umma_f16(d_tmem + part * p.n_part, adesc + 2 * k, bdesc + 2 * k, idesc, (kk | k) != 0 ? 1u : 0u);
            }
          }
          if (kPair) umma2_commit(&empty_bar[s]); else umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (elect_one()) {
        if (kPair) umma2_commit(&tmem_full[buf]); else umma_commit(&tmem_full[buf]);
      }
      __syncwarp();
      if (tr) p.trace[seq * 8 + 2] = clock64();
      ++seq;
    }
  } else if (warp == 1) {
    // pair mode: the peer's MMA warp has nothing to issue
  } else {
    const int ew = warp - 2;          // 0..7
    const int q = warp & 3;           // TMEM lane quadrant this warp may access
    const int half = ew >> 2;         // two warps per quadrant
    int seq = 0;
    EpiCtx c;
    c.stage = staging + ew * (p.stage_bufs * 4096);
    c.stage_cur = c.stage;
    c.stage_bufs = p.stage_bufs;
    c.stage_sel = 0;
    const int row_w = (q * 32 + lane) % p.tile_w, row_h = (q * 32 + lane) / p.tile_w;   // pixel of this thread's row in a tile
    for (Walk wk = walk_begin(); wk.tile < total; walk_next(wk)) {
      int z, w0, h0, n0, m_valid, kc0;
      if (!decode(wk, z, w0, h0, n0, m_valid, kc0)) continue;
      const int num_k = p.taps_h * p.taps_w * (kc0 + p.kc1);
      const int buf = p.tmem_bufs == 2 ? (seq & 1) : 0;
      const uint32_t use = static_cast<uint32_t>(p.tmem_bufs == 2 ? (seq >> 1) : seq);
      c.row = q * 32 + lane;
      c.lane = lane;
      c.px = w0 + row_w;
      c.py = h0 + row_h;
      c.z = z;
      c.n0 = n0;
      c.m_valid = m_valid;
      c.half = half;
      c.xchg = xchg;
      c.seq = seq;
      c.peer_slots = peer_slots;
      c.peer_bar = peer_bar;
      c.tmem_row = tmem_base + buf * p.buf_stride + (static_cast<uint32_t>(q * 32) << 16);
      const bool mine = Epi::kSplit || half == 0;
      if (Epi::kSplit) {
        const int hw = p.block_n / 2;
        c.col_begin = half * hw;
        c.col_end = c.col_begin + hw;
      } else {
        c.col_begin = 0;
        c.col_end = p.block_n;
      }
      typename EpiPrefetch<Epi>::type pre;
      if constexpr (EpiPrefetch<Epi>::value) {
        if (mine) epi.prefetch(c, pre);
      }
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0 && seq < 32;
      if (tr) p.trace[seq * 8 + 3] = clock64();
      mbar_wait(&tmem_full[buf], use & 1u);
      tc_fence_after();
      if (tr) p.trace[seq * 8 + 4] = clock64();
      if (mine) {
        if constexpr (EpiPrefetch<Epi>::value) epi(c, num_k > 0, pre);
        else epi(c, num_k > 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_remote(&tmem_empty[buf], 0); else mbar_arrive(&tmem_empty[buf]);
      }
      if (tr) p.trace[seq * 8 + 5] = clock64();
      ++seq;
    }
    if (lane == 0) bulk_wait_all();   // outstanding TMA stores still read this CTA's shared memory
  }

  tc_fence_before();
  __syncthreads();
  if (clustered) cluster_sync_all();   // a peer may still be writing this CTA's exchange slots / operand stages
  if (warp == 1) {
    if (kPair) tmem_dealloc2(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    else tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ---- host launcher --------------------------------------------------------------------------------
inline int core_tmem_cols(int cols) {
  int c = 32;
  while (c < cols) c <<= 1;
  return c;
}

// `grid` is the logical tile space (x = M tiles, y = N tiles, z = batch); the launch itself uses one
// persistent CTA per SM.
// kPair: CTA pairs (CoreParams::b_rows): `b` must be the B map with block_n / 2-row boxes.
template <class Epi, bool kPair = false>
int launch_core(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, CoreParams p,
                const Epi& epi, dim3 grid, cudaStream_t stream) {
  if (p.block_n % 16 != 0 || p.block_n > 512 || p.tile_w * p.tile_h != kTileM) {
    set_last_error("launch_core: unsupported tile (block_n=%d tile=%dx%d)", p.block_n, p.tile_w, p.tile_h);
    return SSB_ERR_INVALID;
  }
  p.n_parts = (p.block_n + 255) / 256;
  p.n_part = p.block_n / p.n_parts;
  p.b_rows = kPair ? p.block_n / 2 : p.block_n;
  if (kPair && (p.taps_h * p.taps_w != 1 || p.tile_h != 1 || grid.x % 2 != 0 || p.n_parts != 1 || p.cluster_y != 0 ||
                p.block_n % 32 != 0)) {
    set_last_error("launch_core: pair mode needs a linear layer with an even number of M tiles (block_n=%d grid.x=%u)", p.block_n,
                   grid.x);
    return SSB_ERR_INVALID;
  }
  if (p.n_part % 16 != 0 || (Epi::kSplit && (p.block_n / 2) % 32 != 0)) {
    set_last_error("launch_core: block_n %d not splittable for this epilogue", p.block_n);
    return SSB_ERR_INVALID;
  }
  p.buf_stride = core_tmem_cols(p.block_n);
  p.tmem_bufs = (2 * p.buf_stride <= 512) ? 2 : 1;
  p.tmem_cols = p.buf_stride * p.tmem_bufs;
  p.grid_x = static_cast<int>(grid.x);
  p.grid_y = static_cast<int>(grid.y);
  p.grid_z = static_cast<int>(grid.z);
  p.stage_bufs = 2;
  if (p.stages <= 0) {
    // 222 KiB budget: ring as deep as fits next to a double-buffered staging area, at least 2 stages;
    // very wide tiles (block_n 512) fall back to single-buffered staging
    const int sb = core_stage_bytes(p.b_rows);
    int st = (222 * 1024 - 2 * kCoreStagingBytes - 4096) / sb;
    if (st < 2) {
      p.stage_bufs = 1;
      st = (222 * 1024 - kCoreStagingBytes - 4096) / sb;
    }
    p.stages = st > 6 ? 6 : (st < 2 ? 2 : st);
  }
  const int smem = core_smem_bytes(p.b_rows, p.stages, p.stage_bufs);
  auto configure = [&]() -> int {   // per device and template instantiation (common.cuh)
    SSB_CUDA_CHECK(cudaFuncSetAttribute(umma_core_kernel<Epi, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return SSB_OK;
  };
  SSB_DEVICE_CONFIG((&umma_core_kernel<Epi, kPair>), smem, configure());
  const long long total = static_cast<long long>(grid.x) * grid.y * grid.z;
  if (total <= 0) return SSB_OK;
  // diagnostic: SSB_CORE_TRACE=<label> dumps CTA 0's per-tile time stamps of the first launch with that label
  static long long* trace_dev = nullptr;
  static int trace_state = 0;   // 0 = unknown, 1 = armed, 2 = done / off
  bool tracing = false;
  if (trace_state != 2) {
    const char* want = std::getenv("SSB_CORE_TRACE");
    if (want == nullptr) {
      trace_state = 2;
    } else if (p.label != nullptr && std::strcmp(want, p.label) == 0) {
      if (trace_dev == nullptr) cudaMalloc(&trace_dev, 32 * 8 * sizeof(long long));
      cudaMemset(trace_dev, 0, 32 * 8 * sizeof(long long));
      p.trace = trace_dev;
      tracing = true;
      trace_state = 2;
    }
  }
  int ctas = static_cast<int>(total < device_sm_count() ? total : device_sm_count());
  if (p.cluster_y || kPair) {
    if (!kPair && grid.y != 2) {
      set_last_error("launch_core: cluster mode needs exactly two N tiles");
      return SSB_ERR_INVALID;
    }
    const int csize = 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(device_sm_count() / csize * csize));
    cfg.blockDim = dim3(kCoreThreads);
    cfg.dynamicSmemBytes = static_cast<size_t>(smem);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(csize);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent pairs: as many clusters as can be co-resident (a GPC with an odd number of free SMs
    // cannot host a pair on its last SM), so that no cluster waits for a second wave
    // (cached per device and shared-memory size: the registry slot holds smem << 12 | clusters)
    static char occupancy_key;   // one per template instantiation
    int max_clusters = 0;
    const int sms = device_sm_count();
    {
      int* slot = device_config_begin(&occupancy_key);
      if ((*slot >> 12) != smem) {
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, umma_core_kernel<Epi, kPair>, &cfg) != cudaSuccess || n <= 0) {
          cudaGetLastError();
          n = sms / csize;
        }
        *slot = (smem << 12) | n;
      }
      max_clusters = *slot & 0xfff;
      device_config_end();
    }
    // groups of tiles walked by one cluster: M tiles x batch (a cluster spans the N tiles); pair mode: pairs of M
    // tiles x N tiles x batch
    const long long groups = kPair ? static_cast<long long>(grid.x / 2) * grid.y * grid.z : static_cast<long long>(grid.x) * grid.z;
    ctas = static_cast<int>(groups < max_clusters ? groups : max_clusters) * csize;
    SSB_CUDA_CHECK(launch_kernel(umma_core_kernel<Epi, kPair>, dim3(static_cast<unsigned>(ctas)), dim3(kCoreThreads), smem, stream,
                                 static_cast<unsigned>(csize), a0, a1, b, p, epi));
  } else {
    SSB_CUDA_CHECK(launch_kernel(umma_core_kernel<Epi, kPair>, dim3(static_cast<unsigned>(ctas)), dim3(kCoreThreads), smem, stream, 1,
                                 a0, a1, b, p, epi));
  }
  if (tracing) {
    SSB_CUDA_CHECK(cudaStreamSynchronize(stream));
    long long h[32][8];
    SSB_CUDA_CHECK(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
    std::fprintf(stderr, "core trace (%s, %d CTAs, %d stages): tile | mma: wait_empty issue | epi: wait_full epilogue | "
                         "epi period, mma period\n", p.label, ctas, p.stages);
    for (int t = 0; t + 1 < 32 && h[t + 1][3] != 0; ++t)
      std::fprintf(stderr, "%2d | %6lld %6lld | %6lld %6lld | %6lld %6lld\n", t, h[t][1] - h[t][0], h[t][2] - h[t][1],
                   h[t][4] - h[t][3], h[t][5] - h[t][4], h[t + 1][3] - h[t][3], h[t + 1][0] - h[t][0]);
  }
  count_launch();
  prof_mark(stream, p.label);
  return SSB_OK;
}

}  // namespace ssb
