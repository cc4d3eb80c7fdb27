// Fused attention for LightGlue on sm_100a (flash-attention style, persistent CTAs over 128-query tiles):
//   S = Q K^T        tcgen05.mma, fp32 accumulator in TMEM (never leaves the SM)
//   P = exp2(S*c - m) online softmax by 256 threads (two per query row, 64 keys each), fp16 P written
//                    back to TMEM with tcgen05.st (two fp16 per 32-bit cell, lane = query row)
//   O += P V         tcgen05.mma with P as a TMEM A operand and V as an MN-major shared-memory B operand
//                    (V rows = keys, as stored by the QKV epilogue; no transposed copy of V exists)
// The N x M logits of the reference graph (softmax(q k^T / 8) v for self attention, both directions of
// the bidirectional cross attention: oracle/lightglue.py _self_block/_cross_block) are therefore never
// written to HBM.  The running max is only refreshed when it grows by more than 2^8 (the O rescale is
// skipped otherwise), which keeps P <= 256 in fp16 and is exact after the final division by the row sum.
//
// Keeping P in tensor memory takes 64 KB per key block off the shared-memory port (32 KB of stores + 32 KB
// of operand reads; the N = 64 P*V MMAs were shared-memory-read bound) and frees room for a second V stage.
// Budget: 101 KB of shared memory (Q 16, K 2 x 16, V 2 x 16, O staging 16) and 256 TMEM columns (S 128, O 64,
// P 64) per CTA -> two CTAs per SM, so one CTA's
// exponentials (the MUFU-bound part: 16 ex2/clk/SM) overlap the other's MMAs and loads.  (A variant that
// pipelines S/P double-buffered inside one CTA per SM measured 25 % slower: profiles/README.md.)
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = softmax + epilogue (TMEM lane quadrant = warp % 4, key half = (warp-2) / 4).
#pragma once

#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "umma_core.cuh"

namespace ssb {

constexpr int kFaThreads = 320;
constexpr int kFaBlockKeys = 128;
constexpr int kFaSmemBytes = 16384 /*Q*/ + 2 * 16384 /*K*/ + 2 * 16384 /*V*/ + 16384 /*O staging*/ +
                             4096 /*xchg*/ + 256 /*barriers*/ + 1024 /*align*/;

struct FaParams {
  const int* cnt;      // per-image keypoint counts
  int heads;           // z = img * heads + head
  int key_xor;         // keys / values come from image (img ^ key_xor)
  float scale_log2;    // logits scale * log2(e)
  __half* ctx;         // [img][kp][heads*64]
  int kp;
  int q_tiles, zcount; // tile space (set by the launcher)
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
      "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16_u32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand is read from tensor memory (lane = row, two fp16 per cell)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32x2 arithmetic (sm_100: one FMA-pipe instruction for two lanes of work)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  uint64_t ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  uint64_t ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
// Barrier between the two warps (w, w + 4) that share a TMEM lane quadrant, i.e. the two halves of 32 rows.
// (A CTA-wide barrier here made every block wait for the slowest of eight warps spread over four schedulers.)
// 2^x for a pair on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, relative error 7.5e-5, i.e.
// six times below the fp16 rounding of P): n = round(x) through the 1.5 * 2^23 trick, f = x - n in [-0.5, 0.5],
// 2^f ~ c0 + f (c1 + f (c2 + f c3)), and the exponent is added as an integer: bits(t) << 23 == n << 23 (mod 2^32).
// One pair in kPolyEvery (template parameter of the kernel) goes this way to take load off the MUFU
// (16 ex2/clk/SM), the bound of this kernel.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 t = fadd2(x, magic);
  const float2 n = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = ffma2(n, make_float2(-1.0f, -1.0f), x);
  float2 p = ffma2(make_float2(0.05517146f, 0.05517146f), f, make_float2(0.24261086f, 0.24261086f));
  p = ffma2(p, f, make_float2(0.69326099f, 0.69326099f));
  p = ffma2(p, f, make_float2(0.99992809f, 0.99992809f));
  float2 r;
  r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
  r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
  return r;
}
__device__ __forceinline__ void fa_pair_sync(int qd) {
  asm volatile("bar.sync %0, 64;" ::"r"(2 + qd) : "memory");
}
// Diagnostic build of the kernel (SSB_FA_TRACE=1): warp 2 of CTA 0 time-stamps the phases of its first key
// blocks; the launcher prints them after the first launch.  Not instantiated on the product path.
__device__ long long g_fa_trace[64][8];

#define FA_TRACE(i)                                                                              \
  do {                                                                                           \
    if (kTrace && blockIdx.x == 0 && warp == 2 && lane == 0 && kb < 64) g_fa_trace[kb][i] = clock64(); \
  } while (0)

// One unit of work: the 128-query tile `qt` of (image, head) `z`.
struct FaTile {
  int z, img, q0, nq, nk, zk, nblk;
};
__device__ __forceinline__ bool fa_decode(const FaParams& p, int tile, FaTile& t) {
  t.z = tile / p.q_tiles;
  t.q0 = (tile - t.z * p.q_tiles) * 128;
  t.img = t.z / p.heads;
  t.nq = p.cnt[t.img];
  if (t.q0 >= t.nq) return false;   // no queries: nothing to compute, nothing reads these context rows
  t.nk = p.cnt[t.img ^ p.key_xor];
  t.zk = (t.img ^ p.key_xor) * p.heads + (t.z - t.img * p.heads);
  t.nblk = (t.nk + kFaBlockKeys - 1) / kFaBlockKeys;
  return true;
}

// tmQ: 4-D (64, kp, 1, Z) box (64,128,1,1).  tmK, tmV: 3-D (64, kp, Z) box (64,128,1).
// Persistent: two CTAs per SM walk the (z, query tile) space with a stride of gridDim.x.  All mbarrier
// phases are driven by running counters (kb = key blocks processed by this CTA, tq = tiles), so the
// producer runs ahead across tile boundaries: Q of the next tile is loaded as soon as the last Q K^T of
// the current one has retired, K/V blocks keep streaming through their rings, and the first S of the next
// tile is computed underneath the last softmax of the current one.  (As one CTA per tile, ~1/3 of each
// CTA's life went into barrier/TMEM setup and the serial Q -> K -> S -> softmax start-up latency.)
template <bool kTrace, int kPolyEvery>
__global__ void __launch_bounds__(kFaThreads, 2)
flash_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                       const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;            // 2 stages
  constexpr uint32_t kVS = 2u;           // V stages
  uint8_t* sV = smem + 16384 + 32768;    // 2 stages
  // [2 parity][2 halves][128] block maxima: softmax(b+1) may start (its S is computed underneath softmax(b))
  // before the partner thread has read block b's exchange slot, hence two parities; [2][128] row sums follow
  uint8_t* sO = smem + 16384 + 32768 + 32768;   // 4 quadrants x [32 rows x 128 B], swizzled, for the TMA store
  constexpr uint32_t kDataBytes = 16384 + 32768 + 32768 + 16384;
  float* xchg_base = reinterpret_cast<float*>(smem + kDataBytes);
  float* xchg_l = xchg_base + 512;   // [2 parity][2 halves][128] row sums of the finished tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDataBytes + 4096);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 12;   // [2]
  uint64_t* v_empty = bars + 14;  // [2]
  uint64_t* s_full = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint64_t* s_free = bars + 9;
  uint64_t* q_empty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.q_tiles * p.zcount;
  const int stride = static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    mbar_init(&k_full[0], 1);
    mbar_init(&k_full[1], 1);
    mbar_init(&k_empty[0], 1);
    mbar_init(&k_empty[1], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(pv_done, 1);
    mbar_init(s_free, 8);
    mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                // the previous kernel's results (PDL: everything above ran under its tail)
  pdl_launch_dependents();
  const uint32_t tS = tmem;        // 128 columns
  const uint32_t tO = tmem + 128;  // 64 columns
  const uint32_t tP = tmem + 192;  // 64 columns = 128 fp16 probabilities per row

  // Warps 0 and 1 run their loops with all 32 lanes in warp-uniform control flow and let one elected lane
  // issue: descriptors and coordinates then live in uniform registers and every TMA / tcgen05.mma is a
  // single instruction (inside `if (lane == 0)` each one became an ELECT / R2UR / branch sequence of ~100
  // cycles, which sat on the S -> softmax -> P*V critical path twelve times per key block).
  if (warp == 0) {
    uint32_t kb = 0, tq = 0;
    for (int tile = blockIdx.x; tile < total; tile += stride) {
      FaTile t;
      if (!fa_decode(p, tile, t) || t.nblk == 0) continue;
      mbar_wait(q_empty, (tq & 1u) ^ 1u);    // the last Q K^T of the previous tile has read Q
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 16384);
        tma_load_4d(sQ, &tmQ, q_full, 0, t.q0, 0, t.z);
      }
      __syncwarp();
      ++tq;
      for (int j = 0; j < t.nblk; ++j, ++kb) {
        const uint32_t s = kb & 1u;
        mbar_wait(&k_empty[s], ((kb >> 1) & 1u) ^ 1u);   // S(kb-2) has consumed it
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[s], 16384);
          tma_load_3d(sK + s * 16384, &tmK, &k_full[s], 0, j * kFaBlockKeys, t.zk);
        }
        __syncwarp();
        const uint32_t sv = kb % kVS;
        mbar_wait(&v_empty[sv], ((kb / kVS) & 1u) ^ 1u);     // P*V(kb - kVS) has consumed it
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[sv], 16384);
          tma_load_3d(sV + sv * 16384, &tmV, &v_full[sv], 0, j * kFaBlockKeys, t.zk);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = make_idesc_f16(128);          // S: N = 128 keys
    const uint32_t idesc_o = make_idesc_f16(64, 0, 1);     // O: N = 64, B (= V) is MN-major
    const uint64_t qdesc = make_smem_desc_k_sw128(smem_u32(sQ), 1024);
    const uint32_t vbase = smem_u32(sV), kbase = smem_u32(sK);
    // S(b+1) is issued as soon as the softmax warps have pulled S(b) out of TMEM into registers (s_free),
    // i.e. it runs underneath the whole softmax of block b - also across a tile boundary, where it first
    // waits for the next tile's Q; P*V(b) follows once P(b) is in shared memory.
    // issue_s(b, last): S for global block b; `last` = final block of its tile (Q may then be replaced).
    auto issue_s = [&](uint32_t b, bool last) {
      const uint32_t s = b & 1u;
      mbar_wait(&k_full[s], (b >> 1) & 1u);
      tc_fence_after();
      const uint64_t kdesc = make_smem_desc_k_sw128(kbase + s * 16384, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&k_empty[s]);
        umma_commit(s_full);
        if (last) umma_commit(q_empty);
      }
      __syncwarp();
    };
    // next valid tile at or after `tile`
    auto next_tile = [&](int tile, FaTile& t) -> int {
      for (; tile < total; tile += stride)
        if (fa_decode(p, tile, t) && t.nblk > 0) return tile;
      return total;
    };
    uint32_t kb = 0, tq = 0;
    FaTile cur, nxt;
    int tile = next_tile(blockIdx.x, cur);
    if (tile < total) {
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0, cur.nblk == 1);
    }
    while (tile < total) {
      const int ntile = next_tile(tile + stride, nxt);
      for (int j = 0; j < cur.nblk; ++j, ++kb) {
        const bool more = j + 1 < cur.nblk;
        if (more || ntile < total) {
          mbar_wait(s_free, kb & 1u);   // every softmax warp holds S(kb) in registers
          tc_fence_after();
          if (!more) {                  // first block of the next tile: its Q must have landed
            mbar_wait(q_full, (tq + 1) & 1u);
            tc_fence_after();
          }
          issue_s(kb + 1, more ? (j + 2 == cur.nblk) : (nxt.nblk == 1));
        }
        mbar_wait(p_full, kb & 1u);   // softmax(kb) has written P(kb)
        tc_fence_after();
        const uint32_t sv = kb % kVS;
        mbar_wait(&v_full[sv], (kb / kVS) & 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // B: 16 key rows = 2048 B.  A: 16 keys of P = 8 TMEM columns.
            const uint64_t vdesc = make_smem_desc_mn_sw128(vbase + sv * 16384 + k * 2048, 1024, 1024);
            umma_f16_ts(tO, tP + 8 * k, vdesc, idesc_o, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&v_empty[sv]);
          umma_commit(pv_done);
        }
        __syncwarp();
      }
      ++tq;
      tile = ntile;
      cur = nxt;
    }
  } else {
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;          // which 64 keys of the block / which 32 output columns
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tSh = tS + lane_off + half * 64;
    const uint32_t tOh = tO + lane_off + half * 32;
    const uint32_t tPh = tP + lane_off + half * 32;   // my 64 probabilities = 32 cells
    uint8_t* slab = sO + qd * 4096;            // this quadrant's 32 rows x 128 B (both halves)
    uint32_t kb = 0;
    // The epilogue of a tile (O / l -> fp16 context rows) is deferred into the first key block of the NEXT
    // tile: the row sums cross between the two halves in that block's max exchange (same barrier), and O is
    // read after that block's exponentials, when P*V of the finished tile has long retired - instead of
    // stalling all eight warps on the last P*V, a second barrier and 128 strided 16-byte stores per warp
    // (measured: ~4 000 of ~24 000 cycles per tile).  The rows leave through a swizzled staging slab and one
    // TMA store per quadrant.
    int pend = -1;   // finished tile waiting for its epilogue: img << 12 | (q0 / 128) << 4 | head, or -1
    float l_prev = 0.f;
    // stage O * inv for the pending tile and store it; all 64 threads of the quadrant pair call it
    auto flush_pending = [&](float inv) {
      const int pend_img = pend >> 12, pend_row0 = ((pend >> 4) & 0xff) << 7, pend_head = pend & 15;
      const bool valid = (pend_row0 + row) < p.cnt[pend_img];
      float o[32];
      tmem_ld_32x32(tOh, o);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 w;
        w.x = valid ? pack_half2(o[8 * u + 0] * inv, o[8 * u + 1] * inv) : 0u;
        w.y = valid ? pack_half2(o[8 * u + 2] * inv, o[8 * u + 3] * inv) : 0u;
        w.z = valid ? pack_half2(o[8 * u + 4] * inv, o[8 * u + 5] * inv) : 0u;
        w.w = valid ? pack_half2(o[8 * u + 6] * inv, o[8 * u + 7] * inv) : 0u;
        *reinterpret_cast<uint4*>(slab + lane * 128 + (((half * 4 + u) ^ (lane & 7)) << 4)) = w;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      fa_pair_sync(qd);
      if (half == 0 && lane == 0) {
        tma_store_3d(&tmO, slab, pend_head * 64, pend_row0 + qd * 32, pend_img);
        bulk_commit();
      }
      pend = -1;
    };
    for (int tile = blockIdx.x; tile < total; tile += stride) {
      FaTile t;
      if (!fa_decode(p, tile, t)) continue;
      if (t.nblk == 0) {   // no keys: the message is zero
        uint4* dst = reinterpret_cast<uint4*>(p.ctx + (static_cast<size_t>(t.img) * p.kp + t.q0 + row) * (p.heads * 64) +
                                              (t.z - t.img * p.heads) * 64 + half * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[u] = make_uint4(0u, 0u, 0u, 0u);
        continue;
      }
      float m_used = -INFINITY, l = 0.f;
      for (int j = 0; j < t.nblk; ++j, ++kb) {
        FA_TRACE(0);
        mbar_wait(s_full, kb & 1u);
        tc_fence_after();
        FA_TRACE(1);
        const int kvalid = min(64, t.nk - j * kFaBlockKeys - half * 64);  // valid keys in my half (may be <= 0)
        // my 64 logits -> registers (one TMEM read; S is then free for the next block's Q K^T)
        float v[64];
        tmem_ld_32x32(tSh, v);
        tmem_ld_32x32(tSh + 32, v + 32);
        tmem_ld_wait();
        FA_TRACE(2);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        // keys beyond the count (last block of an image only): -inf logits drop out of the maximum and give
        // exp2(-inf) = 0 below, so the common path carries no per-element masking at all
        if (kvalid < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= kvalid) v[i] = -INFINITY;
        }
        // maximum of my 64 logits (raw; the positive scale is applied once)
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 64; ++i) mx[i & 3] = fmaxf(mx[i & 3], v[i]);
        float* xchg = xchg_base + (kb & 1u) * 256;
        float* xl = xchg_l + (kb & 1u) * 256;
        xchg[half * 128 + row] = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        const bool flush = pend >= 0 && j == 0;   // warp-uniform
        if (flush) {
          xl[half * 128 + row] = l_prev;
          if (half == 0 && lane == 0) bulk_wait_read0();   // the previous store has left the staging slab
        }
        FA_TRACE(3);
        fa_pair_sync(qd);
        const float bm = fmaxf(xchg[row], xchg[128 + row]) * p.scale_log2;
        float inv_prev = 0.f;
        if (flush) inv_prev = 1.0f / (xl[row] + xl[128 + row]);
        FA_TRACE(4);
        float alpha = 1.f;
        bool need = false;
        if (j == 0) {
          m_used = bm;
        } else if (bm > m_used + 8.0f) {
          alpha = fast_exp2(m_used - bm);
          m_used = bm;
          need = true;
        }
        // probabilities: packed fp32x2 scale-and-shift and row sums (half the FMA-pipe instructions), one
        // MUFU.EX2 per element, fp16 pairs -> TMEM A operand.  P and O are still being read / written by the
        // previous P*V until pv_done fires: the first 32 exponentials are computed before that wait.
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        const float2 nm2 = make_float2(-m_used, -m_used);
        float2 ls0 = make_float2(0.f, 0.f), ls1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          uint32_t w[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int i = 32 * h2 + 2 * k;
            float2 e = ffma2(make_float2(v[i], v[i + 1]), sc2, nm2);
            if (kPolyEvery > 0 && (k % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) {
              e = exp2_poly2(e);     // FMA pipe instead of the MUFU
            } else {
              e.x = fast_exp2(e.x);
              e.y = fast_exp2(e.y);
            }
            if (k & 1) ls1 = fadd2(ls1, e); else ls0 = fadd2(ls0, e);
            w[k] = pack_half2(e.x, e.y);
          }
          if (h2 == 0) {
            FA_TRACE(5);
            if (kb > 0) {
              mbar_wait(pv_done, (kb - 1) & 1u);
              tc_fence_after();
            }
            FA_TRACE(6);
            if (__any_sync(0xffffffffu, need)) {
              l *= alpha;
              // two 16-column pieces (the rescale is rare: the running maximum grew by more than 2^8)
#pragma unroll 1
              for (int h = 0; h < 2; ++h) {
                float o[16];
                tmem_ld_32x16(tOh + h * 16, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] *= alpha;
                tmem_st_32x16(tOh + h * 16, o);
              }
            }
          }
          tmem_st_32x16_u32(tPh + 16 * h2, w);
        }
        l += (ls0.x + ls0.y) + (ls1.x + ls1.y);
        tmem_st_wait();
        FA_TRACE(7);
        // O of the finished tile is final (its last P*V retired before this block's pv_done wait) and stays
        // untouched until this block's P*V, which needs all eight p_full arrivals
        if (flush) flush_pending(inv_prev);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);   // one arrival per warp: 8 instead of 256 shared-memory atomics
      }
      // hand the finished tile to the next tile's first block (or to the tail below)
      pend = (t.img << 12) | ((t.q0 >> 7) << 4) | (t.z - t.img * p.heads);
      l_prev = l;
    }
    if (pend >= 0) {   // last tile of this CTA
      float* xl = xchg_l + (kb & 1u) * 256;
      xl[half * 128 + row] = l_prev;
      if (half == 0 && lane == 0) bulk_wait_read0();
      fa_pair_sync(qd);
      const float inv = 1.0f / (xl[row] + xl[128 + row]);
      mbar_wait(pv_done, (kb - 1) & 1u);
      tc_fence_after();
      flush_pending(inv);
    }
    if (half == 0 && lane == 0) bulk_wait_all();   // the store still reads this CTA's shared memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

inline int launch_flash_attention(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                  const CUtensorMap& tmO, FaParams p, int q_tiles, int z, cudaStream_t stream, const char* label) {
  using Kernel = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, FaParams);
  // Variant selection is process-wide (environment), the shared-memory opt-in is per device (common.cuh).
  // SSB_FA_POLY = 0 | 3 | 4 | 6 | 8: one pair of exponentials in `poly` on the FMA pipe.  Measured per 18
  // launches at 64 pairs: 0 -> 4.84 ms, 8 -> 4.57, 6 -> 4.60, 4 -> 4.61, 3 -> 4.54.
  static const int trace_env = [] { const char* e = std::getenv("SSB_FA_TRACE"); return e ? std::atoi(e) : 0; }();
  static const int poly = [] { const char* e = std::getenv("SSB_FA_POLY"); return e ? std::atoi(e) : 3; }();
  static std::atomic<int> trace{trace_env};
  Kernel kernel;
  if (trace_env) kernel = flash_attention_kernel<true, 3>;
  else if (poly == 0) kernel = flash_attention_kernel<false, 0>;
  else if (poly == 4) kernel = flash_attention_kernel<false, 4>;
  else if (poly == 6) kernel = flash_attention_kernel<false, 6>;
  else if (poly == 8) kernel = flash_attention_kernel<false, 8>;
  else kernel = flash_attention_kernel<false, 3>;
  auto configure = [&]() -> int {
    SSB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmemBytes));
    // ask for the full shared-memory carveout so that two CTAs (2 x 101 KB) are co-resident per SM
    SSB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    return SSB_OK;
  };
  SSB_DEVICE_CONFIG(kernel, 1, configure());
  p.q_tiles = q_tiles;
  p.zcount = z;
  const int total = q_tiles * z;
  if (total <= 0) return SSB_OK;
  const int resident = 2 * device_sm_count();
  const int ctas = total < resident ? total : resident;
  SSB_CUDA_CHECK(launch_kernel(kernel, dim3(ctas), dim3(kFaThreads), kFaSmemBytes, stream, 1, tmQ, tmK, tmV, tmO, p));
  if (trace.exchange(trace_env ? 2 : 0) == 1) {   // first launch only: dump the phase time stamps of CTA 0 / warp 2
    SSB_CUDA_CHECK(cudaStreamSynchronize(stream));
    static long long h[64][8];
    SSB_CUDA_CHECK(cudaMemcpyFromSymbol(h, g_fa_trace, sizeof(h)));
    std::fprintf(stderr, "fa trace (%s, %d CTAs): kb | wait_s ld max bar exp1 wait_pv exp2 | gap_to_next | period\n", label, ctas);
    for (int b = 0; b + 1 < 48; ++b)
      std::fprintf(stderr, "%2d | %5lld %5lld %5lld %5lld %5lld %5lld %5lld | %5lld | %5lld\n", b, h[b][1] - h[b][0],
                   h[b][2] - h[b][1], h[b][3] - h[b][2], h[b][4] - h[b][3], h[b][5] - h[b][4], h[b][6] - h[b][5],
                   h[b][7] - h[b][6], h[b + 1][0] - h[b][7], h[b + 1][0] - h[b][0]);
  }
  SSB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  prof_mark(stream, label);
  return SSB_OK;
}

}  // namespace ssb
