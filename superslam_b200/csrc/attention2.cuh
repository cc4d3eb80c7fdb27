// Fused attention, second version: TWO 128-query tiles per CTA, one softmax warp group per tile, one thread per query row.
//
// What the first version (round 1: one tile per CTA, two co-resident CTAs per SM, two threads per row) spent per
// 128 x 128 key block and CTA, from the ncu source counters (profiles/ncu_attention_r02.txt): 4 670 warp instructions, of
// which only ~2 050 are the softmax itself (MUFU, FFMA2 / FADD2, F2FP, FMNMX3) - the rest is per-block overhead: the
// max exchange between the two threads of a row (shared memory + a named barrier), barrier polling by four producer / MMA
// warps per SM, address and predicate arithmetic that is repeated per block whatever the block's size.  The kernel is
// issue-bound (58 % of the issue slots, `not selected` among the top stalls), so the overhead is what there is to win.
// Here
//   * a thread owns a whole query row (128 logits per key block): no exchange, no pair barrier, row max and row sum are
//     thread-local;
//   * S is read twice per block, 32 columns at a time with ~64 live registers: a max pass (the exact row maximum of the
//     block; the reference of the exponentials is only raised - and l, O rescaled - when it grows by more than 2^8) and
//     the exponential pass, whose last tcgen05.ld hands S back to the MMA warp (s_free): S(j+1) is computed under the
//     last quarter of softmax(j);
//   * the two tiles of a CTA share every K / V block (half the TMA traffic and barrier traffic per tile) and one
//     producer and one MMA warp; while one warp group computes exponentials the tensor core works for the other.
//   Measured on the way (profiles/README.md): a lazy reference without the max pass (block redone when a row sum betrays
//   an overflow) must keep S until P is written - its softmax warps waited 17 % of the time for the next S; checking per
//   32-key chunk instead, or a second issuer warp, cost more instructions / issue slots than they won.
// TMEM (512 columns, one CTA per SM): S0 | S1 (128 each, fp32 logits), P0 | P1 (64 each: 128 fp16 probabilities per row,
// the A operand of P V), O0 | O1 (64 each).  Shared memory: Q 2 x 32 KB (double-buffered per unit), K and V rings
// (2 x 16 KB each), 2 x 16 KB output staging.
// Warp roles (320 threads): 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = softmax of tile 0,
// 6..9 = softmax of tile 1 (TMEM lane quadrant = warp % 4).
#pragma once

#include "attention.cuh"

namespace ssb {

constexpr int kFa2Threads = 320;
constexpr uint32_t kFa2KS = 2, kFa2VS = 2;   // K / V ring depths
constexpr int kFa2MaxImages = 256;   // keypoint counts cached in shared memory (more images: read from global memory)
constexpr int kFa2SmemBytes = 2 * 32768 /*Q*/ + kFa2KS * 16384 + kFa2VS * 16384 + 2 * 16384 /*O staging*/ + 256 /*barriers*/ +
                              kFa2MaxImages * 4 /*counts*/ + 1024 /*align*/;

// One unit of work: query rows q0 .. q0 + 255 of (image, head) z, i.e. tile 0 and (if it has rows) tile 1.
struct Fa2Unit {
  int z, img, head, q0, nq, nk, zk, nblk;
  bool t1;
};
// `cnt`: the keypoint counts - the kernel's shared-memory copy (every role decodes every unit; from global memory each
// decode is a dependent L2 round trip in front of the unit's first barrier wait)
__device__ __forceinline__ bool fa2_decode(const FaParams& p, const int* cnt, int unit, Fa2Unit& t) {
  const int q_pairs = (p.q_tiles + 1) >> 1;
  t.z = unit / q_pairs;
  t.q0 = (unit - t.z * q_pairs) * 256;
  t.img = t.z / p.heads;
  t.head = t.z - t.img * p.heads;
  t.nq = cnt[t.img];
  if (t.q0 >= t.nq) return false;
  t.t1 = t.q0 + 128 < t.nq;
  t.nk = cnt[t.img ^ p.key_xor];
  t.zk = (t.img ^ p.key_xor) * p.heads + t.head;
  t.nblk = (t.nk + kFaBlockKeys - 1) / kFaBlockKeys;
  return true;
}

template <int kPolyEvery>
__global__ void __launch_bounds__(kFa2Threads, 1)
flash_attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // [2 units][2 tiles][128 x 128 B]
  uint8_t* sK = sQ + 2 * 32768;
  uint8_t* sV = sK + kFa2KS * 16384;
  uint8_t* sO = sV + kFa2VS * 16384;                    // [2 tiles][4 quadrants][32 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * 16384);
  uint64_t* q_full = bars;          // [2]
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* k_full = bars + 4;      // [KS]
  uint64_t* k_empty = bars + 6;
  uint64_t* v_full = bars + 8;      // [VS]
  uint64_t* v_empty = bars + 10;
  uint64_t* s_full = bars + 12;     // [2 tiles]  MMA -> softmax: S is in tensor memory
  uint64_t* p_full = bars + 14;     // [2 tiles]  softmax -> MMA: P is written (and S has been read): 4 warps
  uint64_t* pv_done = bars + 16;    // [2 tiles]  MMA -> softmax: P V has retired (P and O may be touched)
  uint64_t* s_free = bars + 18;     // [2 tiles]  softmax -> MMA: the last column of S is in registers: 4 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  int* s_cnt = reinterpret_cast<int*>(bars + 32);   // [kFa2MaxImages]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_pairs = (p.q_tiles + 1) >> 1;
  const int total = q_pairs * p.zcount;
  const int stride = static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&s_free[i], 4);
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
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                // the previous kernel's results (PDL: everything above ran under its tail)
  pdl_launch_dependents();
  const int images = p.zcount / p.heads;
  const int* cnt = p.cnt;
  if (images <= kFa2MaxImages && total > 2 * stride) {   // only when a CTA walks several units (see umma_core.cuh)
    for (int i = threadIdx.x; i < images; i += blockDim.x) s_cnt[i] = p.cnt[i];
    __syncthreads();
    cnt = s_cnt;
  }

  // next valid unit at or after `unit`
  auto next_unit = [&](int unit, Fa2Unit& t) -> int {
    for (; unit < total; unit += stride)
      if (fa2_decode(p, cnt, unit, t) && t.nblk > 0) return unit;
    return total;
  };

  if (warp == 0) {
    // ---- producer ----
    uint32_t u = 0, kb = 0;
    Fa2Unit t;
    for (int unit = next_unit(blockIdx.x, t); unit < total; unit = next_unit(unit + stride, t), ++u) {
      const uint32_t qb = u & 1u;
      mbar_wait(&q_empty[qb], ((u >> 1) & 1u) ^ 1u);   // every S of the unit that used this buffer has retired
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[qb], t.t1 ? 32768u : 16384u);
        tma_load_4d(sQ + qb * 32768, &tmQ, &q_full[qb], 0, t.q0, 0, t.z);
        if (t.t1) tma_load_4d(sQ + qb * 32768 + 16384, &tmQ, &q_full[qb], 0, t.q0 + 128, 0, t.z);
      }
      __syncwarp();
      for (int j = 0; j < t.nblk; ++j, ++kb) {
        const uint32_t sk = kb % kFa2KS, sv = kb % kFa2VS;
        mbar_wait(&k_empty[sk], ((kb / kFa2KS) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[sk], 16384);
          tma_load_3d(sK + sk * 16384, &tmK, &k_full[sk], 0, j * kFaBlockKeys, t.zk);
        }
        __syncwarp();
        mbar_wait(&v_empty[sv], ((kb / kFa2VS) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[sv], 16384);
          tma_load_3d(sV + sv * 16384, &tmV, &v_full[sv], 0, j * kFaBlockKeys, t.zk);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (whole warp in uniform control flow, one elected lane issues) ----
    const uint32_t idesc_s = make_idesc_f16(128);          // S: N = 128 keys
    const uint32_t idesc_o = make_idesc_f16(64, 0, 1);     // O: N = 64, B (= V) is MN-major
    const uint32_t qbase = smem_u32(sQ), kbase = smem_u32(sK), vbase = smem_u32(sV);
    // S of tile `wg` for the global key block `b`, with Q from unit buffer `qb`
    auto issue_s = [&](int wg, uint32_t qb, uint32_t b) {
      const uint64_t qdesc = make_smem_desc_k_sw128(qbase + qb * 32768 + wg * 16384, 1024);
      const uint64_t kdesc = make_smem_desc_k_sw128(kbase + (b % kFa2KS) * 16384, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + wg * 128, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[wg]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int wg, uint32_t b, bool first) {
      const uint32_t sv = b % kFa2VS;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // B: 16 key rows = 2048 B.  A: 16 keys of P = 8 TMEM columns.
          const uint64_t vdesc = make_smem_desc_mn_sw128(vbase + sv * 16384 + k * 2048, 1024, 1024);
          umma_f16_ts(tmem + 384 + wg * 64, tmem + 256 + wg * 64 + 8 * k, vdesc, idesc_o, (!first || k != 0) ? 1u : 0u);
        }
        umma_commit(&pv_done[wg]);
      }
      __syncwarp();
    };
    auto wait_k = [&](uint32_t b) {
      mbar_wait(&k_full[b % kFa2KS], (b / kFa2KS) & 1u);
      tc_fence_after();
    };
    auto release_k = [&](uint32_t b) {   // after the last S that reads K block b has been issued
      if (elect_one()) umma_commit(&k_empty[b % kFa2KS]);
      __syncwarp();
    };
    uint32_t u = 0, kb = 0, c[2] = {0, 0};
    Fa2Unit cur, nxt;
    int unit = next_unit(blockIdx.x, cur);
    if (unit < total) {
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      wait_k(0);
      issue_s(0, 0, 0);
      if (cur.t1) issue_s(1, 0, 0);
      release_k(0);
      if (cur.nblk == 1) {
        if (elect_one()) umma_commit(&q_empty[0]);
        __syncwarp();
      }
    }
    while (unit < total) {
      const int nunit = next_unit(unit + stride, nxt);
      const bool has_next = nunit < total;
      const uint32_t qb = u & 1u, nqb = qb ^ 1u;
      for (int j = 0; j < cur.nblk; ++j, ++kb) {
        const bool more = j + 1 < cur.nblk;
        const bool s_next = more || has_next;             // some S reads K block kb + 1
        if (s_next) {
          wait_k(kb + 1);
          if (!more) {                                    // first block of the next unit: its Q must have landed
            mbar_wait(&q_full[nqb], ((u + 1) >> 1) & 1u);
            tc_fence_after();
          }
        }
        mbar_wait(&v_full[kb % kFa2VS], (kb / kFa2VS) & 1u);
        tc_fence_after();
        // per tile: S(kb + 1) as soon as the softmax warps hold the last column of S(kb) (s_free: under the last quarter
        // of softmax(kb)), P(kb) V(kb) when P(kb) is written.  Tile 0 first: its warps run about half a block ahead.
        mbar_wait(&s_free[0], c[0] & 1u);
        tc_fence_after();
        if (more) issue_s(0, qb, kb + 1);
        else if (has_next) issue_s(0, nqb, kb + 1);
        mbar_wait(&p_full[0], c[0] & 1u);
        tc_fence_after();
        issue_pv(0, kb, j == 0);
        ++c[0];
        if (cur.t1) {
          mbar_wait(&s_free[1], c[1] & 1u);
          tc_fence_after();
        }
        if (more) {
          if (cur.t1) issue_s(1, qb, kb + 1);
        } else if (has_next && nxt.t1) {
          issue_s(1, nqb, kb + 1);                        // S1's buffer is free: every S1 issued so far has been read
        }
        if (cur.t1) {
          mbar_wait(&p_full[1], c[1] & 1u);
          tc_fence_after();
          issue_pv(1, kb, j == 0);
          ++c[1];
        }
        if (s_next) release_k(kb + 1);
        if (elect_one()) umma_commit(&v_empty[kb % kFa2VS]);
        __syncwarp();
        // Q buffers: the unit's last S (block nblk - 1) has just been issued when j == nblk - 2; a next unit of one
        // block has had its only S issued when !more
        if (j + 2 == cur.nblk) {
          if (elect_one()) umma_commit(&q_empty[qb]);
          __syncwarp();
        }
        if (!more && has_next && nxt.nblk == 1) {
          if (elect_one()) umma_commit(&q_empty[nqb]);
          __syncwarp();
        }
      }
      unit = nunit;
      cur = nxt;
      ++u;
    }
  } else {
    // ---- softmax + epilogue: warp group wg owns tile wg, a thread owns a query row ----
    const int wg = (warp - 2) >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tS = tmem + lane_off + wg * 128;
    const uint32_t tP = tmem + lane_off + 256 + wg * 64;
    const uint32_t tO = tmem + lane_off + 384 + wg * 64;
    uint8_t* slab = sO + wg * 16384 + qd * 4096;   // this warp's 32 rows x 128 B
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    uint32_t c = 0;   // key blocks this warp group has processed
    Fa2Unit t;
    for (int unit = next_unit(blockIdx.x, t); unit < total; unit = next_unit(unit + stride, t)) {
      if (wg == 1 && !t.t1) continue;
      const int my_q0 = t.q0 + wg * 128;
      float m_used = 0.f, l = 0.f;
      for (int j = 0; j < t.nblk; ++j, ++c) {
        mbar_wait(&s_full[wg], c & 1u);
        tc_fence_after();
        const int kvalid = t.nk - j * kFaBlockKeys;   // valid keys in this block (>= 1; < 128 only in an image's last block)
        // The 128 logits of the row are read ONCE into registers and S goes back to the MMA warp at once: S(j+1) is
        // computed under the whole softmax(j).  (The two-pass version read S twice, 32 columns at a time: its max pass
        // was bound by the tcgen05.ld latency - 19 % of the softmax warps' time - and they waited another 12 % for S.)
        float s[128];
        tmem_ld_32x32(tS, s);
        tmem_ld_32x32(tS + 32, s + 32);
        tmem_ld_32x32(tS + 64, s + 64);
        tmem_ld_32x32(tS + 96, s + 96);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[wg]);
        if (kvalid < kFaBlockKeys) {   // warp-uniform, last block of an image only: keys beyond the count -> exp2(-inf) = 0
#pragma unroll
          for (int e = 0; e < 128; ++e)
            if (e >= kvalid) s[e] = -INFINITY;
        }
        // exact row maximum of the block (raw logits; the positive scale is applied once)
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int e = 0; e < 128; ++e) mx[e & 3] = fmaxf(mx[e & 3], s[e]);
        const float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * p.scale_log2;
        // Reference of the exponentials: the running maximum, refreshed only when a block exceeds it by more than 2^8
        // (P <= 256 in fp16, exact after the final division by l); then l and O are rescaled.
        float alpha = 1.0f;
        bool need = false;
        if (j == 0) {
          m_used = bm;
        } else if (bm > m_used + 8.0f) {
          alpha = fast_exp2(m_used - bm);
          m_used = bm;
          need = true;
        }
        const float2 nm2 = make_float2(-m_used, -m_used);
        float2 ls0 = make_float2(0.f, 0.f), ls1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t w[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float2 e = ffma2(make_float2(s[32 * i + 2 * k], s[32 * i + 2 * k + 1]), sc2, nm2);
            if (kPolyEvery > 0 && (k % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) {
              e = exp2_poly2(e);     // FMA pipe instead of the MUFU
            } else {
              e.x = fast_exp2(e.x);
              e.y = fast_exp2(e.y);
            }
            if (k & 1) ls1 = fadd2(ls1, e); else ls0 = fadd2(ls0, e);
            w[k] = pack_half2(e.x, e.y);
          }
          if (i == 0 && c > 0) {   // P and O are still in use by the previous P V until pv_done fires
            mbar_wait(&pv_done[wg], (c - 1) & 1u);
            tc_fence_after();
            if (__any_sync(0xffffffffu, need)) {   // rare: the running maximum grew by more than 2^8
              l *= alpha;
#pragma unroll 1
              for (int h = 0; h < 4; ++h) {
                float o[16];
                tmem_ld_32x16(tO + h * 16, o);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) o[e] *= alpha;
                tmem_st_32x16(tO + h * 16, o);
              }
            }
          }
          tmem_st_32x16_u32(tP + 16 * i, w);
        }
        const float lsum = (ls0.x + ls0.y) + (ls1.x + ls1.y);
        l += lsum;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[wg]);   // one arrival per warp; also tells the MMA warp that S has been read
      }
      // ---- tile epilogue: O / l -> fp16 context rows (this warp's 32 rows through its staging slab, one TMA store)
      mbar_wait(&pv_done[wg], (c - 1) & 1u);
      tc_fence_after();
      {
        const bool valid = my_q0 + row < t.nq;
        const float inv = 1.0f / l;
        if (lane == 0) bulk_wait_read0();   // the previous store has left the slab
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float o[32];
          tmem_ld_32x32(tO + h * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 wv;
            wv.x = valid ? pack_half2(o[8 * g + 0] * inv, o[8 * g + 1] * inv) : 0u;
            wv.y = valid ? pack_half2(o[8 * g + 2] * inv, o[8 * g + 3] * inv) : 0u;
            wv.z = valid ? pack_half2(o[8 * g + 4] * inv, o[8 * g + 5] * inv) : 0u;
            wv.w = valid ? pack_half2(o[8 * g + 6] * inv, o[8 * g + 7] * inv) : 0u;
            *reinterpret_cast<uint4*>(slab + lane * 128 + (((h * 4 + g) ^ (lane & 7)) << 4)) = wv;
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, slab, t.head * 64, my_q0 + qd * 32, t.img);
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_all();   // the store still reads this CTA's shared memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Units without keys (nblk == 0) produce a zero message; the main kernel skips them, this one writes the zeros.
__global__ void fa2_zero_ctx_kernel(const FaParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const int img = blockIdx.y;
  const int nq = p.cnt[img], nk = p.cnt[img ^ p.key_xor];
  if (nk > 0) return;
  const int row = blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= nq) return;
  uint4* dst = reinterpret_cast<uint4*>(p.ctx + (static_cast<size_t>(img) * p.kp + row) * (p.heads * 64));
  dst[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);   // 32 lanes x 16 B = one 512-byte row
}

inline int launch_flash_attention2(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV,
                                   const CUtensorMap& tmO, FaParams p, int q_tiles, int z, int images, cudaStream_t stream,
                                   const char* label) {
  using Kernel = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, FaParams);
  // one pair of exponentials in SSB_FA_POLY on the FMA pipe (default 3; measured per 9 self-attention launches at 64
  // pairs with S held in registers: 0 -> 2.04 ms, 2 -> 2.09, 3 -> 2.04, 4 -> 1.98.  4 is within noise of 3 and moved the
  // score error of the trained-like test to 1.08e-3 of the logit scale, over the 1e-3 the parity tests allow.)
  static const int poly = [] { const char* e = std::getenv("SSB_FA_POLY"); return e ? std::atoi(e) : 3; }();
  Kernel kernel;
  if (poly == 0) kernel = flash_attention2_kernel<0>;
  else if (poly == 2) kernel = flash_attention2_kernel<2>;
  else if (poly == 4) kernel = flash_attention2_kernel<4>;
  else kernel = flash_attention2_kernel<3>;
  auto configure = [&]() -> int {
    SSB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFa2SmemBytes));
    return SSB_OK;
  };
  SSB_DEVICE_CONFIG(kernel, 1, configure());
  p.q_tiles = q_tiles;
  p.zcount = z;
  const int units = ((q_tiles + 1) / 2) * z;
  if (units <= 0) return SSB_OK;
  if (p.key_xor != 0) {   // cross attention only: an image may face a partner without keypoints
    SSB_CUDA_CHECK(launch_kernel(fa2_zero_ctx_kernel, dim3((p.kp + 7) / 8, images), dim3(32, 8), 0, stream, 1, p));
    count_launch();
  }
  const int sms = device_sm_count();
  const int ctas = units < sms ? units : sms;
  SSB_CUDA_CHECK(launch_kernel(kernel, dim3(ctas), dim3(kFa2Threads), kFa2SmemBytes, stream, 1, tmQ, tmK, tmV, tmO, p));
  count_launch();
  prof_mark(stream, label);
  return SSB_OK;
}

}  // namespace ssb
