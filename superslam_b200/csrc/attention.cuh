// Building blocks of the fused attention kernel (attention2.cuh): launch parameters, tcgen05.st / TMEM-operand MMA
// wrappers, packed fp32x2 arithmetic, the MUFU and the FMA-pipe exponentials.
//
// The kernel keeps the N x M logits of the reference graph (softmax(q k^T / 8) v for self attention, both directions of
// the bidirectional cross attention: oracle/lightglue.py _self_block / _cross_block) on the SM:
//   S = Q K^T         tcgen05.mma, fp32 accumulator in TMEM
//   P = exp2(S*c - m) online softmax, fp16 P written back to TMEM with tcgen05.st (two fp16 per 32-bit cell, lane =
//                     query row)
//   O += P V          tcgen05.mma with P as a TMEM A operand and V as an MN-major shared-memory B operand
//                     (V rows = keys, as stored by the QKV epilogue; no transposed copy of V exists)
// The running max is only refreshed when it grows by more than 2^8 (the O rescale is skipped otherwise), which keeps
// P <= 256 in fp16 and is exact after the final division by the row sum.  Keeping P in tensor memory takes 64 KB per key
// block off the shared-memory port (32 KB of stores + 32 KB of operand reads).
// (The first kernel of this file - one query tile per CTA, two CTAs per SM, two threads per row with a max exchange
// through shared memory - was replaced by attention2.cuh in round 2; its measurements are in profiles/README.md.)
#pragma once

#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "umma_core.cuh"

namespace ssb {

constexpr int kFaBlockKeys = 128;

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

}  // namespace ssb
