// Microbenchmark: ex2.approx throughput per SM (alone, and inside the softmax inner-loop instruction mix),
// and the polynomial exp2 on the FMA pipe, at several warps/SM.  Build: nvcc -arch=sm_100a -O3 -o mufu_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// degree-3 polynomial 2^f on [0,1) + exponent insertion (Cody-Waite, FA4 style)
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float fl = floorf(x);
  const float f = x - fl;
  float p = 0.0555041f;
  p = fmaf(p, f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (static_cast<int>(fl) << 23));
}

template <int MODE>
__global__ void probe(float* out, int iters, float seed) {
  float v[16], acc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed * (i + 1) + threadIdx.x * 1e-3f;
  __shared__ uint4 sm[1024];
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {   // pure MUFU
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ex2(v[i]) - 1.0f;
    } else if (MODE == 1) {   // softmax mix: ffma, ex2, fadd, pack, sts
      float e[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        e[i] = ex2(fmaf(v[i], seed, -1.0f));
        acc[i & 3] += e[i];
      }
      uint4 w0, w1;
      __half2 h;
#define PK(a, b) (h = __floats2half2_rn(a, b), *reinterpret_cast<uint32_t*>(&h))
      w0.x = PK(e[0], e[1]); w0.y = PK(e[2], e[3]); w0.z = PK(e[4], e[5]); w0.w = PK(e[6], e[7]);
      w1.x = PK(e[8], e[9]); w1.y = PK(e[10], e[11]); w1.z = PK(e[12], e[13]); w1.w = PK(e[14], e[15]);
      sm[threadIdx.x] = w0;
      sm[(threadIdx.x + 512) & 1023] = w1;
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += 1e-6f;
    } else if (MODE == 2) {   // polynomial only
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ex2_poly(v[i]) - 1.0f;
    } else {   // half MUFU, half polynomial
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ((i & 1) ? ex2_poly(v[i]) : ex2(v[i])) - 1.0f;
    }
  }
  float s = acc[0] + acc[1] + acc[2] + acc[3];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s + sm[0].x;
}

template <int MODE>
void run(const char* name, int warps_per_sm) {
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 4096;
  const int threads = 32 * (warps_per_sm > 16 ? 16 : warps_per_sm), blocks = 148 * (warps_per_sm > 16 ? warps_per_sm / 16 : 1);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  probe<MODE><<<blocks, threads>>>(out, 64, 0.5f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  probe<MODE><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ops = double(blocks) * threads * iters * 16;
  printf("%-10s warps/SM %2d: %.3f ms  %.1f Gop/s  = %.2f op/clk/SM at %d MHz nominal\n", name, warps_per_sm, ms,
         ops / ms * 1e-6, ops / (ms * 1e-3) / 148 / (clk_khz * 1e3), clk_khz / 1000);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 16, 32}) {
    run<0>("mufu", w);
    run<1>("softmax", w);
    run<2>("poly", w);
    run<3>("half", w);
  }
  return 0;
}
