// Microbenchmark: tcgen05.ld (LDTM) throughput of one SM - how many cycles a 128 x 128 fp32 logit block costs to read from
// tensor memory, with 4 warps (one per lane quadrant) and 8 warps (two per quadrant), .x32 and .x16 shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tools/tmem_probe.cu ; run on one B200.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ void ldtm(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void ldtm<32>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
template <>
__device__ __forceinline__ void ldtm<16>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// every warp reads `cols` columns of its lane quadrant `iters` times; MUFU_WARPS extra warps run ex2 beside them
template <int X>
__global__ void probe(uint32_t* out, long long* cycles, int iters, int cols, int ld_warps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  uint32_t acc = 0;
  float f = lane * 1e-3f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < ld_warps) {
    const uint32_t base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    for (int it = 0; it < iters; ++it) {
      for (int c = 0; c < cols; c += 2 * X) {   // two loads in flight
        uint32_t a[X], b[X];
        ldtm<X>(base + c, a);
        ldtm<X>(base + c + X, b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < X; ++i) acc ^= a[i] + b[i];
      }
    }
  } else {   // MUFU load beside the reads
    for (int it = 0; it < iters * (cols / 32); ++it) {
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f));
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(f);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int X>
static void run(const char* name, int warps, int ld_warps, int cols) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  probe<X><<<148, warps * 32>>>(out, cyc, 10, cols, ld_warps);
  probe<X><<<148, warps * 32>>>(out, cyc, iters, cols, ld_warps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = static_cast<double>(iters) * cols * 4.0 * 32 * ld_warps;   // per SM
  std::printf("%-34s warps %2d (ld %d) cols %3d: %8.0f cycles per pass of the block, %6.1f B/clk/SM  (%s)\n", name, warps, ld_warps,
              cols, static_cast<double>(h[0]) / iters, bytes / static_cast<double>(h[0]), cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<32>("x32, 4 warps, 128 cols", 4, 4, 128);
  run<16>("x16, 4 warps, 128 cols", 4, 4, 128);
  run<32>("x32, 8 warps (2 per quadrant)", 8, 8, 128);
  run<32>("x32, 4 ld warps + 4 MUFU warps", 8, 4, 128);
  run<32>("x32, 8 ld warps + 8 MUFU warps", 16, 8, 128);
  return 0;
}
