// Hardware probe (run once on a B200): which shared-memory descriptor conventions does tcgen05.mma
// accept for (a) A operands whose start address is shifted by whole 128-byte rows inside a 128B-swizzled
// buffer (needed to reuse one halo tile for all 3x3 taps) and (b) an MN-major B operand (V in P*V)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe tools/umma_probe.cu && /tmp/umma_probe
#include <cstdio>
#include <cstdlib>
#include "../superslam_b200/csrc/common.cuh"

using namespace ssb;

__device__ __forceinline__ float aval(int r, int c) { return float(((r * 7 + c * 3) % 13) - 6); }
__device__ __forceinline__ float bval(int n, int c) { return float(((n * 5 + c * 11) % 9) - 4) * 0.5f; }

// mode 0: K-major A shifted by `shift` rows with base_offset `bo`; B K-major [64 n][64 k].
// mode 1: A K-major unshifted; B MN-major stored [64 k][64 n] (n contiguous), lbo = `bo` bytes.
// mode 2: un-swizzled K-major A [128 x 16] and B [64 x 16] (8-row x 16-byte core matrices stored densely:
//         element (r, k) at (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2), one K=16 MMA; `bo` selects which
//         descriptor field carries which stride: 0 -> LBO = 128 (K direction), SBO = 256 (row groups);
//         1 -> swapped.  (conv1a on the tensor cores: K = 9 taps + bias, padded to 16.)
__device__ __forceinline__ uint64_t desc_k_noswizzle(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;   // layout type 0 = SWIZZLE_NONE
}
__global__ void probe(int mode, int shift, int bo, int* mismatches, float* sample) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __half* A = reinterpret_cast<__half*>(smem);               // 192 rows x 128 B
  __half* B = reinterpret_cast<__half*>(smem + 192 * 128);   // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 256 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int i = tid; i < 192 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(A) + off) = __float2half(aval(r, c));
  }
  for (int i = tid; i < 64 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;  // stored row r, 16-byte chunk c>>3
    const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    // mode 0: row = n, col = k.  mode 1: row = k, col = n.
    const float v = mode == 0 ? bval(r, c) : bval(c, r);
    *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(B) + off) = __float2half(v);
  }
  if (mode == 2) {
    __syncthreads();
    for (int i = tid; i < 128 * 16; i += blockDim.x) {
      const int r = i / 16, k = i % 16;
      const int off = (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
      *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(A) + off) = __float2half(aval(r, k));
    }
    for (int i = tid; i < 64 * 16; i += blockDim.x) {
      const int r = i / 16, k = i % 16;
      const int off = (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
      *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(B) + off) = __float2half(bval(r, k));
    }
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(A) + shift * 128;
    const uint32_t b0 = smem_u32(B);
    if (mode == 2) {
      const uint32_t lbo = bo == 0 ? 128 : 256, sbo = bo == 0 ? 256 : 128;
      umma_f16(tmem, desc_k_noswizzle(a0, lbo, sbo), desc_k_noswizzle(b0, lbo, sbo), make_idesc_f16(64), 0);
    }
    for (int k = 0; k < (mode == 2 ? 0 : 4); ++k) {
      uint64_t ad, bd;
      uint32_t idesc;
      if (mode == 0) {
        ad = make_smem_desc_k_sw128(a0, 1024, bo) + 2 * k;
        bd = make_smem_desc_k_sw128(b0, 1024) + 2 * k;
        idesc = make_idesc_f16(64);
      } else {
        ad = make_smem_desc_k_sw128(a0, 1024) + 2 * k;
        bd = make_smem_desc_mn_sw128(b0 + k * 2048, bo, 1024);
        idesc = make_idesc_f16(64, 0, 1);
      }
      umma_f16(tmem, ad, bd, idesc, k != 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  float v[32];
  int bad = 0;
  for (int col = 0; col < 64; col += 32) {
    tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16) + col, v);
    tmem_ld_wait();
    const int m = warp * 32 + lane;
    for (int j = 0; j < 32; ++j) {
      float e = 0.f;
      for (int c = 0; c < (mode == 2 ? 16 : 64); ++c) e += aval(m + shift, c) * bval(col + j, c);
      if (fabsf(e - v[j]) > 1e-3f) ++bad;
      if (m == 5 && col + j == 3) {
        sample[0] = v[j];
        sample[1] = e;
      }
    }
  }
  if (bad) atomicAdd(mismatches, bad);
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

int main() {
  int* d_bad;
  float* d_s;
  cudaMalloc(&d_bad, 4);
  cudaMalloc(&d_s, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  auto run = [&](int mode, int shift, int bo) {
    cudaMemset(d_bad, 0, 4);
    cudaMemset(d_s, 0, 8);
    probe<<<1, 128, 48 * 1024, 0>>>(mode, shift, bo, d_bad, d_s);
    cudaError_t e = cudaDeviceSynchronize();
    int bad = -1;
    float s[2] = {0, 0};
    if (e == cudaSuccess) {
      cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(s, d_s, 8, cudaMemcpyDeviceToHost);
    }
    printf("mode %d shift %d bo/lbo %4d -> %s mismatches %d (got %.2f expect %.2f)\n", mode, shift, bo,
           e == cudaSuccess ? "ok " : cudaGetErrorString(e), bad, s[0], s[1]);
    return e == cudaSuccess;
  };
  const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 17};
  for (int s : shifts) {
    if (!run(0, s, 0)) return 1;
    if ((s & 7) != 0 && !run(0, s, s & 7)) return 1;
  }
  const int lbos[] = {16, 1024, 2048, 4096, 8192};
  for (int l : lbos)
    if (!run(1, 0, l)) return 1;
  run(2, 0, 0);
  run(2, 0, 1);
  return 0;
}
