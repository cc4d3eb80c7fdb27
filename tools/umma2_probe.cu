// Hardware probe for CTA-pair MMAs (tcgen05.mma.cta_group::2): run once on a B200 before conv_pipe.cuh / umma_core.cuh
// rely on the conventions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma2_probe tools/umma2_probe.cu && build/umma2_probe
// A cluster of two CTAs computes D[256 x N] = A[256 x 64] * B[N x 64]^T with ONE instruction stream issued by the
// leader (cluster rank 0):
//   * A: each CTA holds ITS 128 rows (K-major, 128B swizzle) at the same shared-memory offset;
//   * B: each CTA holds N/2 rows (rank 0: n = 0 .. N/2-1, rank 1: the rest) at the same offset;
//   * D: each CTA's tensor memory receives its own 128 rows x all N columns;
//   * tensor memory is allocated by the same warp of both CTAs with cta_group::2;
//   * completion: tcgen05.commit.cta_group::2 ... multicast::cluster with mask 0b11 arrives on the barrier at the same
//     offset in both CTAs.
// The questions answered: instruction-descriptor M field (256 >> 4), the N split of B, which rows land where.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <unistd.h>
#include "../superslam_b200/csrc/common.cuh"

using namespace ssb;

__device__ __forceinline__ float aval(int r, int c) { return float(((r * 7 + c * 3) % 13) - 6); }
__device__ __forceinline__ float bval(int n, int c) { return float(((n * 5 + c * 11) % 9) - 4) * 0.5f; }

// n_total = 64 or 128
__global__ void __cluster_dims__(2, 1, 1) probe2(int n_total, int* mismatches, float* sample) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* A = smem;                 // 128 rows x 128 B
  uint8_t* B = smem + 128 * 128;     // up to 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 192 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  const int rank = static_cast<int>(cluster_ctarank());
  const int nh = n_total / 2;
  for (int i = tid; i < 128 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__half*>(A + off) = __float2half(aval(rank * 128 + r, c));
  }
  for (int i = tid; i < nh * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__half*>(B + off) = __float2half(bval(rank * nh + r, c));
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = (1u << 4) | ((static_cast<uint32_t>(n_total) >> 3) << 17) | ((256u >> 4) << 24);
    for (int k = 0; k < 4; ++k)
      umma2_f16(tmem, make_smem_desc_k_sw128(smem_u32(A), 1024) + 2 * k, make_smem_desc_k_sw128(smem_u32(B), 1024) + 2 * k, idesc,
                k != 0);
    umma2_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  float v[32];
  int bad = 0;
  for (int col = 0; col < n_total; col += 32) {
    tmem_ld_32x32(tmem + (uint32_t(warp * 32) << 16) + col, v);
    tmem_ld_wait();
    const int m = rank * 128 + warp * 32 + lane;
    for (int j = 0; j < 32; ++j) {
      float e = 0.f;
      for (int c = 0; c < 64; ++c) e += aval(m, c) * bval(col + j, c);
      if (fabsf(e - v[j]) > 1e-3f) ++bad;
      if (m == 133 && col + j == 37) {
        sample[0] = v[j];
        sample[1] = e;
      }
    }
  }
  if (bad) atomicAdd(mismatches + rank, bad);
  if (tid == 0) sample[2 + rank] = __uint_as_float(tmem);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}


// ---- feature tests of the pair protocol (one per process: `umma2_probe <mode>`; a hang is killed by `timeout`) ----------
// mode 1: tcgen05.alloc.cta_group::2 of 512 columns, cluster given as a LAUNCH attribute (not __cluster_dims__)
// mode 2: + remote mbarrier arrive (release.cluster) from both CTAs on the leader's barrier, leader waits acquire.cluster
// mode 3: + pair TMA loads (cp.async.bulk.tensor ... cta_group::2): both CTAs' boxes counted on the leader's barrier
// mode 4: + two multicast commits back to back after an un-swizzled K = 16 pair MMA (conv1a's operand layout)
__global__ void feature(int mode, const __grid_constant__ CUtensorMap tm, int* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);   // [0] remote arrivals, [1] tx bytes, [2], [3] commits
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rank = static_cast<int>(cluster_ctarank());
  if (tid == 0) {
    mbar_init(&bar[0], 2 * 4);   // one arrival per warp of both CTAs
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 1);
    mbar_init(&bar[3], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(slot, 512);
    tmem_relinquish2();
  }
  for (int i = tid; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) out[rank * 8 + 0] = 1 + static_cast<int>(tmem);   // reached: allocation done
  if (mode >= 2) {
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive_cluster(&bar[0], 0);
    if (rank == 0 && warp == 1) mbar_wait_cluster(&bar[0], 0);
    if (tid == 32) out[rank * 8 + 1] = 1;
  }
  if (mode >= 3) {
    if (tid == 0) {
      if (rank == 0) mbar_arrive_expect_tx(&bar[1], 2 * 4096);
      tma_load_3d_pair(smem + 8192, &tm, &bar[1], 0, rank * 32, 0);
    }
    if (rank == 0 && warp == 1) mbar_wait_cluster(&bar[1], 0);
    if (tid == 32) out[rank * 8 + 2] = 1;
  }
  if (mode >= 4) {
    if (rank == 0 && warp == 1) {
      uint64_t d = 0;   // un-swizzled K-major core matrices: LBO 128 (K), SBO 256 (8-row groups)
      d |= static_cast<uint64_t>((smem_u32(smem) >> 4) & 0x3FFF);
      d |= static_cast<uint64_t>(128 >> 4) << 16;
      d |= static_cast<uint64_t>(256 >> 4) << 32;
      d |= static_cast<uint64_t>(1) << 46;
      if (elect_one()) {
        umma2_f16(tmem + 256, d, d + (4096 >> 4), make_idesc2_f16(64), 0u);
        umma2_commit(&bar[2]);
        umma2_commit(&bar[3]);
      }
      __syncwarp();
    }
    mbar_wait(&bar[2], 0);
    mbar_wait(&bar[3], 0);
    if (tid == 32) out[rank * 8 + 3] = 1;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem, 512);
  if (tid == 0) out[rank * 8 + 4] = 1;
}

// ---- MMA rate: cycles per tcgen05.mma (M = 128 per CTA, K = 16, SW128 K-major operands from shared memory), issued back
// to back by one thread: one CTA (cta_group::1) against a pair (cta_group::2), N = 64 / 128 / 256.
template <bool kPairMode>
__global__ void mma_rate(int n_total, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* A = smem;                 // 128 rows x 128 B
  uint8_t* B = smem + 128 * 128;     // up to 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  const int rank = kPairMode ? static_cast<int>(cluster_ctarank()) : 0;
  for (int i = tid; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) {
    if (kPairMode) { tmem_alloc2(slot, 512); tmem_relinquish2(); } else { tmem_alloc(slot, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (kPairMode) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (rank == 0 && tid < 32) {
    const uint32_t idesc = kPairMode ? make_idesc2_f16(n_total) : make_idesc_f16(n_total);
    const uint64_t ad = make_smem_desc_k_sw128(smem_u32(A), 1024), bd = make_smem_desc_k_sw128(smem_u32(B), 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (kPairMode) umma2_f16(tmem + (it & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, 1u);
          else umma_f16(tmem + (it & 1) * 256, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      if (kPairMode) umma2_commit(bar); else umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  } else if (kPairMode) {
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (kPairMode) cluster_sync_all();
  if (tid < 32) {
    if (kPairMode) tmem_dealloc2(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

template <bool kPairMode>
void run_rate(int n) {
  long long* cyc;
  cudaMallocManaged(&cyc, 148 * 8);
  cudaFuncSetAttribute(mma_rate<kPairMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 56 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPairMode ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int iters = 4000;
  cudaLaunchKernelEx(&cfg, mma_rate<kPairMode>, n, 100, cyc);
  cudaLaunchKernelEx(&cfg, mma_rate<kPairMode>, n, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s N=%3d: %6.1f cycles per MMA (M=128 per CTA, K=16; math %d)  %s\n", kPairMode ? "pair  " : "single", n,
         double(cyc[0]) / (4.0 * iters), n / 2, cudaGetErrorString(e));
  cudaFree(cyc);
}

int run_feature(int mode);

int main(int argc, char** argv) {
  if (argc > 1 && std::atoi(argv[1]) == 9) {
    for (int n : {64, 128, 256}) { run_rate<false>(n); run_rate<true>(n); }
    return 0;
  }
  if (argc > 1) return run_feature(std::atoi(argv[1]));
  int* d_bad;
  float* d_s;
  cudaMalloc(&d_bad, 8);
  cudaMalloc(&d_s, 16);
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int n : {64, 128}) {
    cudaMemset(d_bad, 0, 8);
    cudaMemset(d_s, 0, 16);
    probe2<<<2, 128, 40 * 1024>>>(n, d_bad, d_s);
    cudaError_t e = cudaDeviceSynchronize();
    int bad[2] = {-1, -1};
    float s[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) {
      cudaMemcpy(bad, d_bad, 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(s, d_s, 16, cudaMemcpyDeviceToHost);
    }
    unsigned t0, t1;
    memcpy(&t0, &s[2], 4);
    memcpy(&t1, &s[3], 4);
    printf("cta_group::2 M=256 N=%3d -> %s mismatches rank0 %d rank1 %d (row 133 col 37: got %.2f expect %.2f) tmem base %#x / %#x\n", n,
           e == cudaSuccess ? "ok " : cudaGetErrorString(e), bad[0], bad[1], s[0], s[1], t0, t1);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}

#include <cuda.h>
#include <cudaTypedefs.h>
int run_feature(int mode) {
  int* out;
  cudaMallocManaged(&out, 64);
  memset(out, 0, 64);
  __half* w;
  cudaMalloc(&w, 64 * 64 * 2);
  cudaMemset(w, 0, 64 * 64 * 2);
  // weight-like matrix [64 rows][64 k] fp16, 32-row boxes, 128B swizzle
  CUtensorMap tm;
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&enc), cudaEnableDefault, &q);
  cuuint64_t dims[3] = {64, 64, 1};
  cuuint64_t strides[2] = {128, 64 * 128};
  cuuint32_t box[3] = {64, 32, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", int(r)); return 2; }
  cudaFuncSetAttribute(feature, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 48 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, feature, mode, tm, out);
  printf("feature mode %d: launch %s\n", mode, cudaGetErrorString(e));
  fflush(stdout);
  for (int i = 0; i < 30; ++i) {
    if (cudaStreamQuery(0) != cudaErrorNotReady) break;
    struct timespec ts = {0, 100000000};
    nanosleep(&ts, nullptr);
  }
  e = cudaStreamQuery(0);
  printf("feature mode %d: %s | rank0 alloc %d arrive %d tma %d mma %d end %d | rank1 alloc %d arrive %d tma %d mma %d end %d\n", mode,
         e == cudaSuccess ? "completed" : (e == cudaErrorNotReady ? "HANG" : cudaGetErrorString(e)), out[0], out[1], out[2], out[3], out[4],
         out[8], out[9], out[10], out[11], out[12]);
  fflush(stdout);
  if (e == cudaErrorNotReady) _exit(3);
  return e == cudaSuccess ? 0 : 1;
}
