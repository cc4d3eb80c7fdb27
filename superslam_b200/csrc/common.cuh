// Shared device/host helpers for the sm_100a kernels: error handling, PTX wrappers for mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05 (TMEM alloc / MMA / commit / ld) and descriptor builders.
// Bit layouts follow the PTX ISA tcgen05 matrix/instruction descriptor tables (the same fields
// CUTLASS names in cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <utility>

namespace ssb {

// ------------------------------------------------------------------------------------------------
// host-side error plumbing: every C-ABI entry returns an int status, nothing throws across it.
// ------------------------------------------------------------------------------------------------
// Status codes: identical to the macros of include/superslam_b200.h (kept as macros so both headers
// can be included together).
#ifndef SSB_OK
#define SSB_OK 0
#define SSB_ERR_INVALID 1   /* bad argument */
#define SSB_ERR_CUDA 2      /* a CUDA runtime/driver call failed */
#define SSB_ERR_IO 3        /* weight archive unreadable / malformed */
#define SSB_ERR_EXHAUSTED 4 /* descriptor pool has no free slot */
#define SSB_ERR_NODEVICE 5  /* no sm_100 device */
#endif

void set_last_error(const char* fmt, ...);
// Number of kernels this library has launched in this process (bench.py reports the delta).
void count_launch(int n = 1);
long long launch_count();
// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg):
// prof_begin() marks the start of a step, prof_mark() is called after every kernel launch with that
// kernel's label, prof_collect() (after a stream sync) folds event deltas into per-label totals.
void prof_enable(bool on);
bool prof_enabled();
void prof_begin(cudaStream_t stream);
void prof_mark(cudaStream_t stream, const char* label);
void prof_collect();
int prof_report(char* buf, size_t bytes);
const char* last_error();
// Per-device launch configuration.  cudaFuncSetAttribute (the > 48 KB dynamic shared-memory opt-in, the carveout),
// the SM count and cluster occupancy belong to ONE device, while every handle of this library names its own
// device_id and contexts may live on several host threads: a process-wide `static bool configured` would leave the
// second GPU's kernels without the opt-in.  device_config_begin() takes the registry lock and returns the int slot
// of (current device, key), zero on first use; the caller configures the device if the slot is below what it
// needs, stores the new value and calls device_config_end().  device_sm_count() is cached per device the same way.
int* device_config_begin(const void* key);
void device_config_end();
int device_sm_count();

#define SSB_CUDA_CHECK(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ssb::set_last_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                  \
                          cudaGetErrorString(_e));                                            \
      return SSB_ERR_CUDA;                                                               \
    }                                                                                         \
  } while (0)

#define SSB_CHECK(cond, code, ...)                                                            \
  do {                                                                                        \
    if (!(cond)) {                                                                            \
      ssb::set_last_error(__VA_ARGS__);                                                       \
      return (code);                                                                          \
    }                                                                                         \
  } while (0)

// Run `stmt` (an int-status expression) once per device and `want` level for `key`.
#define SSB_DEVICE_CONFIG(key, want, stmt)                                                    \
  do {                                                                                        \
    int* _slot = ssb::device_config_begin(reinterpret_cast<const void*>(key));                \
    int _st = SSB_OK;                                                                         \
    if (*_slot < (want)) {                                                                    \
      _st = (stmt);                                                                           \
      if (_st == SSB_OK) *_slot = (want);                                                     \
    }                                                                                         \
    ssb::device_config_end();                                                                 \
    if (_st != SSB_OK) return _st;                                                            \
  } while (0)

#define SSB_RETURN_IF(expr)                                                                   \
  do {                                                                                        \
    int _s = (expr);                                                                          \
    if (_s != SSB_OK) return _s;                                                         \
  } while (0)

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// Every kernel of the pair pipeline is launched with cudaLaunchAttributeProgrammaticStreamSerialization (launch_kernel
// below): its CTAs may become resident - and run their prologue: barrier initialisation, TMEM allocation, tensor-map
// prefetch - while the previous kernel of the stream is still draining, instead of after the grid-to-grid launch
// latency (2 - 4 us x ~100 launches per step).  pdl_wait() blocks until the previous kernel has completed and its
// writes are visible; it must be executed by EVERY CTA before its first global-memory access and before it exits (a
// grid whose CTAs all left without waiting would "complete" before its predecessor and release ITS successor early).
// pdl_launch_dependents() lets the next kernel's CTAs be scheduled as soon as resources allow.  Both are no-ops for a
// kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU box.  The bound is wall
// time (4 s on %globaltimer, checked every 1024 polls) - a try_wait may itself sleep for a system-dependent time, so
// a poll count says little - and the message names the barrier (shared-memory address) and the phase waited for.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
static __device__ __noinline__ void mbar_timeout(uint32_t bar_addr, uint32_t parity) {
  printf("ssb: mbarrier wait timed out (block %d,%d,%d thread %d, barrier smem+%#x, parity %u)\n", blockIdx.x, blockIdx.y,
         blockIdx.z, threadIdx.x, bar_addr, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) mbar_timeout(smem_u32(bar), parity);
    }
  }
}

// ---- thread-block clusters: distributed shared memory --------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `local` (a shared::cta address) in the CTA with rank `rank`
__device__ __forceinline__ uint32_t cluster_map(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Asynchronous store of two floats into another CTA's shared memory; completion (8 bytes) is signalled on
// an mbarrier in THAT CTA, whose phase completes once the expected byte count has landed - the data is
// then visible to threads that observe the phase with an ordinary (CTA-scope) wait.  No cluster-scope
// acquire is needed on the consumer side (which would invalidate its L1).
__device__ __forceinline__ void st_async_f32x2(uint32_t remote_addr, float a, float b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(remote_addr),
               "f"(a), "f"(b), "r"(remote_bar)
               : "memory");
}

// ---- TMA ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1), "r"(c2)
      : "memory");
}

// Multicast form: the box lands at the same CTA-relative shared-memory offset of every CTA in `cta_mask`
// (cluster ranks), and each of them gets the complete_tx on its own barrier at the same offset.
__device__ __forceinline__ void tma_load_3d_mcast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                                  int c1, int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}

// TMA stores (shared -> global, bulk async group).  OOB parts of the box are clipped.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 / TMEM ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(slot_in_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers fp16/bf16 inputs with fp32 accumulation.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, arriving on the barrier at this offset in every CTA of `cta_mask` (multicast operand stages are
// released to every producer that writes them).
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread l of the warp receives row (lane_base + l).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- CTA pairs (cta_group::2) --------------------------------------------------------------------
// Two CTAs of a cluster (ranks 0 and 1, the two SMs of a TPC) execute ONE tcgen05.mma stream issued by the leader
// (rank 0): M = 256, each CTA supplies its own 128 rows of A and N/2 rows of B from the same shared-memory offsets and
// receives its 128 rows x N columns in its own tensor memory - every CTA reads half of B, which is what moves the
// shared-memory-operand-bound N = 64 convolutions (conventions checked on hardware: tools/umma2_probe.cu).  Every tcgen05
// instruction of such a kernel carries cta_group::2; tensor memory is allocated by the same warp of both CTAs.
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed.
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
// Instruction descriptor of a pair MMA: fp16 A/B, fp32 accumulate, M = 256 (128 rows per CTA), N = n (n/2 rows of B per CTA).
__device__ __forceinline__ uint32_t make_idesc2_f16(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}
// Arrive - with release semantics at cluster scope - on the barrier at the same offset in CTA `rank` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_map(smem_u32(bar), rank)) : "memory");
}
// The same without cluster-scope release (default semantics: release at CTA scope): for arrivals that publish nothing the
// remote waiter reads through the generic proxy - "this CTA has finished reading its tensor memory", or operands that
// this SM's own tensor core reads from this CTA's shared memory after the writer's fence.proxy.async.  (The cluster-scope
// release costs every arriving thread a MEMBAR: the top stall of the first pair kernel.)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_map(smem_u32(bar), rank)) : "memory");
}
// Wait that also acquires what other CTAs of the cluster released before arriving.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  uint64_t t0 = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    if ((++spins & 1023u) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) mbar_timeout(smem_u32(bar), parity);
    }
  }
}
// TMA load of a CTA pair: the box lands in THIS CTA's shared memory, the bytes are counted on the LEADER's barrier
// (the shared::cluster address of rank 0's copy of it).
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_map(smem_u32(bar), 0)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_map(smem_u32(bar), 0)), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}

// ---- descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 64 fp16 (128 B),
// 8-row swizzle atoms (1024 B) stacked every `sbo_bytes`.  base_offset is the (addr>>7)&7 phase for
// start addresses that are not 1024-byte aligned.
__device__ __forceinline__ uint64_t make_smem_desc_k_sw128(uint32_t smem_addr, uint32_t sbo_bytes,
                                                           uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);       // start address   [0,14)
  d |= static_cast<uint64_t>(1) << 16;                         // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32; // SBO             [32,46)
  d |= static_cast<uint64_t>(1) << 46;                         // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(base_offset & 7) << 49;           // base offset     [49,52)
  d |= static_cast<uint64_t>(2) << 61;                         // layout: SWIZZLE_128B
  return d;
}
// MN-major operand with 128-byte swizzle (used for V in P*V): 64 MN-elements (128 B) contiguous per
// K-row, 8 K-rows per atom (1024 B), atoms along K every `sbo_bytes`, along MN every `lbo_bytes`.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                            uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor for kind::f16: fp16 A/B, fp32 accumulate, M=128, N=n.
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t n, uint32_t a_mn_major = 0,
                                                   uint32_t b_mn_major = 0) {
  return (1u << 4)            // c_format = F32
         | (0u << 7)          // a_format = F16
         | (0u << 10)         // b_format = F16
         | (a_mn_major << 15) // a_major
         | (b_mn_major << 16) // b_major
         | ((n >> 3) << 17)   // n_dim
         | ((128u >> 4) << 24);  // m_dim
}

// ---- small math helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

#endif  // __CUDACC__

// Host: launch `kernel` with programmatic stream serialization (see pdl_wait above) unless SSB_PDL=0, optionally as
// thread-block clusters of `cluster` CTAs.  Returns the launch status.
// Measured on B200 (profiles/README.md, round 2): programmatic edges cut the single-pair call by 3 - 13 % (a latency-bound
// chain of ~100 small launches) but cost ~2 % of the throughput of a 64-pair step, whose kernels each fill the GPU for
// 50 - 300 us.  So the scope is set by the caller from the batch it is about to enqueue.
bool pdl_enabled();
bool pdl_set_scope(bool on);   // for the calling thread; returns the previous setting
constexpr int kPdlMaxImages = 8;   // PDL for calls on at most this many images
struct PdlScope {
  bool old;
  explicit PdlScope(int images) : old(pdl_set_scope(images <= kPdlMaxImages)) {}
  ~PdlScope() { pdl_set_scope(old); }
};
#ifdef __CUDACC__
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 unsigned cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

// Host: encode TMA tensor maps (driver entry point resolved through the runtime, no libcuda link).
// `elem_strides` (optional, rank entries): traversal stride per dimension - with stride s the box covers
// box[i] elements of the tensor and ceil(box[i] / s) of them are loaded (strided convolutions).
int encode_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box,
                    const uint32_t* elem_strides = nullptr);

}  // namespace ssb
