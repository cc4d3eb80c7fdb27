// Host-side plumbing shared by the SuperPoint and LightGlue runtimes: last-error string, TMA tensor
// map encoding through the driver entry point, device checks.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace ssb {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// SSB_PDL=0|1 forces it off / on; otherwise the calling thread's current scope decides (PdlScope, common.cuh)
static thread_local bool g_pdl_scope = true;
bool pdl_enabled() {
  static const int forced = [] { const char* e = std::getenv("SSB_PDL"); return e == nullptr ? -1 : (std::atoi(e) != 0 ? 1 : 0); }();
  return forced >= 0 ? forced != 0 : g_pdl_scope;
}
bool pdl_set_scope(bool on) {
  const bool old = g_pdl_scope;
  g_pdl_scope = on;
  return old;
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

namespace {
struct DeviceConfig {
  std::mutex mu;
  std::map<std::pair<int, const void*>, int> slots;
  int sms[64] = {};
};
DeviceConfig& device_config() {
  static DeviceConfig c;
  return c;
}
}  // namespace
int* device_config_begin(const void* key) {
  DeviceConfig& c = device_config();
  c.mu.lock();
  int dev = 0;
  cudaGetDevice(&dev);
  return &c.slots[std::make_pair(dev, key)];
}
void device_config_end() { device_config().mu.unlock(); }
int device_sm_count() {
  DeviceConfig& c = device_config();
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(c.mu);
  int& sms = c.sms[dev & 63];
  if (sms == 0) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

namespace {
struct Prof {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<const char*> label;
  size_t n = 0;
  std::map<std::string, std::pair<double, long long>> acc;
  std::mutex mu;
};
Prof& prof() {
  static Prof p;
  return p;
}
cudaEvent_t prof_event(Prof& p) {
  if (p.n >= p.ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    p.ev.push_back(e);
    p.label.push_back(nullptr);
  }
  return p.ev[p.n];
}
}  // namespace

void prof_enable(bool on) {
  std::lock_guard<std::mutex> g(prof().mu);
  prof().on = on;
}
bool prof_enabled() { return prof().on; }
void prof_begin(cudaStream_t stream) {
  Prof& p = prof();
  if (!p.on) return;
  std::lock_guard<std::mutex> g(p.mu);
  cudaEventRecord(prof_event(p), stream);
  p.label[p.n++] = nullptr;  // nullptr = step start marker
}
void prof_mark(cudaStream_t stream, const char* label) {
  Prof& p = prof();
  if (!p.on) return;
  std::lock_guard<std::mutex> g(p.mu);
  cudaEventRecord(prof_event(p), stream);
  p.label[p.n++] = label ? label : "?";
}
void prof_collect() {
  Prof& p = prof();
  std::lock_guard<std::mutex> g(p.mu);
  for (size_t i = 1; i < p.n; ++i) {
    if (p.label[i] == nullptr) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.ev[i - 1], p.ev[i]) == cudaSuccess) {
      auto& a = p.acc[p.label[i]];
      a.first += ms;
      a.second += 1;
    }
  }
  p.n = 0;
}
int prof_report(char* buf, size_t bytes) {
  Prof& p = prof();
  std::lock_guard<std::mutex> g(p.mu);
  std::string out;
  char line[256];
  for (auto& kv : p.acc) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.second, kv.second.first);
    out += line;
  }
  p.acc.clear();
  if (out.size() + 1 > bytes) return SSB_ERR_INVALID;
  std::memcpy(buf, out.c_str(), out.size() + 1);
  return SSB_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int encode_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  SSB_CHECK(fn != nullptr, SSB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank),
                  const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSB_CHECK(r == CUDA_SUCCESS, SSB_ERR_CUDA,
            "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]",
            static_cast<int>(r), rank, (unsigned long long)dims[0],
            (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
            (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0,
            rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return SSB_OK;
}

}  // namespace ssb
