// Test double of the C-ABI (include/superslam_b200.h): the entry points the C++ adapter calls, with canned, fully
// predictable behaviour and call recording, so that the adapter + the reference's StereoFrontEnd above it can be RUN
// on a machine without a GPU (tests/test_dropin_adapter.py).  What is checked there is the adapter's own work:
// cv::Mat -> (pointer, stride, channels), keypoint / DMatch construction, slot ownership through
// shared_ptr deleters, status -> empty-result mapping.  It is linked ONLY into oracle/_ref/libdropin_fake.so;
// nothing in the product can reach it.  TEST INFRASTRUCTURE.
//
// Canned rules (mirrored by the test):
//   extract, image i:   n = min(max_keypoints, 4 * px(0,0));  px(0,0) == 255 -> SSB_ERR_CUDA for the whole call
//                       keypoint k = (px(0,1) + 2k, px(1,0) + k % 7), score 1 / (1 + k)      [px(1,0) checks row_stride]
//                       descriptor row k, column c (fp32 in fake "device" memory) = px(0,2) + k + c / 1024
//                       slots from a LIFO free list (include/DescriptorPool.h:25-44); none free -> slot -1, null
//                       pointer, SSB_ERR_EXHAUSTED, keypoints still delivered (src/SuperPoint.cc:724-727)
//   match, query i:     i % 3 == 0 -> unmatched (-1, 0);  else train (7 i + 3) % n1, score 0.25 + 0.5 (i % 2)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "superslam_b200.h"

struct ssb_superpoint {
  int max_kp = 0;
  std::vector<int> free_slots, refs;
  std::vector<std::vector<float>> slot_mem;
};
struct ssb_lightglue {
  bool clone = false;
  int width = 0, height = 0;
};

// Served mode (fake_set_servers): extract / match are answered by callbacks - the test plugs the CPU oracle in - and the
// slot memory holds real fp16 rows [n][256], so that the adapter classes can be compared with the reference's own
// SuperPoint / LightGlue classes running over the same networks (tests/test_oracle_ref_e2e.py).
using ExtractServer = int (*)(const uint8_t* image, int height, int width, int row_stride, int max_keypoints, float* xy,
                              float* score, unsigned short* desc_f16);   // returns the keypoint count
using MatchServer = void (*)(const float* xy0, int n0, const unsigned short* d0, const float* xy1, int n1,
                             const unsigned short* d1, int image_width, int image_height, int32_t* matches0, float* mscores0);
static ExtractServer g_extract_server = nullptr;
static MatchServer g_match_server = nullptr;

namespace {
ssb_superpoint* g_sp = nullptr;
int g_release_calls = 0, g_sp_destroyed = 0, g_lg_alive = 0;
struct MatchCall {
  int clone = -1, device_path = -1, n0 = 0, n1 = 0, slot0 = -2, slot1 = -2;
  double xy0_sum = 0, xy1_sum = 0, d0_first = 0, d1_first = 0;
} g_last;

int slot_of(const void* p) {
  if (!g_sp || !p) return -1;
  for (size_t s = 0; s < g_sp->slot_mem.size(); ++s)
    if (g_sp->slot_mem[s].data() == p) return static_cast<int>(s);
  return -1;
}
int run_match(ssb_lightglue* lg, int device_path, const float* xy0, int n0, const float* xy1, int n1, int32_t* m, float* s) {
  if (!lg || n0 < 0 || n1 < 0) return SSB_ERR_INVALID;
  g_last.clone = lg->clone, g_last.device_path = device_path, g_last.n0 = n0, g_last.n1 = n1;
  g_last.xy0_sum = g_last.xy1_sum = 0;
  for (int i = 0; i < 2 * n0; ++i) g_last.xy0_sum += xy0[i];
  for (int i = 0; i < 2 * n1; ++i) g_last.xy1_sum += xy1[i];
  for (int i = 0; i < n0; ++i) {
    const bool hit = n1 > 0 && i % 3 != 0;
    m[i] = hit ? (7 * i + 3) % n1 : -1;
    s[i] = hit ? 0.25f + 0.5f * static_cast<float>(i % 2) : 0.0f;
  }
  return SSB_OK;
}
}  // namespace

extern "C" {

int ssb_sp_create(const char* weights_path, int max_keypoints, double, int, int num_slots, int, ssb_superpoint** out) {
  *out = nullptr;
  if (!weights_path || std::string(weights_path) == "missing") return SSB_ERR_IO;
  ssb_superpoint* sp = new ssb_superpoint;
  sp->max_kp = max_keypoints;
  const int n = num_slots > 0 ? num_slots : 8;
  for (int i = n - 1; i >= 0; --i) sp->free_slots.push_back(i);
  sp->refs.assign(n, 0);
  sp->slot_mem.assign(n, std::vector<float>(static_cast<size_t>(max_keypoints) * 256));
  g_sp = *out = sp;
  return SSB_OK;
}
void ssb_sp_destroy(ssb_superpoint* sp) {
  if (!sp) return;
  if (g_sp == sp) g_sp = nullptr;
  ++g_sp_destroyed;
  delete sp;
}
int ssb_sp_extract(ssb_superpoint* sp, const uint8_t* const* images, int batch, int height, int width, int row_stride,
                   int channels, float* const* xy, float* const* score, int* count, void** desc_dev, int* slot) {
  if (!sp || batch < 1 || height < 2 || width * channels < 3 || row_stride < width * channels) return SSB_ERR_INVALID;
  if (!g_extract_server && images[0][0] == 255) return SSB_ERR_CUDA;
  int rc = SSB_OK;
  for (int i = 0; i < batch; ++i) {
    const uint8_t* im = images[i];
    if (g_extract_server) {   // served mode: gray images only; rows land in the slot as fp16
      if (channels != 1) return SSB_ERR_INVALID;
      std::vector<unsigned short> rows(static_cast<size_t>(sp->max_kp) * 256);
      count[i] = g_extract_server(im, height, width, row_stride, sp->max_kp, xy[i], score[i], rows.data());
      if (sp->free_slots.empty()) {
        slot[i] = -1, desc_dev[i] = nullptr, rc = SSB_ERR_EXHAUSTED;
        continue;
      }
      const int s = sp->free_slots.back();
      sp->free_slots.pop_back();
      sp->refs[s] = 1;
      std::memcpy(sp->slot_mem[s].data(), rows.data(), sizeof(unsigned short) * 256 * count[i]);
      slot[i] = s, desc_dev[i] = sp->slot_mem[s].data();
      continue;
    }
    int n = 4 * im[0];
    if (n > sp->max_kp) n = sp->max_kp;
    count[i] = n;
    for (int k = 0; k < n; ++k) {
      xy[i][2 * k] = static_cast<float>(im[1]) + 2.0f * k;
      xy[i][2 * k + 1] = static_cast<float>(im[row_stride]) + static_cast<float>(k % 7);
      score[i][k] = 1.0f / (1.0f + k);
    }
    if (sp->free_slots.empty()) {
      slot[i] = -1, desc_dev[i] = nullptr, rc = SSB_ERR_EXHAUSTED;
      continue;
    }
    const int s = sp->free_slots.back();
    sp->free_slots.pop_back();
    sp->refs[s] = 1;
    float* d = sp->slot_mem[s].data();
    for (int k = 0; k < n; ++k)
      for (int c = 0; c < 256; ++c) d[k * 256 + c] = static_cast<float>(im[2]) + k + c / 1024.0f;
    slot[i] = s, desc_dev[i] = d;
  }
  return rc;
}
int ssb_sp_slot_retain(ssb_superpoint* sp, int slot) {
  if (!sp || slot < 0 || slot >= static_cast<int>(sp->refs.size()) || sp->refs[slot] <= 0) return SSB_ERR_INVALID;
  ++sp->refs[slot];
  return SSB_OK;
}
int ssb_sp_slot_release(ssb_superpoint* sp, int slot) {
  ++g_release_calls;
  if (!sp || slot < 0 || slot >= static_cast<int>(sp->refs.size()) || sp->refs[slot] <= 0) return SSB_ERR_INVALID;
  if (--sp->refs[slot] == 0) sp->free_slots.push_back(slot);
  return SSB_OK;
}
int ssb_sp_slots_in_use(ssb_superpoint* sp) { return sp ? static_cast<int>(sp->refs.size() - sp->free_slots.size()) : 0; }

int ssb_lg_create(const char* weights_path, int image_width, int image_height, int, int, ssb_lightglue** out) {
  *out = nullptr;
  if (!weights_path || std::string(weights_path) == "missing") return SSB_ERR_IO;
  *out = new ssb_lightglue{false, image_width, image_height};
  ++g_lg_alive;
  return SSB_OK;
}
int ssb_lg_clone_context(ssb_lightglue* src, int image_width, int image_height, ssb_lightglue** out) {
  *out = nullptr;
  if (!src) return SSB_ERR_INVALID;
  *out = new ssb_lightglue{true, image_width, image_height};
  ++g_lg_alive;
  return SSB_OK;
}
void ssb_lg_destroy(ssb_lightglue* lg) {
  if (!lg) return;
  --g_lg_alive;
  delete lg;
}
int ssb_lg_match_device(ssb_lightglue* lg, const float* xy0, int n0, const void* desc0_dev, const float* xy1, int n1,
                        const void* desc1_dev, int32_t* matches0, float* mscores0) {
  if (g_match_server) {
    if (!lg || !desc0_dev || !desc1_dev) return SSB_ERR_INVALID;
    g_match_server(xy0, n0, static_cast<const unsigned short*>(desc0_dev), xy1, n1,
                   static_cast<const unsigned short*>(desc1_dev), lg->width, lg->height, matches0, mscores0);
    return SSB_OK;
  }
  g_last.slot0 = slot_of(desc0_dev), g_last.slot1 = slot_of(desc1_dev);
  g_last.d0_first = g_last.d1_first = 0;
  return run_match(lg, 1, xy0, n0, xy1, n1, matches0, mscores0);
}
int ssb_lg_match_host(ssb_lightglue* lg, const float* xy0, int n0, const float* desc0_f32, const float* xy1, int n1,
                      const float* desc1_f32, int32_t* matches0, float* mscores0) {
  g_last.slot0 = g_last.slot1 = -2;
  g_last.d0_first = n0 > 0 ? desc0_f32[0] : 0, g_last.d1_first = n1 > 0 ? desc1_f32[0] : 0;
  return run_match(lg, 0, xy0, n0, xy1, n1, matches0, mscores0);
}
static float half_bits_to_float(unsigned short h) {
  const uint32_t sign = (h & 0x8000u) << 16, exp = (h >> 10) & 31u, man = h & 1023u;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else {   // subnormal: renormalise
      int e = -1;
      uint32_t m = man;
      do { ++e; m <<= 1; } while (!(m & 1024u));
      bits = sign | ((112 - e) << 23) | ((m & 1023u) << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7F800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}
int ssb_desc_to_host_f32(int, const void* desc_dev_f16, int count, int dim, float* out) {
  if (!desc_dev_f16 || count < 0 || dim != 256) return SSB_ERR_INVALID;
  if (g_extract_server) {   // served mode: the slot holds fp16 rows
    const unsigned short* h = static_cast<const unsigned short*>(desc_dev_f16);
    for (long i = 0; i < static_cast<long>(count) * dim; ++i) out[i] = half_bits_to_float(h[i]);
    return SSB_OK;
  }
  std::memcpy(out, desc_dev_f16, sizeof(float) * static_cast<size_t>(count) * dim);
  return SSB_OK;
}

// ---- EigenPlaces / rectifier / RGB-D doubles ----------------------------------------------------------------------
//   ep_compute: descriptor[c] = px(c % (width * channels)) of row 1 (checks row_stride) + 1, not normalised
//   ep_add / ep_query: a plain cosine index (double dot products, rows normalised on insertion) with the semantics of
//                      src/PlaceRecognizer.cc:21-52; the test compares with oracle/eigenplaces.py::CosineDescriptorIndex
//   rect_remap: out(y, x) = in(y % src_h, x % src_w) + 1;  rgbd_process: records every scalar argument,
//               out_xy = xy + 0.25, stereo = (x, depth(lround(y), lround(x)) / depth_factor, y), has = depth != 0
struct ssb_eigenplaces {
  int in_w = 0, in_h = 0;
  std::vector<uint64_t> ids;
  std::vector<std::vector<double>> rows;
};
struct ssb_rectifier {
  int dh = 0, dw = 0, sh = 0, sw = 0;
  float m00 = 0, m11 = 0;
};
struct ssb_rgbd {
  int max_kp = 0;
};
namespace {
double g_rgbd[32];
int g_rect_stride = 0;
std::vector<double> unit(const float* d, int dim) {
  double n = 0;
  for (int i = 0; i < dim; ++i) n += static_cast<double>(d[i]) * d[i];
  n = std::sqrt(n);
  std::vector<double> r(dim);
  for (int i = 0; i < dim; ++i) r[i] = n > 1e-12 ? d[i] / n : d[i];
  return r;
}
}  // namespace

int ssb_ep_create(const char* weights_path, int input_width, int input_height, int, int, ssb_eigenplaces** out) {
  *out = nullptr;
  if (!weights_path || std::string(weights_path) == "missing") return SSB_ERR_IO;
  *out = new ssb_eigenplaces;
  (*out)->in_w = input_width, (*out)->in_h = input_height;
  return SSB_OK;
}
void ssb_ep_destroy(ssb_eigenplaces* ep) { delete ep; }
int ssb_ep_descriptor_dim(ssb_eigenplaces*) { return 512; }
int ssb_ep_compute(ssb_eigenplaces* ep, const uint8_t* const* images, int count, int height, int width, int row_stride,
                   int channels, float* descriptors) {
  if (!ep || count < 1 || height < 2 || row_stride < width * channels) return SSB_ERR_INVALID;
  if (images[0][0] == 255) return SSB_ERR_CUDA;
  for (int i = 0; i < count; ++i)
    for (int c = 0; c < 512; ++c) descriptors[i * 512 + c] = images[i][row_stride + c % (width * channels)] + 1.0f;
  return SSB_OK;
}
int ssb_ep_add(ssb_eigenplaces* ep, uint64_t keyframe_id, const float* descriptor, int dim) {
  if (!ep || !descriptor || dim <= 0) return SSB_ERR_INVALID;
  ep->ids.push_back(keyframe_id);
  ep->rows.push_back(unit(descriptor, dim));
  return SSB_OK;
}
int ssb_ep_index_size(ssb_eigenplaces* ep) { return ep ? static_cast<int>(ep->ids.size()) : 0; }
int ssb_ep_query(ssb_eigenplaces* ep, const float* descriptor, int dim, uint64_t exclude_recent, int top_k,
                 float min_score, uint64_t* keyframe_ids, float* scores, int capacity, int* n_out) {
  *n_out = 0;
  if (!ep || !descriptor) return SSB_ERR_INVALID;
  const size_t M = ep->ids.size();
  if (M == 0 || M <= exclude_recent) return SSB_OK;
  const std::vector<double> q = unit(descriptor, dim);
  std::vector<std::pair<float, uint64_t>> c;
  for (size_t i = 0; i < M - exclude_recent; ++i) {
    double s = 0;
    for (int k = 0; k < dim; ++k) s += ep->rows[i][k] * q[k];
    if (static_cast<float>(s) >= min_score) c.emplace_back(static_cast<float>(s), ep->ids[i]);
  }
  std::stable_sort(c.begin(), c.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
  if (top_k > 0 && c.size() > static_cast<size_t>(top_k)) c.resize(top_k);
  for (size_t i = 0; i < c.size() && static_cast<int>(i) < capacity; ++i, ++*n_out)
    scores[i] = c[i].first, keyframe_ids[i] = c[i].second;
  return SSB_OK;
}

int ssb_rect_create(const float* map_x, const float* map_y, int dst_height, int dst_width, int src_height, int src_width,
                    int, int, ssb_rectifier** out) {
  *out = nullptr;
  if (!map_x || !map_y || (dst_height * dst_width) % 4) return SSB_ERR_INVALID;
  *out = new ssb_rectifier{dst_height, dst_width, src_height, src_width, map_x[0], map_y[dst_width + 1]};
  return SSB_OK;
}
void ssb_rect_destroy(ssb_rectifier* r) { delete r; }
int ssb_rect_remap(ssb_rectifier* r, const uint8_t* const* images, int count, int row_stride, uint8_t* const* out) {
  if (!r || row_stride < r->sw) return SSB_ERR_INVALID;
  g_rect_stride = row_stride;
  for (int i = 0; i < count; ++i)
    for (int y = 0; y < r->dh; ++y)
      for (int x = 0; x < r->dw; ++x)
        out[i][y * r->dw + x] = static_cast<uint8_t>(images[i][(y % r->sh) * row_stride + x % r->sw] + 1);
  return SSB_OK;
}

int ssb_rgbd_create(int max_keypoints, int, int, int, ssb_rgbd** out) {
  *out = new ssb_rgbd{max_keypoints};
  return SSB_OK;
}
void ssb_rgbd_destroy(ssb_rgbd* r) { delete r; }
int ssb_rgbd_process(ssb_rgbd* r, const float* xy, int n, const void* depth, int depth_type, int height, int width,
                     int row_stride, const double* camera, const double* dist, int n_dist, double bf, double depth_factor,
                     double max_depth, float* out_xy, double* out_stereo, uint8_t* out_has_depth) {
  if (!r || n > r->max_kp) return SSB_ERR_INVALID;
  double* g = g_rgbd;
  g[0] = n, g[1] = depth_type, g[2] = height, g[3] = width, g[4] = row_stride;
  for (int i = 0; i < 4; ++i) g[5 + i] = camera[i];
  g[9] = n_dist, g[10] = bf, g[11] = depth_factor, g[12] = max_depth, g[13] = dist != nullptr;
  for (int i = 0; i < 14; ++i) g[14 + i] = (dist && i < n_dist) ? dist[i] : -1;
  for (int i = 0; i < n; ++i) {
    const long u = std::lround(xy[2 * i]), v = std::lround(xy[2 * i + 1]);
    double z = 0;
    if (u >= 0 && v >= 0 && u < width && v < height) {
      const unsigned char* row = static_cast<const unsigned char*>(depth) + v * row_stride;
      z = depth_type == 0 ? reinterpret_cast<const uint16_t*>(row)[u] : reinterpret_cast<const float*>(row)[u];
    }
    out_xy[2 * i] = xy[2 * i] + 0.25f, out_xy[2 * i + 1] = xy[2 * i + 1] + 0.25f;
    out_stereo[3 * i] = xy[2 * i], out_stereo[3 * i + 1] = z / depth_factor, out_stereo[3 * i + 2] = xy[2 * i + 1];
    out_has_depth[i] = z != 0;
  }
  return SSB_OK;
}
void fake_last_rgbd(double* out28) { std::memcpy(out28, g_rgbd, 28 * sizeof(double)); }
int fake_rect_stride(void) { return g_rect_stride; }

void fake_set_servers(ExtractServer e, MatchServer m) { g_extract_server = e, g_match_server = m; }

// ---- inspection hooks for the test ----------------------------------------------------------------
int fake_slots_in_use(void) { return ssb_sp_slots_in_use(g_sp); }
int fake_release_calls(void) { return g_release_calls; }
int fake_sp_destroyed(void) { return g_sp_destroyed; }
int fake_lg_alive(void) { return g_lg_alive; }
void fake_last_match(int* info6, double* sums4) {
  info6[0] = g_last.clone, info6[1] = g_last.device_path, info6[2] = g_last.n0, info6[3] = g_last.n1;
  info6[4] = g_last.slot0, info6[5] = g_last.slot1;
  sums4[0] = g_last.xy0_sum, sums4[1] = g_last.xy1_sum, sums4[2] = g_last.d0_first, sums4[3] = g_last.d1_first;
}
}
