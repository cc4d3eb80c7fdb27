// CPU test double of the ssb_fe_* entry points the multi-device driver (superslam_b200/csrc/multigpu.cpp) calls:
// every output is a simple function of the image bytes and of the device the front end was created on, so
// tests/test_multigpu_driver.py can check sharding, step walk and scatter without a GPU.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../include/superslam_b200.h"

struct Step {
  int pairs;
  std::vector<int> count;
  std::vector<float> xy, score, mscores, ur;
  std::vector<int32_t> matches;
  std::vector<uint8_t> hd;
};
struct ssb_frontend {
  int K, max_pairs, device;
  std::deque<Step> inflight;
  int max_inflight_seen = 0, submits = 0;
};
static thread_local std::string g_err;
static int g_fail_device = -1;

extern "C" {
const char* ssb_last_error(void) { return g_err.c_str(); }
void fake_fe_fail_on_device(int d) { g_fail_device = d; }
int ssb_fe_create(const char*, const char*, int max_keypoints, double, int, int, int, float, int max_pairs, int device_id,
                  ssb_frontend** out) {
  *out = new ssb_frontend{max_keypoints, max_pairs, device_id, {}, 0, 0};
  return SSB_OK;
}
void ssb_fe_destroy(ssb_frontend* fe) { delete fe; }
int ssb_fe_submit(ssb_frontend* fe, const uint8_t* const* images, int pairs, int h, int w, int row_stride) {
  if (pairs < 1 || pairs > fe->max_pairs) {
    g_err = "fake: pairs exceeds capacity";
    return SSB_ERR_INVALID;
  }
  if (fe->inflight.size() >= 2) {
    g_err = "fake: two steps already in flight";
    return SSB_ERR_INVALID;
  }
  if (fe->device == g_fail_device) {
    g_err = "fake: injected failure";
    return SSB_ERR_CUDA;
  }
  const size_t K = fe->K;
  Step s;
  s.pairs = pairs;
  s.count.resize(2 * pairs), s.xy.resize(2 * pairs * K * 2), s.score.resize(2 * pairs * K);
  s.matches.resize(pairs * K), s.mscores.resize(pairs * K), s.ur.resize(pairs * K), s.hd.resize(pairs * K);
  for (int i = 0; i < 2 * pairs; ++i) {
    unsigned sum = 0;
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) sum += images[i][static_cast<size_t>(y) * row_stride + x];
    s.count[i] = static_cast<int>(sum % (K + 1));
    for (size_t k = 0; k < K; ++k) {
      s.xy[(i * K + k) * 2] = static_cast<float>(sum % 1000) + k;
      s.xy[(i * K + k) * 2 + 1] = static_cast<float>(i & 1);
      s.score[i * K + k] = static_cast<float>(sum) * 0.5f - k;
    }
    if ((i & 1) == 0) {
      const size_t p = i / 2;
      for (size_t k = 0; k < K; ++k) {
        s.matches[p * K + k] = static_cast<int32_t>((sum + k) % 97) - 1;
        s.mscores[p * K + k] = static_cast<float>((sum + 3 * k) % 11) / 11.0f;
        s.ur[p * K + k] = static_cast<float>(fe->device);        // which device served this pair
        s.hd[p * K + k] = static_cast<uint8_t>((sum + k) & 1);
      }
    }
  }
  fe->inflight.push_back(std::move(s));
  ++fe->submits;
  if (static_cast<int>(fe->inflight.size()) > fe->max_inflight_seen) fe->max_inflight_seen = fe->inflight.size();
  return SSB_OK;
}
int ssb_fe_collect(ssb_frontend* fe, int* pairs, int* count, float* xy, float* score, int32_t* matches0, float* mscores0,
                   float* ur, uint8_t* hd) {
  if (fe->inflight.empty()) {
    g_err = "fake: nothing in flight";
    return SSB_ERR_INVALID;
  }
  Step s = std::move(fe->inflight.front());
  fe->inflight.pop_front();
  if (pairs) *pairs = s.pairs;
  if (count) std::memcpy(count, s.count.data(), s.count.size() * 4);
  if (xy) std::memcpy(xy, s.xy.data(), s.xy.size() * 4);
  if (score) std::memcpy(score, s.score.data(), s.score.size() * 4);
  if (matches0) std::memcpy(matches0, s.matches.data(), s.matches.size() * 4);
  if (mscores0) std::memcpy(mscores0, s.mscores.data(), s.mscores.size() * 4);
  if (ur) std::memcpy(ur, s.ur.data(), s.ur.size() * 4);
  if (hd) std::memcpy(hd, s.hd.data(), s.hd.size());
  return SSB_OK;
}
}
