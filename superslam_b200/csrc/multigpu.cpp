// Multi-device driver of the frame-pair front end (include/superslam_b200.h, ssb_mg_*).
//
// Stereo pairs are independent, so the reference's single front end (src/SuperSLAM.cc:82-86,107: one SuperPoint, one
// LightGlue, one StereoFrontEnd) scales to the GPUs of a box without any data-path exchange: one host thread and one
// ssb_frontend per device, pair p of a call goes to device p mod G (round-robin keeps the per-pair latency of a
// stream balanced), every device walks its share in steps of at most max_pairs_per_device pairs through the
// streaming calls (ssb_fe_submit / ssb_fe_collect: the upload of step i+1 runs under the kernels of step i), and the
// "result gather" is each worker writing its pairs' rows straight into the caller's arrays.
//
// This file deliberately uses nothing but the public C-ABI (superslam_b200.h): it is the code a maintainer would
// otherwise write above the library, and tests/test_multigpu_driver.py links it against a CPU test double of
// ssb_fe_* to check the sharding, step walk and scatter without a device.
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/superslam_b200.h"

namespace {

struct Job {
  const uint8_t* const* images = nullptr;
  int pairs = 0, h = 0, w = 0, row_stride = 0;
  int* count = nullptr;
  float* xy = nullptr;
  float* score = nullptr;
  int32_t* matches0 = nullptr;
  float* mscores0 = nullptr;
  float* ur = nullptr;
  uint8_t* hd = nullptr;
};

struct Worker {
  int index = 0, device = 0;
  ssb_frontend* fe = nullptr;
  std::thread thread;
  // scratch of one step (max_pairs pairs), the shape ssb_fe_collect writes
  std::vector<int> count;
  std::vector<float> xy, score, mscores, ur;
  std::vector<int32_t> matches;
  std::vector<uint8_t> hd;
  int status = SSB_OK;
  std::string error;
};

}  // namespace

struct ssb_multigpu {
  int K = 0, max_pairs = 0;
  std::vector<std::unique_ptr<Worker>> workers;
  std::mutex mu;
  std::condition_variable cv_start, cv_done;
  unsigned long long generation = 0;   // bumped per job
  int pending = 0;
  bool quit = false;
  Job job;
  std::string error;

  // worker w owns pairs w, w + G, w + 2G, ... of the call
  void run_share(Worker& wk, const Job& j) {
    const int G = static_cast<int>(workers.size());
    std::vector<int> mine;
    for (int p = wk.index; p < j.pairs; p += G) mine.push_back(p);
    wk.status = SSB_OK;
    if (mine.empty()) return;
    const size_t Ks = static_cast<size_t>(K);
    const int steps = (static_cast<int>(mine.size()) + max_pairs - 1) / max_pairs;
    std::vector<std::vector<const uint8_t*>> ptrs(steps);
    auto submit = [&](int s) -> int {
      const int first = s * max_pairs, n = std::min(max_pairs, static_cast<int>(mine.size()) - first);
      ptrs[s].resize(2 * n);
      for (int i = 0; i < n; ++i) {
        ptrs[s][2 * i] = j.images[2 * mine[first + i]];
        ptrs[s][2 * i + 1] = j.images[2 * mine[first + i] + 1];
      }
      return ssb_fe_submit(wk.fe, ptrs[s].data(), n, j.h, j.w, j.row_stride);
    };
    auto collect = [&](int s) -> int {
      const int first = s * max_pairs;
      int n = 0;
      const int st = ssb_fe_collect(wk.fe, &n, wk.count.data(), wk.xy.data(), wk.score.data(), wk.matches.data(),
                                    wk.mscores.data(), wk.ur.data(), wk.hd.data());
      if (st != SSB_OK) return st;
      for (int i = 0; i < n; ++i) {   // scatter the step's rows to the pairs' places in the caller's arrays
        const size_t g = static_cast<size_t>(mine[first + i]), l = static_cast<size_t>(i);
        if (j.count) j.count[2 * g] = wk.count[2 * l], j.count[2 * g + 1] = wk.count[2 * l + 1];
        if (j.xy) std::memcpy(j.xy + 2 * g * Ks * 2, wk.xy.data() + 2 * l * Ks * 2, 2 * Ks * 2 * sizeof(float));
        if (j.score) std::memcpy(j.score + 2 * g * Ks, wk.score.data() + 2 * l * Ks, 2 * Ks * sizeof(float));
        if (j.matches0) std::memcpy(j.matches0 + g * Ks, wk.matches.data() + l * Ks, Ks * sizeof(int32_t));
        if (j.mscores0) std::memcpy(j.mscores0 + g * Ks, wk.mscores.data() + l * Ks, Ks * sizeof(float));
        if (j.ur) std::memcpy(j.ur + g * Ks, wk.ur.data() + l * Ks, Ks * sizeof(float));
        if (j.hd) std::memcpy(j.hd + g * Ks, wk.hd.data() + l * Ks, Ks);
      }
      return SSB_OK;
    };
    int st = submit(0);
    int in_flight = st == SSB_OK ? 1 : 0;
    for (int s = 0; s < steps && st == SSB_OK; ++s) {
      if (s + 1 < steps) {
        st = submit(s + 1);
        if (st != SSB_OK) break;
        ++in_flight;
      }
      st = collect(s);
      --in_flight;
    }
    if (st != SSB_OK) {
      const char* e = ssb_last_error();   // thread-local in the library: read it on the worker's own thread
      wk.error = e ? e : "";
      int n = 0;
      while (in_flight-- > 0) ssb_fe_collect(wk.fe, &n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    }
    wk.status = st;
  }

  void worker_loop(Worker* wk) {
    unsigned long long seen = 0;
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_start.wait(lk, [&] { return quit || generation != seen; });
        if (quit) return;
        seen = generation;
        j = job;
      }
      run_share(*wk, j);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--pending == 0) cv_done.notify_all();
      }
    }
  }
};

extern "C" {

int ssb_mg_create(const char* sp_weights, const char* lg_weights, int max_keypoints, double keypoint_threshold,
                  int remove_borders, int lg_image_width, int lg_image_height, float min_disparity,
                  int max_pairs_per_device, const int* device_ids, int n_devices, ssb_multigpu** out) {
  if (out == nullptr) return SSB_ERR_INVALID;
  *out = nullptr;
  if (n_devices < 1 || n_devices > 64 || max_pairs_per_device < 1 || max_keypoints < 1) return SSB_ERR_INVALID;
  try {
    std::unique_ptr<ssb_multigpu> mg(new ssb_multigpu);
    mg->K = max_keypoints;
    mg->max_pairs = max_pairs_per_device;
    const size_t P = static_cast<size_t>(max_pairs_per_device), K = static_cast<size_t>(max_keypoints);
    int st = SSB_OK;
    for (int i = 0; i < n_devices && st == SSB_OK; ++i) {
      std::unique_ptr<Worker> w(new Worker);
      w->index = i;
      w->device = device_ids ? device_ids[i] : i;
      st = ssb_fe_create(sp_weights, lg_weights, max_keypoints, keypoint_threshold, remove_borders, lg_image_width,
                         lg_image_height, min_disparity, max_pairs_per_device, w->device, &w->fe);
      w->count.resize(2 * P), w->xy.resize(2 * P * K * 2), w->score.resize(2 * P * K);
      w->matches.resize(P * K), w->mscores.resize(P * K), w->ur.resize(P * K), w->hd.resize(P * K);
      mg->workers.push_back(std::move(w));
    }
    if (st != SSB_OK) {   // ssb_last_error() already names the device that failed
      for (auto& w : mg->workers)
        if (w->fe) ssb_fe_destroy(w->fe);
      return st;
    }
    for (auto& w : mg->workers) w->thread = std::thread(&ssb_multigpu::worker_loop, mg.get(), w.get());
    *out = mg.release();
    return SSB_OK;
  } catch (...) {
    return SSB_ERR_INVALID;
  }
}

void ssb_mg_destroy(ssb_multigpu* mg) {
  if (mg == nullptr) return;
  {
    std::lock_guard<std::mutex> lk(mg->mu);
    mg->quit = true;
  }
  mg->cv_start.notify_all();
  for (auto& w : mg->workers) {
    if (w->thread.joinable()) w->thread.join();
    if (w->fe) ssb_fe_destroy(w->fe);
  }
  delete mg;
}

int ssb_mg_device_count(ssb_multigpu* mg) { return mg ? static_cast<int>(mg->workers.size()) : -1; }
const char* ssb_mg_last_error(ssb_multigpu* mg) { return mg ? mg->error.c_str() : "mg is null"; }
int ssb_mg_device_of_pair(ssb_multigpu* mg, int pair) {
  if (mg == nullptr || pair < 0) return -1;
  return mg->workers[static_cast<size_t>(pair) % mg->workers.size()]->device;
}

int ssb_mg_process(ssb_multigpu* mg, const uint8_t* const* images, int pairs, int height, int width, int row_stride,
                   int* count, float* xy, float* score, int32_t* matches0, float* mscores0, float* stereo_ur,
                   uint8_t* has_depth) {
  if (mg == nullptr) return SSB_ERR_INVALID;
  if (images == nullptr || pairs < 0 || row_stride < width) {
    mg->error = "bad image arguments";
    return SSB_ERR_INVALID;
  }
  if (pairs == 0) return SSB_OK;
  try {
    std::unique_lock<std::mutex> lk(mg->mu);
    mg->job = Job{images, pairs, height, width, row_stride, count, xy, score, matches0, mscores0, stereo_ur, has_depth};
    mg->pending = static_cast<int>(mg->workers.size());
    ++mg->generation;
    mg->cv_start.notify_all();
    mg->cv_done.wait(lk, [&] { return mg->pending == 0; });
    for (auto& w : mg->workers) {
      if (w->status != SSB_OK) {
        mg->error = "device " + std::to_string(w->device) + ": " + w->error;
        return w->status;
      }
    }
    return SSB_OK;
  } catch (...) {
    mg->error = "unexpected C++ exception";
    return SSB_ERR_INVALID;
  }
}

}  // extern "C"
