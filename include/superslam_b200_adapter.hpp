// Drop-in adapter: implements SuperSLAM's inference interfaces on top of the C-ABI.
//
// Compile this header inside the reference tree (it needs the reference's own
// include/InferenceInterfaces.h, include/DescriptorPool.h and OpenCV core) and link
// libsuperslam_b200.so.  It replaces the reference's TensorRT-backed classes one for one:
//
//   reference (include/SuperPoint.h:37-52, include/LightGlue.h:33-57)   this header
//   ------------------------------------------------------------------  ---------------------------
//   SuperPoint(engine, max_kp, thresh, borders) + initialize()          superslam_b200::SuperPointB200
//   SuperPoint::infer(image, keypoints, cv::Mat descriptors)            SuperPointB200::infer (host path of the demos)
//   LightGlue(engine, w, h) + initialize()                              superslam_b200::LightGlueB200
//   LightGlue(shared_engine(), w, h)                                    LightGlueB200(other, w, h)
//   EigenPlaces(engine, w, h) + initialize()  (include/EigenPlaces.h:21-66)  superslam_b200::EigenPlacesB200
//
// Error behaviour is the reference's: interface methods never throw; failure -> empty Features /
// MatchResult (src/SuperPoint.cc:895-899, src/LightGlue.cc:381-390); initialize() returns bool.
// Keypoints are built exactly like src/SuperPoint.cc:716 (size 1, angle -1, response = score) and
// matches like src/LightGlue.cc:352-361 (increasing queryIdx, distance = 1 - score).
#pragma once

#include <memory>
#include <string>
#include <utility>
#include <vector>

#include <cstdlib>

#include "InferenceInterfaces.h"  // reference header: superslam::IFeatureExtractor / IFeatureMatcher
#include "PlaceRecognizer.h"      // reference header: superslam::IPlaceRecognizer, LoopCandidate
#include "superslam_b200.h"

namespace superslam_b200 {

class SuperPointB200 : public superslam::IFeatureExtractor {
 public:
  SuperPointB200(std::string weights_file, int max_keypoints, double keypoint_threshold, int remove_borders,
                 int device_id = 0)
      : weights_(std::move(weights_file)), max_kp_(max_keypoints), thresh_(keypoint_threshold),
        borders_(remove_borders), device_(device_id) {}
  ~SuperPointB200() override = default;   // the extractor dies with its last live descriptor handle (see sp_)

  bool initialize() {
    ssb_superpoint* sp = nullptr;
    if (ssb_sp_create(weights_.c_str(), max_kp_, thresh_, borders_, /*num_slots=*/8, device_, &sp) != SSB_OK) return false;
    sp_ = std::shared_ptr<ssb_superpoint>(sp, [](ssb_superpoint* p) { ssb_sp_destroy(p); });
    return true;
  }

  superslam::Features extract(const cv::Mat& image) override {
    std::vector<superslam::Features> f = run({&image});
    return f.empty() ? superslam::Features{} : std::move(f[0]);
  }
  // SuperPoint::infer (include/SuperPoint.h:45-46, src/SuperPoint.cc:427-531): the host path the reference's demo
  // programs use (tests/test_superpoint_only.cc:71, tests/test_superpoint_cosine_matching.cc:192) - keypoints plus
  // L2-normalised CV_32F [n, 256] descriptors on the host.  false on a failed inference, true with zero keypoints.
  bool infer(const cv::Mat& image, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors) {
    keypoints.clear();
    descriptors = cv::Mat();
    std::vector<superslam::Features> f = run({&image});
    if (f.empty()) return false;
    keypoints = std::move(f[0].keypoints);
    const superslam::DeviceDescriptors& d = f[0].descriptors;
    if (keypoints.empty()) return true;
    if (d.empty()) return false;   // pool exhausted: the reference's host path has no pool, so report the failure
    descriptors.create(d.count, d.dim, CV_32F);
    return ssb_desc_to_host_f32(device_, d.data, d.count, d.dim, descriptors.ptr<float>()) == SSB_OK;
  }
  std::pair<superslam::Features, superslam::Features> extract_stereo(const cv::Mat& left,
                                                                    const cv::Mat& right) override {
    if (left.rows != right.rows || left.cols != right.cols)
      return {};  // "stereo pair must share resolution" (src/SuperPoint.cc:761-764)
    if (left.channels() != right.channels())   // the reference grays each side on its own (:768-773): two single passes
      return {extract(left), extract(right)};
    std::vector<superslam::Features> f = run({&left, &right});
    if (f.size() != 2) return {};
    return {std::move(f[0]), std::move(f[1])};
  }

 private:
  std::vector<superslam::Features> run(std::vector<const cv::Mat*> imgs) {
    std::vector<superslam::Features> out;
    if (!sp_ || imgs.empty() || imgs[0]->empty() || imgs[0]->depth() != CV_8U) return out;
    const int b = static_cast<int>(imgs.size());
    const int h = imgs[0]->rows, w = imgs[0]->cols, ch = imgs[0]->channels();
    std::vector<const uint8_t*> ptr(b);
    std::vector<cv::Mat> packed;   // the C-ABI takes one row stride per call: images whose steps differ are packed first
    size_t step = imgs[0]->step[0];
    for (int i = 1; i < b; ++i)
      if (imgs[i]->step[0] != step) packed.resize(b);
    for (int i = 0; i < b; ++i) {
      if (!packed.empty()) {
        packed[i] = imgs[i]->clone();
        step = packed[i].step[0];
      }
      ptr[i] = packed.empty() ? imgs[i]->data : packed[i].data;
    }
    std::vector<std::vector<float>> xy(b, std::vector<float>(2 * max_kp_)), sc(b, std::vector<float>(max_kp_));
    std::vector<float*> xyp(b), scp(b);
    for (int i = 0; i < b; ++i) xyp[i] = xy[i].data(), scp[i] = sc[i].data();
    std::vector<int> count(b, 0), slot(b, -1);
    std::vector<void*> desc(b, nullptr);
    const int st = ssb_sp_extract(sp_.get(), ptr.data(), b, h, w, static_cast<int>(step), ch, xyp.data(),
                                  scp.data(), count.data(), desc.data(), slot.data());
    if (st != SSB_OK && st != SSB_ERR_EXHAUSTED) return out;
    out.resize(b);
    for (int i = 0; i < b; ++i) {
      superslam::Features& f = out[i];
      f.keypoints.reserve(count[i]);
      for (int k = 0; k < count[i]; ++k)
        f.keypoints.emplace_back(xy[i][2 * k], xy[i][2 * k + 1], 1.0f, -1, sc[i][k]);
      f.descriptors.count = count[i];
      f.descriptors.dim = 256;
      f.descriptors.slot = slot[i];
      if (slot[i] >= 0) {  // DescriptorPool::make (include/DescriptorPool.h:62-76)
        f.descriptors.data = desc[i];
        // The deleter owns a share of the extractor: "a handle may outlive the pool" (include/DescriptorPool.h:71-75,
        // where the reference captures the shared FreeList, not `this`).  The slot memory and the free list live
        // inside ssb_superpoint, so it is destroyed only after the last Features handle has let go.
        std::shared_ptr<ssb_superpoint> sp = sp_;
        const int s = slot[i];
        f.descriptors.slot_ref = std::shared_ptr<void>(desc[i], [sp, s](void*) { ssb_sp_slot_release(sp.get(), s); });
      }
    }
    return out;
  }

  std::string weights_;
  int max_kp_;
  double thresh_;
  int borders_, device_;
  std::shared_ptr<ssb_superpoint> sp_;   // shared with every live descriptor handle
};

class LightGlueB200 : public superslam::IFeatureMatcher {
 public:
  LightGlueB200(std::string weights_file, int image_width, int image_height, int max_keypoints = 1024,
                int device_id = 0)
      : weights_(std::move(weights_file)), w_(image_width), h_(image_height), max_kp_(max_keypoints),
        device_(device_id) {}
  // Shared-weights context for a second thread (loop closure): src/SuperSLAM.cc:129-133.
  LightGlueB200(const LightGlueB200& primary, int image_width, int image_height)
      : w_(image_width), h_(image_height), max_kp_(primary.max_kp_), device_(primary.device_), shared_(primary.lg_) {}
  ~LightGlueB200() override { ssb_lg_destroy(lg_); }

  bool initialize() {
    if (shared_) return ssb_lg_clone_context(shared_, w_, h_, &lg_) == SSB_OK;
    return ssb_lg_create(weights_.c_str(), w_, h_, max_kp_, device_, &lg_) == SSB_OK;
  }

  MatchResult match(const std::vector<cv::KeyPoint>& kp0, const cv::Mat& d0, const std::vector<cv::KeyPoint>& kp1,
                    const cv::Mat& d1) override {
    MatchResult r;
    if (!lg_ || kp0.empty() || kp1.empty()) return r;
    cv::Mat a = d0, b = d1;
    if (a.type() != CV_32F) d0.convertTo(a, CV_32F);
    if (b.type() != CV_32F) d1.convertTo(b, CV_32F);
    if (!a.isContinuous()) a = a.clone();
    if (!b.isContinuous()) b = b.clone();
    const std::vector<float> x0 = flatten(kp0), x1 = flatten(kp1);
    std::vector<int32_t> m(kp0.size());
    std::vector<float> s(kp0.size());
    if (ssb_lg_match_host(lg_, x0.data(), static_cast<int>(kp0.size()), a.ptr<float>(), x1.data(),
                          static_cast<int>(kp1.size()), b.ptr<float>(), m.data(), s.data()) != SSB_OK)
      return r;
    fill(m, s, &r);
    return r;
  }
  MatchResult match(const std::vector<cv::KeyPoint>& kp0, const superslam::DeviceDescriptors& d0,
                    const std::vector<cv::KeyPoint>& kp1, const superslam::DeviceDescriptors& d1) override {
    MatchResult r;
    if (!lg_ || d0.empty() || d1.empty() || kp0.empty() || kp1.empty()) return r;
    const std::vector<float> x0 = flatten(kp0), x1 = flatten(kp1);
    std::vector<int32_t> m(kp0.size());
    std::vector<float> s(kp0.size());
    if (ssb_lg_match_device(lg_, x0.data(), static_cast<int>(kp0.size()), d0.data, x1.data(),
                            static_cast<int>(kp1.size()), d1.data, m.data(), s.data()) != SSB_OK)
      return r;
    fill(m, s, &r);
    return r;
  }
  cv::Mat descriptors_to_host(const superslam::DeviceDescriptors& d) override {
    if (d.empty()) return cv::Mat();
    cv::Mat out(d.count, d.dim, CV_32F);
    if (ssb_desc_to_host_f32(device_, d.data, d.count, d.dim, out.ptr<float>()) != SSB_OK) return cv::Mat();
    return out;
  }

 private:
  static std::vector<float> flatten(const std::vector<cv::KeyPoint>& kp) {
    std::vector<float> v(kp.size() * 2);
    for (size_t i = 0; i < kp.size(); ++i) v[2 * i] = kp[i].pt.x, v[2 * i + 1] = kp[i].pt.y;
    return v;
  }
  static void fill(const std::vector<int32_t>& m, const std::vector<float>& s, MatchResult* r) {
    for (int i = 0; i < static_cast<int>(m.size()); ++i) {
      if (m[i] < 0) continue;  // unmatched query keypoint (src/LightGlue.cc:353-355)
      cv::DMatch dm;
      dm.queryIdx = i;
      dm.trainIdx = m[i];
      dm.distance = 1.0f - s[i];
      r->matches.push_back(dm);
    }
  }
  std::string weights_;
  int w_, h_, max_kp_, device_;
  ssb_lightglue* shared_ = nullptr;
  ssb_lightglue* lg_ = nullptr;
};

// superslam::IPlaceRecognizer (include/PlaceRecognizer.h:20-36) over ssb_ep_*: the global-descriptor network
// and the cosine index both live on the GPU.  Runs on the loop worker thread like the reference's class.
class EigenPlacesB200 : public superslam::IPlaceRecognizer {
 public:
  EigenPlacesB200(std::string weights_file, int input_width, int input_height, int device_id = 0)
      : weights_(std::move(weights_file)), w_(input_width), h_(input_height), device_(device_id) {
    if (const char* s = std::getenv("SUPERSLAM_LOOP_MIN_SCORE"))   // src/EigenPlaces.cc:30-34
      min_score_ = static_cast<float>(std::atof(s));
  }
  ~EigenPlacesB200() override { ssb_ep_destroy(ep_); }

  bool initialize() { return ssb_ep_create(weights_.c_str(), w_, h_, /*max_batch=*/1, device_, &ep_) == SSB_OK; }

  cv::Mat compute_global_descriptor(const cv::Mat& image) override {
    if (!ep_ || image.empty() || image.depth() != CV_8U) return cv::Mat();   // src/EigenPlaces.cc:146-147
    const uint8_t* ptr = image.data;
    cv::Mat desc(1, ssb_ep_descriptor_dim(ep_), CV_32F);
    if (ssb_ep_compute(ep_, &ptr, 1, image.rows, image.cols, static_cast<int>(image.step[0]), image.channels(),
                       desc.ptr<float>()) != SSB_OK)
      return cv::Mat();
    return desc;
  }
  void add(size_t keyframe_id, const cv::Mat& global_descriptor) override {
    if (global_descriptor.empty()) return;   // a failed compute_global_descriptor: nothing to index
    cv::Mat row = (global_descriptor.isContinuous() ? global_descriptor : global_descriptor.clone()).reshape(1, 1);
    if (row.type() != CV_32F) row.convertTo(row, CV_32F);
    if (!row.isContinuous()) row = row.clone();
    ssb_ep_add(ep_, keyframe_id, row.ptr<float>(), row.cols);
  }
  std::vector<superslam::LoopCandidate> query(const cv::Mat& global_descriptor, size_t excludeRecent,
                                              int topK) override {
    std::vector<superslam::LoopCandidate> out;
    if (global_descriptor.empty()) return out;
    cv::Mat row = (global_descriptor.isContinuous() ? global_descriptor : global_descriptor.clone()).reshape(1, 1);
    if (row.type() != CV_32F) row.convertTo(row, CV_32F);
    if (!row.isContinuous()) row = row.clone();
    const int cap = topK > 0 ? topK : ssb_ep_index_size(ep_);
    if (cap <= 0) return out;
    std::vector<uint64_t> ids(cap);
    std::vector<float> sc(cap);
    int n = 0;
    if (ssb_ep_query(ep_, row.ptr<float>(), row.cols, excludeRecent, topK, min_score_, ids.data(), sc.data(), cap,
                     &n) != SSB_OK)
      return out;
    for (int i = 0; i < n; ++i) out.push_back({static_cast<size_t>(ids[i]), sc[i]});
    return out;
  }

 private:
  std::string weights_;
  int w_, h_, device_;
  float min_score_ = 0.75f;   // include/EigenPlaces.h:55
  ssb_eigenplaces* ep_ = nullptr;
};

// cv::remap(img, out, M1, M2, cv::INTER_LINEAR) on the device (examples/stereo/euroc.cc:176-177): construct once per
// camera with the CV_32F maps of cv::initUndistortRectifyMap (euroc.cc:118-133), then
//   rect_l(imLeft, imLeftRect);   // instead of cv::remap(imLeft, imLeftRect, M1l, M2l, cv::INTER_LINEAR)
class RemapB200 {
 public:
  RemapB200(const cv::Mat& map_x, const cv::Mat& map_y, cv::Size src_size, int device_id = 0) {
    cv::Mat mx = map_x.isContinuous() ? map_x : map_x.clone(), my = map_y.isContinuous() ? map_y : map_y.clone();
    if (mx.type() == CV_32FC1 && my.type() == CV_32FC1 && mx.size() == my.size())
      ssb_rect_create(mx.ptr<float>(), my.ptr<float>(), mx.rows, mx.cols, src_size.height, src_size.width, 2,
                      device_id, &r_);
    dst_ = mx.size();
  }
  ~RemapB200() { ssb_rect_destroy(r_); }
  RemapB200(const RemapB200&) = delete;
  RemapB200& operator=(const RemapB200&) = delete;
  bool operator()(const cv::Mat& src, cv::Mat& dst) const {
    if (!r_ || src.type() != CV_8UC1) return false;
    dst.create(dst_, CV_8UC1);
    const uint8_t* in = src.data;
    uint8_t* out = dst.data;   // freshly created: continuous
    return ssb_rect_remap(r_, &in, 1, static_cast<int>(src.step[0]), &out) == SSB_OK;
  }

 private:
  ssb_rectifier* r_ = nullptr;
  cv::Size dst_;
};

// The part of RgbdFrontEnd::process after ext_->extract (src/RgbdFrontEnd.cc:27-58) as one device call;
// a maintainer replaces that loop by
//   post.run(raw, depth, K_.fx(), K_.fy(), K_.px(), K_.py(), dist_coeffs_, K_.fx() * K_.baseline(),
//            depth_factor_, max_depth_, undist, stereo, has_depth);
class RgbdPostB200 {
 public:
  explicit RgbdPostB200(int max_keypoints, cv::Size max_size, int device_id = 0) {
    ssb_rgbd_create(max_keypoints, max_size.height, max_size.width, device_id, &r_);
  }
  ~RgbdPostB200() { ssb_rgbd_destroy(r_); }
  RgbdPostB200(const RgbdPostB200&) = delete;
  RgbdPostB200& operator=(const RgbdPostB200&) = delete;
  // stereo: n x (uL, uR or NaN, v) doubles - the fields of gtsam::StereoPoint2
  bool run(const std::vector<cv::Point2f>& raw, const cv::Mat& depth, double fx, double fy, double cx, double cy,
           const cv::Mat& dist_coeffs, double bf, double depth_factor, double max_depth,
           std::vector<cv::Point2f>& undist, std::vector<double>& stereo, std::vector<char>& has_depth) const {
    const int n = static_cast<int>(raw.size());
    undist.resize(n);
    stereo.resize(static_cast<size_t>(n) * 3);
    has_depth.assign(n, 0);
    if (!r_ || n == 0) return r_ != nullptr;
    const int type = depth.type() == CV_16U ? 0 : (depth.type() == CV_32F ? 1 : -1);
    cv::Mat zero;
    const cv::Mat* dm = &depth;
    if (type < 0 || depth.empty()) {   // sampleDepth returns 0 for any other type and outside the map
      zero = cv::Mat::zeros(depth.empty() ? cv::Size(1, 1) : depth.size(), CV_16U);   // (src/RgbdFrontEnd.cc:12-20)
      dm = &zero;
    }
    cv::Mat d64;   // (cv::Mat::reshape throws on an empty matrix: no distortion model -> no coefficients)
    if (!dist_coeffs.empty()) {
      const cv::Mat dc = dist_coeffs.isContinuous() ? dist_coeffs : dist_coeffs.clone();
      dc.reshape(1, 1).convertTo(d64, CV_64F);
    }
    const double cam[4] = {fx, fy, cx, cy};
    return ssb_rgbd_process(r_, reinterpret_cast<const float*>(raw.data()), n, dm->data, dm == &zero ? 0 : type, dm->rows,
                            dm->cols, static_cast<int>(dm->step[0]), cam, d64.empty() ? nullptr : d64.ptr<double>(),
                            d64.empty() ? 0 : d64.cols, bf, depth_factor, max_depth,
                            reinterpret_cast<float*>(undist.data()), stereo.data(),
                            reinterpret_cast<uint8_t*>(has_depth.data())) == SSB_OK;
  }

 private:
  ssb_rgbd* r_ = nullptr;
};

}  // namespace superslam_b200
