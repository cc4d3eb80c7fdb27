// Drop-in harness: the REFERENCE's own caller, StereoFrontEnd::process (/root/reference/src/StereoFrontEnd.cc:10-49,
// compiled from the source where it lies), running on top of the PRODUCT's adapter classes
// (include/superslam_b200_adapter.hpp: SuperPointB200 / LightGlueB200 behind superslam::IFeatureExtractor /
// IFeatureMatcher) - i.e. exactly the object graph src/SuperSLAM.cc:107-138 builds, with the TensorRT classes swapped
// for the adapter.  It also replays the two other call shapes the reference makes through the matcher interface:
//   dropin_track            VoEstimator::track        last keyframe <-> frame, device descriptors (src/VoEstimator.cc:240-246)
//   dropin_promote_keyframe make_keyframe_msg         descriptors_to_host                          (src/VoEstimator.cc:106)
//   dropin_verify           LoopCloser::verify        host descriptors on a cloned context         (src/LoopCloser.cc:44-53,
//                                                                                                   src/SuperSLAM.cc:129-133)
// OpenCV / GTSAM value types come from the functional stand-ins in oracle/stubs_cv and oracle/stubs.
// oracle/Makefile links it twice into oracle/_ref/: against the real libsuperslam_b200.so (libdropin.so, GPU parity
// test + the no-GPU error path) and against oracle/fake_capi.cpp (libdropin_fake.so: the adapter's marshalling,
// slot ref-counting and error mapping checked on the CPU).  TEST INFRASTRUCTURE.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "StereoFrontEnd.h"
#include "superslam_b200_adapter.hpp"

namespace {
struct Harness {
  std::unique_ptr<superslam_b200::SuperPointB200> sp;
  std::unique_ptr<superslam_b200::LightGlueB200> lg, loop_lg;
  std::unique_ptr<superslam::StereoFrontEnd> fe;
  bool sp_ok = false, lg_ok = false, loop_ok = false;
  superslam::StereoFrame frame;          // the frame process() returned last
  superslam::StereoFrame last_keyframe;  // VoEstimator::last_keyframe_ (src/VoEstimator.cc:327): keeps its slot alive
  cv::Mat keyframe_desc;                 // KeyframeRecord::descriptors_left of the last promoted keyframe
  std::vector<cv::KeyPoint> keyframe_kp;
  std::vector<superslam::StereoFrame> held;  // further live copies (a keyframe window), to drive the pool to exhaustion
};

int copy_matches(const MatchResult& m, int cap, int* query, int* train, float* distance) {
  const int n = static_cast<int>(m.matches.size());
  for (int i = 0; i < n && i < cap; ++i) {
    query[i] = m.matches[i].queryIdx;
    train[i] = m.matches[i].trainIdx;
    distance[i] = m.matches[i].distance;
  }
  return n;
}
}  // namespace

extern "C" {

// SuperSLAM::SuperSLAM (src/SuperSLAM.cc:107-138): extractor, tracking matcher, loop matcher sharing its weights,
// stereo front end over raw interface pointers.  Returns a handle even when an initialize() fails (the reference's
// facade logs and carries on with a broken object): status bits 1 = SuperPoint, 2 = LightGlue, 4 = loop matcher.
void* dropin_create(const char* sp_weights, const char* lg_weights, int image_width, int image_height, int max_keypoints,
                    double keypoint_threshold, int remove_borders, float min_disparity, int* status) {
  Harness* h = new Harness;
  h->sp.reset(new superslam_b200::SuperPointB200(sp_weights, max_keypoints, keypoint_threshold, remove_borders));
  h->lg.reset(new superslam_b200::LightGlueB200(lg_weights, image_width, image_height, max_keypoints));
  h->sp_ok = h->sp->initialize();
  h->lg_ok = h->lg->initialize();
  h->loop_lg.reset(new superslam_b200::LightGlueB200(*h->lg, image_width, image_height));
  h->loop_ok = h->lg_ok && h->loop_lg->initialize();
  superslam::IFeatureExtractor* ext = h->sp.get();
  superslam::IFeatureMatcher* mat = h->lg.get();
  h->fe.reset(new superslam::StereoFrontEnd(ext, mat, gtsam::Cal3_S2Stereo(500, 500, 0, image_width / 2.0,
                                                                           image_height / 2.0, 0.1), min_disparity));
  if (status) *status = (h->sp_ok ? 1 : 0) | (h->lg_ok ? 2 : 0) | (h->loop_ok ? 4 : 0);
  return h;
}

void dropin_destroy(void* handle) { delete static_cast<Harness*>(handle); }

// StereoFrontEnd::process(left, right, timestamp).  Images are cv::Mat headers over the caller's rows (channels 1 or
// 3).  Outputs sized `cap`: keypoints_left (x, y), response, stereo (uL, uR, v), has_depth; desc = {count, dim, slot,
// data != nullptr}.  Returns the number of left keypoints.
int dropin_process(void* handle, const uint8_t* left, const uint8_t* right, int height, int width, int row_stride,
                   int row_stride_right, int channels, double timestamp, int cap, float* xy, float* response, float* size_angle,
                   double* stereo, char* has_depth, int* desc) {
  Harness* h = static_cast<Harness*>(handle);
  // channels: low 4 bits = left (and right); bits 4.. = right's channel count when it differs
  const int ch_l = channels & 15, ch_r = (channels >> 4) ? (channels >> 4) : ch_l;
  const cv::Mat l(height, width, CV_MAKETYPE(CV_8U, ch_l), const_cast<uint8_t*>(left), row_stride);
  const cv::Mat r(height, width, CV_MAKETYPE(CV_8U, ch_r), const_cast<uint8_t*>(right), row_stride_right);
  h->frame = h->fe->process(l, r, timestamp);
  const superslam::StereoFrame& f = h->frame;
  const int n = static_cast<int>(f.keypoints_left.size());
  for (int i = 0; i < n && i < cap; ++i) {
    const cv::KeyPoint& k = f.keypoints_left[i];
    xy[2 * i] = k.pt.x, xy[2 * i + 1] = k.pt.y;
    response[i] = k.response;
    size_angle[2 * i] = k.size, size_angle[2 * i + 1] = k.angle;
    stereo[3 * i] = f.stereo[i].uL(), stereo[3 * i + 1] = f.stereo[i].uR(), stereo[3 * i + 2] = f.stereo[i].v();
    has_depth[i] = f.has_depth[i];
  }
  desc[0] = f.descriptors_left.count, desc[1] = f.descriptors_left.dim, desc[2] = f.descriptors_left.slot;
  desc[3] = f.descriptors_left.data != nullptr;
  return n;
}

// VoEstimator: `last_keyframe_ = frame` (src/VoEstimator.cc:327) - the copy shares the descriptor slot - and
// make_keyframe_msg's descriptors_to_host (:106).  Copies the host descriptors [count, dim] to `out` when given;
// returns the row count (0 for an empty handle -> empty Mat).
int dropin_promote_keyframe(void* handle, float* out, int cap_rows) {
  Harness* h = static_cast<Harness*>(handle);
  h->last_keyframe = h->frame;
  superslam::IFeatureMatcher* mat = h->lg.get();
  h->keyframe_desc = mat->descriptors_to_host(h->last_keyframe.descriptors_left);
  h->keyframe_kp = h->last_keyframe.keypoints_left;
  if (h->keyframe_desc.empty()) return 0;
  if (h->keyframe_desc.type() != CV_32F || !h->keyframe_desc.isContinuous()) return -1;
  const int rows = h->keyframe_desc.rows;
  if (out && rows <= cap_rows)
    std::memcpy(out, h->keyframe_desc.ptr<float>(), sizeof(float) * static_cast<size_t>(rows) * h->keyframe_desc.cols);
  return rows;
}

// VoEstimator::track's match: queryIdx = last keyframe, trainIdx = current frame, device descriptors on both sides.
int dropin_track(void* handle, int cap, int* query, int* train, float* distance) {
  Harness* h = static_cast<Harness*>(handle);
  superslam::IFeatureMatcher* mat = h->lg.get();
  const MatchResult m = mat->match(h->last_keyframe.keypoints_left, h->last_keyframe.descriptors_left,
                                   h->frame.keypoints_left, h->frame.descriptors_left);
  return copy_matches(m, cap, query, train, distance);
}

// LoopCloser::verify's match on the loop worker's own matcher: host descriptors of the stored keyframe (candidate)
// against host descriptors of the current frame (query).
int dropin_verify(void* handle, int cap, int* query, int* train, float* distance) {
  Harness* h = static_cast<Harness*>(handle);
  superslam::IFeatureMatcher* mat = h->loop_lg.get();
  const cv::Mat cur = mat->descriptors_to_host(h->frame.descriptors_left);
  const MatchResult m = mat->match(h->keyframe_kp, h->keyframe_desc, h->frame.keypoints_left, cur);
  return copy_matches(m, cap, query, train, distance);
}

// SuperPoint::infer-style host path of the adapter (the reference's demo programs): keypoints + CV_32F descriptors.
// Returns the keypoint count, -1 when infer() reports a failure.
int dropin_infer(void* handle, const uint8_t* image, int height, int width, int row_stride, int channels, int cap,
                 float* xy, float* response, float* desc) {
  Harness* h = static_cast<Harness*>(handle);
  const cv::Mat img(height, width, CV_MAKETYPE(CV_8U, channels), const_cast<uint8_t*>(image), row_stride);
  std::vector<cv::KeyPoint> kps;
  cv::Mat d;
  if (!h->sp->infer(img, kps, d)) return -1;
  const int n = static_cast<int>(kps.size());
  if (n > 0 && (d.rows != n || d.cols != 256 || d.type() != CV_32F)) return -2;
  for (int i = 0; i < n && i < cap; ++i) {
    xy[2 * i] = kps[i].pt.x, xy[2 * i + 1] = kps[i].pt.y;
    response[i] = kps[i].response;
    std::memcpy(desc + static_cast<size_t>(i) * 256, d.ptr<float>(i), 256 * sizeof(float));
  }
  return n;
}

// Keeps one more copy of the current frame alive; returns how many are held.
int dropin_hold_frame(void* handle) {
  Harness* h = static_cast<Harness*>(handle);
  h->held.push_back(h->frame);
  return static_cast<int>(h->held.size());
}

// Drops the frames the harness holds (their descriptor slots go back to the pool).
void dropin_release_frames(void* handle) {
  Harness* h = static_cast<Harness*>(handle);
  h->frame = superslam::StereoFrame();
  h->last_keyframe = superslam::StereoFrame();
  h->held.clear();
}

// ---- the other adapter classes, driven the way the reference drives the interfaces they stand behind ----------------

// IPlaceRecognizer (include/PlaceRecognizer.h:20-36) as LoopCloser uses it (src/LoopCloser.cc:29-41):
// compute_global_descriptor on the keyframe image, query against the index, add.
void* dropin_place_create(const char* weights, int in_w, int in_h, int* ok) {
  superslam_b200::EigenPlacesB200* ep = new superslam_b200::EigenPlacesB200(weights, in_w, in_h);
  *ok = ep->initialize() ? 1 : 0;
  return ep;
}
void dropin_place_destroy(void* h) { delete static_cast<superslam_b200::EigenPlacesB200*>(h); }
// returns the descriptor length (0 = empty Mat)
int dropin_place_compute(void* h, const uint8_t* image, int height, int width, int row_stride, int channels, float* out) {
  superslam::IPlaceRecognizer* pr = static_cast<superslam_b200::EigenPlacesB200*>(h);
  const cv::Mat img(height, width, CV_MAKETYPE(CV_8U, channels), const_cast<uint8_t*>(image), row_stride);
  const cv::Mat d = pr->compute_global_descriptor(img);
  if (d.empty()) return 0;
  std::memcpy(out, d.ptr<float>(), sizeof(float) * d.total());
  return static_cast<int>(d.total());
}
// descriptor handed over as a column vector of doubles when as_f64_column != 0 (add / query must reshape + convert)
void dropin_place_add(void* h, size_t id, const float* desc, int dim, int as_f64_column) {
  superslam::IPlaceRecognizer* pr = static_cast<superslam_b200::EigenPlacesB200*>(h);
  cv::Mat d(1, dim, CV_32F, const_cast<float*>(desc));
  if (as_f64_column) {
    cv::Mat c;
    d.reshape(1, dim).convertTo(c, CV_64F);
    pr->add(id, c);
  } else {
    pr->add(id, d);
  }
}
int dropin_place_query(void* h, const float* desc, int dim, size_t exclude_recent, int top_k, size_t* ids, float* scores) {
  superslam::IPlaceRecognizer* pr = static_cast<superslam_b200::EigenPlacesB200*>(h);
  const cv::Mat d(1, dim, CV_32F, const_cast<float*>(desc));
  const std::vector<superslam::LoopCandidate> c = pr->query(d, exclude_recent, top_k);
  for (size_t i = 0; i < c.size(); ++i) ids[i] = c[i].keyframe_id, scores[i] = c[i].score;
  return static_cast<int>(c.size());
}

// RemapB200 in place of cv::remap(img, out, M1, M2, cv::INTER_LINEAR) (examples/stereo/euroc.cc:176-177)
int dropin_remap(const float* map_x, const float* map_y, int dst_h, int dst_w, const uint8_t* src, int src_h, int src_w,
                 int src_stride, uint8_t* dst) {
  const cv::Mat mx(dst_h, dst_w, CV_32FC1, const_cast<float*>(map_x)), my(dst_h, dst_w, CV_32FC1, const_cast<float*>(map_y));
  superslam_b200::RemapB200 rect(mx, my, cv::Size(src_w, src_h));
  const cv::Mat in(src_h, src_w, CV_8UC1, const_cast<uint8_t*>(src), src_stride);
  cv::Mat out;
  if (!rect(in, out)) return 0;
  if (out.rows != dst_h || out.cols != dst_w || out.type() != CV_8UC1 || !out.isContinuous()) return -1;
  std::memcpy(dst, out.data, static_cast<size_t>(dst_h) * dst_w);
  return 1;
}

// RgbdPostB200::run in place of the tail of RgbdFrontEnd::process (src/RgbdFrontEnd.cc:27-58).  depth_type: 0 = CV_16U,
// 1 = CV_32F, 2 = CV_8U (an unsupported type: sampleDepth returns 0); dist as a CV_32F column when dist_f32_column != 0.
int dropin_rgbd_post(const float* xy, int n, void* depth, int depth_type, int dh, int dw, int depth_stride,
                     const double* cam4, const double* dist, int nd, int dist_f32_column, double bf, double depth_factor,
                     double max_depth, float* out_xy, double* out_stereo, char* out_has) {
  superslam_b200::RgbdPostB200 post(4096, cv::Size(dw, dh));
  std::vector<cv::Point2f> raw(n), und;
  for (int i = 0; i < n; ++i) raw[i] = cv::Point2f(xy[2 * i], xy[2 * i + 1]);
  const int types[3] = {CV_16U, CV_32F, CV_8U};
  const cv::Mat dm(dh, dw, types[depth_type], depth, depth_stride);
  cv::Mat D;
  if (nd > 0) {
    D = cv::Mat(1, nd, CV_64F, const_cast<double*>(dist));
    if (dist_f32_column) D.reshape(1, nd).convertTo(D, CV_32F);
  }
  std::vector<double> st;
  std::vector<char> has;
  if (!post.run(raw, dm, cam4[0], cam4[1], cam4[2], cam4[3], D, bf, depth_factor, max_depth, und, st, has)) return 0;
  if (static_cast<int>(und.size()) != n || static_cast<int>(st.size()) != 3 * n || static_cast<int>(has.size()) != n) return -1;
  for (int i = 0; i < n; ++i) {
    out_xy[2 * i] = und[i].x, out_xy[2 * i + 1] = und[i].y;
    for (int k = 0; k < 3; ++k) out_stereo[3 * i + k] = st[3 * i + k];
    out_has[i] = has[i];
  }
  return 1;
}
}
