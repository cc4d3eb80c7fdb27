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
                   int channels, double timestamp, int cap, float* xy, float* response, float* size_angle,
                   double* stereo, char* has_depth, int* desc) {
  Harness* h = static_cast<Harness*>(handle);
  const int type = CV_MAKETYPE(CV_8U, channels);
  const cv::Mat l(height, width, type, const_cast<uint8_t*>(left), row_stride);
  const cv::Mat r(height, width, type, const_cast<uint8_t*>(right), row_stride);
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
}
