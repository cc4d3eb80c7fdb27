// C-ABI shim that RUNS the reference's own StereoFrontEnd::process (/root/reference/src/StereoFrontEnd.cc:10-49)
// and StereoFrame::backproject (/root/reference/src/StereoFrame.cc:5-13), compiled from the sources where they lie,
// behind mock IFeatureExtractor / IFeatureMatcher objects that hand back the caller's keypoints and matches - the
// same construction as the reference's tests/test_stereo_frontend.cc.  OpenCV / GTSAM value types are the small
// functional stand-ins of oracle/stubs/ (the code under test only reads keypoint floats and writes three doubles).
// Built by oracle/Makefile into oracle/_ref/libref_frontend.so.  TEST INFRASTRUCTURE.
#include <utility>
#include <vector>

#include "StereoFrontEnd.h"

namespace {
using superslam::DeviceDescriptors;
using superslam::Features;

struct MockExtractor : superslam::IFeatureExtractor {
  Features l, r;
  Features extract(const cv::Mat&) override { return l; }
  std::pair<Features, Features> extract_stereo(const cv::Mat&, const cv::Mat&) override { return {l, r}; }
};
struct MockMatcher : superslam::IFeatureMatcher {
  MatchResult m;
  MatchResult match(const std::vector<cv::KeyPoint>&, const cv::Mat&, const std::vector<cv::KeyPoint>&,
                    const cv::Mat&) override { return m; }
  MatchResult match(const std::vector<cv::KeyPoint>&, const DeviceDescriptors&, const std::vector<cv::KeyPoint>&,
                    const DeviceDescriptors&) override { return m; }
  cv::Mat descriptors_to_host(const DeviceDescriptors&) override { return cv::Mat(); }
};
}  // namespace

extern "C" {
// xy_l [nl][2], xy_r [nr][2] float keypoints; (query, train) [nm] match indices; out_stereo [nl][3], out_has [nl]
void ref_stereo_frontend_process(const float* xy_l, int nl, const float* xy_r, int nr, const int* query, const int* train,
                                 int nm, float min_disparity, double* out_stereo, char* out_has) {
  MockExtractor ext;
  MockMatcher mat;
  for (int i = 0; i < nl; ++i) ext.l.keypoints.emplace_back(xy_l[2 * i], xy_l[2 * i + 1], 1.0f);
  for (int i = 0; i < nr; ++i) ext.r.keypoints.emplace_back(xy_r[2 * i], xy_r[2 * i + 1], 1.0f);
  for (int i = 0; i < nm; ++i) mat.m.matches.emplace_back(query[i], train[i], 0.0f);
  superslam::StereoFrontEnd fe(&ext, &mat, gtsam::Cal3_S2Stereo(500, 500, 0, 320, 240, 0.5), min_disparity);
  const superslam::StereoFrame f = fe.process(cv::Mat(), cv::Mat(), 0.0);
  for (int i = 0; i < nl; ++i) {
    out_stereo[3 * i] = f.stereo[i].uL();
    out_stereo[3 * i + 1] = f.stereo[i].uR();
    out_stereo[3 * i + 2] = f.stereo[i].v();
    out_has[i] = f.has_depth[i];
  }
}
// StereoFrame::backproject for one stereo observation under pose (R row-major, t) and calibration (fx, fy, px, py, b)
void ref_stereo_frame_backproject(const double* stereo3, const double* R9, const double* t3, const double* cal5,
                                  double* out3) {
  superslam::StereoFrame f;
  f.stereo = {gtsam::StereoPoint2(stereo3[0], stereo3[1], stereo3[2])};
  f.has_depth = {1};
  f.pose = gtsam::Pose3(R9, t3);
  const gtsam::Point3 p = f.backproject(0, gtsam::Cal3_S2Stereo(cal5[0], cal5[1], 0, cal5[2], cal5[3], cal5[4]));
  out3[0] = p.x(), out3[1] = p.y(), out3[2] = p.z();
}
}
