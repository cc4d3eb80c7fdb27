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

// ---- RgbdFrontEnd::process (/root/reference/src/RgbdFrontEnd.cc:23-58), the real code; cv::undistortPoints is
// forwarded to `undistort` (the test passes cv2.undistortPoints).  depth_type 0 = CV_16U, 1 = CV_32F.
#include <opencv4/opencv2/calib3d.hpp>

#include "RgbdFrontEnd.h"

extern "C" void ref_rgbd_frontend_process(const float* xy, int n, void* depth, int depth_type, int dh, int dw,
                                          const double* cam4, double baseline, double* dist, int nd,
                                          double depth_factor, double max_depth, cv::UndistortPointsFn undistort,
                                          float* out_xy, double* out_stereo, char* out_has) {
  MockExtractor ext;
  for (int i = 0; i < n; ++i) ext.l.keypoints.emplace_back(xy[2 * i], xy[2 * i + 1], 1.0f);
  double k9[9] = {cam4[0], 0, cam4[2], 0, cam4[1], cam4[3], 0, 0, 1};
  const cv::Mat K(3, 3, CV_64F, k9, 3 * sizeof(double));
  const cv::Mat D = nd > 0 ? cv::Mat(1, nd, CV_64F, dist, nd * sizeof(double)) : cv::Mat();
  const size_t esz = depth_type == 0 ? 2 : 4;
  const cv::Mat depth_m(dh, dw, depth_type == 0 ? CV_16U : CV_32F, depth, dw * esz);
  cv::undistort_points_hook() = undistort;
  superslam::RgbdFrontEnd fe(&ext, gtsam::Cal3_S2Stereo(cam4[0], cam4[1], 0, cam4[2], cam4[3], baseline), depth_factor,
                             max_depth, K, D);
  const superslam::StereoFrame f = fe.process(cv::Mat(), depth_m, 0.0);
  for (int i = 0; i < n; ++i) {
    out_xy[2 * i] = f.keypoints_left[i].pt.x;
    out_xy[2 * i + 1] = f.keypoints_left[i].pt.y;
    out_stereo[3 * i] = f.stereo[i].uL();
    out_stereo[3 * i + 1] = f.stereo[i].uR();
    out_stereo[3 * i + 2] = f.stereo[i].v();
    out_has[i] = f.has_depth[i];
  }
}
