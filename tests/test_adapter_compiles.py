"""The C++ adapter (include/superslam_b200_adapter.hpp) implements the reference's own abstract interfaces, so it
can only be compiled against the reference's headers.  This image has no OpenCV C++ headers; a declaration-only
stub of the few cv:: types involved (tests/stubs/) lets g++ type-check the adapter against the REAL
InferenceInterfaces.h / PlaceRecognizer.h / DescriptorPool.h of /root/reference: every override must match the
interface's signature, every C-ABI call the header's prototype."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

REF_INC = "/root/reference/include"


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference headers not available on this machine")
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_adapter_type_checks_against_reference_interfaces(tmp_path):
    src = tmp_path / "tu.cpp"
    src.write_text(
        '#include "superslam_b200_adapter.hpp"\n'
        "// instantiate every adapter so that missing overrides (abstract classes) are errors\n"
        "void instantiate(const cv::Mat& m) {\n"
        '  superslam_b200::SuperPointB200 sp("w.ssbw", 1024, 0.005, 4);\n'
        '  superslam_b200::LightGlueB200 lg("w.ssbw", 640, 480, 1024);\n'
        "  superslam_b200::LightGlueB200 lg2(lg, 640, 480);\n"
        '  superslam_b200::EigenPlacesB200 ep("w.ssbw", 512, 512);\n'
        "  superslam::IFeatureExtractor* e = &sp;\n"
        "  superslam::IFeatureMatcher* mt = &lg;\n"
        "  superslam::IPlaceRecognizer* pr = &ep;\n"
        "  auto f = e->extract_stereo(m, m);\n"
        "  MatchResult r = mt->match(f.first.keypoints, f.first.descriptors, f.second.keypoints, f.second.descriptors);\n"
        "  cv::Mat h = mt->descriptors_to_host(f.first.descriptors);\n"
        "  pr->add(1, pr->compute_global_descriptor(m));\n"
        "  (void)pr->query(h, 10, 3); (void)r; (void)lg2;\n"
        "  superslam_b200::RemapB200 rect(m, m, cv::Size(752, 480));\n"
        "  cv::Mat out; rect(m, out);\n"
        "  superslam_b200::RgbdPostB200 post(1024, cv::Size(640, 480));\n"
        "  std::vector<cv::Point2f> raw, und; std::vector<double> st; std::vector<char> has;\n"
        "  post.run(raw, m, 500, 500, 320, 240, m, 40.0, 5000.0, 8.0, und, st, has);\n"
        "}\n")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror=overloaded-virtual",
           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", REF_INC, "-I", os.path.join(ROOT, "include"), str(src)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
