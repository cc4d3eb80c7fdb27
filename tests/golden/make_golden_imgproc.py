"""Golden vectors for the image front door, produced by OpenCV itself (cv2 of the authoring image
runs the C++ the reference links: cv::initUndistortRectifyMap, cv::remap, cv::undistortPoints).

    python tests/golden/make_golden_imgproc.py        # writes tests/golden/imgproc_cv2.npz (~60 KB)

The maps are stored too, so the consumers need no OpenCV."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from superslam_b200.synth import synth_pair  # noqa: E402

K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]]) / 4.7      # EuRoC cam0, scaled to 160x102
K[2, 2] = 1
D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0])
P = np.array([[435.2 / 4.7, 0, 367.45 / 4.7], [0, 435.2 / 4.7, 252.2 / 4.7], [0, 0, 1]])
D8 = np.array([-0.2, 0.05, 0.001, -0.0005, 0.01, 0.02, -0.01, 0.003])


def main():
    h, w = 102, 160
    img, _ = synth_pair(h, w, 60, 11)
    R = cv2.Rodrigues(np.array([0.0077, -0.0049, 0.0016]))[0]
    m1, m2 = cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32F)
    # a second, wilder map: large distortion pushes border pixels outside the source (constant border 0)
    m1b, m2b = cv2.initUndistortRectifyMap(K, D * 3.0, R, P * np.array([[0.8, 1, 1], [1, 0.8, 1], [1, 1, 1]]), (w, h),
                                           cv2.CV_32F)
    rng = np.random.default_rng(7)
    pts = rng.uniform([0, 0], [w, h], (300, 2)).astype(np.float32)
    # a colour image whose channels differ (B = img, G = inverted, R = shifted) and OpenCV's own gray conversion
    bgr = np.stack([img, 255 - img, np.roll(img, 7, axis=1)], -1)
    out = {
        "bgr": bgr, "gray": cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY),
        "image": img, "map_x": m1, "map_y": m2, "remap": cv2.remap(img, m1, m2, cv2.INTER_LINEAR),
        "map_x_wide": m1b, "map_y_wide": m2b, "remap_wide": cv2.remap(img, m1b, m2b, cv2.INTER_LINEAR),
        "camera": np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]]), "dist5": D, "dist8": D8, "points": pts,
        "undist5": cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2),
        "undist8": cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D8, None, K).reshape(-1, 2),
        "cv2_version": np.array(cv2.__version__),
    }
    path = os.path.join(ROOT, "tests", "golden", "imgproc_cv2.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; border zeros in the wide remap:", int((out["remap_wide"] == 0).sum()))


if __name__ == "__main__":
    main()
