#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the REFERENCE's own torch module
(/root/reference/utils/convert_superpoint_to_onnx.py: SuperPoint + DenseSuperPoint(nms_radius=4) with
/root/reference/weights/superpoint_v1.pth) in the authoring container.  /root/reference does not exist
on the GPU box, so the outputs are committed here and this script is only re-run by hand:

    python tests/golden/make_golden.py

Fixtures
  superpoint_ref_small.npz   2x(120x160) synthetic images: dense scores (after NMS) and descriptor grid,
                             straight from the reference module (fp32).
  superpoint_ref_odd.npz     1x(99x131) (sizes not divisible by 8): scores + grid, for the floor-pool /
                             narrower-score-map behaviour (SURVEY.md §0 "Keypoint coordinates").
  superpoint_ref_c2.npz      the C2 workload pair (640x480, seed 1234): per image the (h, w, score) of
                             every pixel with score > 0.005 inside the 4-px border, computed from the
                             reference module's score map, and the fp16 descriptor-grid rows at those
                             pixels' cells (so select+gather can be checked without the 2.5 MB grid).
  superpoint_ref_kitti.npz   1241x376 seed 1234: score-map shape and candidates (h, w, score) only.
The LightGlue reference cannot be run offline (un-vendored package, no weights) - no fixture.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/utils")

import convert_superpoint_to_onnx as ref  # noqa: E402  (the reference's model definition)

from superslam_b200.synth import synth_image, synth_pair  # noqa: E402


def ref_model():
    sp = ref.SuperPoint()
    sp.load_state_dict(torch.load("/root/reference/weights/superpoint_v1.pth", map_location="cpu", weights_only=True))
    return ref.DenseSuperPoint(sp.eval(), 4).eval()


def run(model, imgs_u8):
    x = torch.from_numpy(imgs_u8).float() * np.float32(1.0 / 255.0)  # cv::Mat::convertTo(CV_32F, 1/255)
    with torch.no_grad():
        s, d = model(x[:, None])
    return s.numpy(), d.numpy()


def candidates(scores, rb=4, thr=0.005):
    h, w = np.nonzero(scores.astype(np.float64) > thr)
    k = (h >= rb) & (h < scores.shape[0] - rb) & (w >= rb) & (w < scores.shape[1] - rb)
    h, w = h[k], w[k]
    return np.stack([h, w], 1).astype(np.int32), scores[h, w].astype(np.float32)


def main():
    torch.set_num_threads(8)
    m = ref_model()
    imgs = np.stack([synth_image(120, 160, 100, 40), synth_image(120, 160, 101, 40)])
    s, d = run(m, imgs)
    np.savez_compressed(os.path.join(HERE, "superpoint_ref_small.npz"), images=imgs, scores=s, grid=d)

    img = synth_image(99, 131, 102, 30)[None]
    s, d = run(m, img)
    np.savez_compressed(os.path.join(HERE, "superpoint_ref_odd.npz"), images=img, scores=s, grid=d)

    l, r = synth_pair(480, 640, 1234)
    s, d = run(m, np.stack([l, r]))
    out = {}
    for i in range(2):
        hw, sc = candidates(s[i])
        cells = np.stack([np.minimum(hw[:, 0] // 8, d.shape[2] - 1), np.minimum(hw[:, 1] // 8, d.shape[3] - 1)], 1)
        out[f"hw{i}"] = hw
        out[f"score{i}"] = sc
        out[f"rows{i}"] = d[i][:, cells[:, 0], cells[:, 1]].T.astype(np.float16)
    out["score_shape"] = np.array(s.shape[1:], np.int32)
    out["grid_shape"] = np.array(d.shape[1:], np.int32)
    np.savez_compressed(os.path.join(HERE, "superpoint_ref_c2.npz"), **out)

    img = synth_image(376, 1241, 1234)[None]
    s, d = run(m, img)
    hw, sc = candidates(s[0])
    np.savez_compressed(os.path.join(HERE, "superpoint_ref_kitti.npz"), hw=hw, score=sc,
                        score_shape=np.array(s.shape[1:], np.int32), grid_shape=np.array(d.shape[1:], np.int32))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
