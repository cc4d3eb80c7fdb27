#!/usr/bin/env python
"""Convert a Magic Leap SuperPoint checkpoint (.pth state dict) to the SSBW archive the C++ runtime
and the oracle load.  Tensor names and shapes are kept as in the checkpoint
(/root/reference/utils/convert_superpoint_to_onnx.py:38-49 defines the 12 conv layers).

    python tools/convert_superpoint_weights.py /root/reference/weights/superpoint_v1.pth \
        superslam_b200/weights/superpoint_v1.ssbw
"""
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from superslam_b200.weights_io import save_archive  # noqa: E402

ORDER = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b",
         "convPa", "convPb", "convDa", "convDb"]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    state = torch.load(src, map_location="cpu", weights_only=True)
    if isinstance(state, dict):
        state = state.get("model", state.get("state_dict", state))
    out = OrderedDict()
    for name in ORDER:
        out[name + ".weight"] = state[name + ".weight"].float().numpy()
        out[name + ".bias"] = state[name + ".bias"].float().numpy()
    save_archive(dst, out)
    n = sum(v.size for v in out.values())
    print(f"wrote {dst}: {len(out)} tensors, {n} parameters")


if __name__ == "__main__":
    main()
