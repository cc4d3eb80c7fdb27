#!/usr/bin/env python
"""Convert a cvg/LightGlue checkpoint (e.g. superpoint_lightglue.pth, fetched by the `lightglue`
package the reference's exporter uses: /root/reference/utils/convert_lightglue_to_onnx.py:69) to the
SSBW archive the C++ runtime loads.  Accepts both key spellings (self_attn.{i}.* and
transformers.{i}.self_attn.*).

    python tools/convert_lightglue_weights.py superpoint_lightglue.pth superslam_b200/weights/lightglue.ssbw
    python tools/convert_lightglue_weights.py --synthetic 7 /tmp/lightglue_synth.ssbw
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from superslam_b200.lightglue_weights import make_random_weights, save_state_dict  # noqa: E402


def main():
    if sys.argv[1] == "--synthetic":
        sd = make_random_weights(int(sys.argv[2]))
        dst = sys.argv[3]
    else:
        sd = torch.load(sys.argv[1], map_location="cpu", weights_only=True)
        if isinstance(sd, dict) and "state_dict" in sd:
            sd = sd["state_dict"]
        dst = sys.argv[2]
    save_state_dict(sd, dst)
    print(f"wrote {dst}")


if __name__ == "__main__":
    main()
