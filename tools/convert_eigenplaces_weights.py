#!/usr/bin/env python
"""Convert an EigenPlaces (ResNet18, 512-d) state dict to the SSBW archive `ssb_ep_create` loads.  The reference gets the
trained model from torch.hub at export time (/root/reference/utils/convert_eigenplaces_to_onnx.py:54-60:
torch.hub.load("gmberton/eigenplaces", "get_trained_model", backbone="ResNet18", fc_output_dim=512)); save its
state dict once where there is network access -

    torch.save(model.state_dict(), "eigenplaces_resnet18_512.pth")

- and convert it here (tensor names are kept: backbone.{0,1,4..7}.*, aggregation.1.p, aggregation.3.{weight,bias};
BatchNorm `num_batches_tracked` counters are dropped, BN is folded into the convolutions at load time by the library):

    python tools/convert_eigenplaces_weights.py eigenplaces_resnet18_512.pth eigenplaces_resnet18_512.ssbw
    python tools/convert_eigenplaces_weights.py --synthetic 11 /tmp/eigenplaces_synth.ssbw
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from superslam_b200.eigenplaces_weights import make_random_weights, save_state_dict  # noqa: E402


def main():
    if sys.argv[1] == "--synthetic":
        sd, dst = make_random_weights(int(sys.argv[2])), sys.argv[3]
    else:
        sd = torch.load(sys.argv[1], map_location="cpu", weights_only=True)
        if isinstance(sd, dict) and "state_dict" in sd:
            sd = sd["state_dict"]
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
        dst = sys.argv[2]
    save_state_dict(sd, dst)
    print(f"wrote {dst}")


if __name__ == "__main__":
    main()
