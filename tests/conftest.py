import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SP_WEIGHTS = os.path.join(ROOT, "superslam_b200", "weights", "superpoint_v1.ssbw")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def sp_weights():
    from oracle import superpoint as osp

    return osp.load_weights(SP_WEIGHTS)


@pytest.fixture(scope="session")
def lg_weights():
    from oracle import lightglue as olg

    return olg.make_random_weights()
