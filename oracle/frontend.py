"""Oracle: the callers either side of the networks (host logic), numpy.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
  * stereo_postfilter  /root/reference/src/StereoFrontEnd.cc:22-47
  * dmatches           /root/reference/src/LightGlue.cc:352-361
  * FreeList           /root/reference/include/DescriptorPool.h:25-44
"""
from __future__ import annotations

import numpy as np


def dmatches(matches0: np.ndarray, mscores0: np.ndarray):
    """postprocess_outputs: increasing queryIdx, drop j<0, distance = 1 - score."""
    q = np.nonzero(matches0 >= 0)[0].astype(np.int32)
    return q, matches0[q].astype(np.int32), (np.float32(1.0) - mscores0[q].astype(np.float32))


def stereo_postfilter(xy_left: np.ndarray, xy_right: np.ndarray, query: np.ndarray, train: np.ndarray,
                      min_disparity: float = 1.0):
    """StereoFrontEnd::process after the match: stereo[i] = (uL, NaN, v) by default; for each match
    keep it iff uL-uR >= min_disparity (float compare) and |vL-vR| <= 2.0f."""
    n = xy_left.shape[0]
    stereo = np.empty((n, 3), np.float64)
    stereo[:, 0] = xy_left[:, 0]
    stereo[:, 1] = np.nan
    stereo[:, 2] = xy_left[:, 1]
    has_depth = np.zeros(n, np.int8)
    md = np.float32(min_disparity)
    for i, j in zip(query.tolist(), train.tolist()):
        if i < 0 or j < 0 or i >= n or j >= xy_right.shape[0]:
            continue
        uL, v = np.float32(xy_left[i, 0]), np.float32(xy_left[i, 1])
        uR = np.float32(xy_right[j, 0])
        if np.float32(uL - uR) < md:
            continue
        if abs(np.float32(xy_left[i, 1] - xy_right[j, 1])) > np.float32(2.0):
            continue
        stereo[i] = (uL, uR, v)
        has_depth[i] = 1
    return stereo, has_depth


class FreeList:
    """LIFO free list over n slots; acquire() == -1 when exhausted (DescriptorPool.h:25-44)."""

    def __init__(self, n: int):
        self.n = n
        self.free = list(range(n - 1, -1, -1))

    def acquire(self) -> int:
        return self.free.pop() if self.free else -1

    def release(self, slot: int) -> None:
        self.free.append(slot)

    def in_use(self) -> int:
        return self.n - len(self.free)
