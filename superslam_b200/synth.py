"""Deterministic synthetic stereo inputs: seeded random polygons, blur, noise.

The reference ships no datasets and there is no network; this generator gives images whose
SuperPoint response saturates max_keypoints (uniform noise does not).  Right image = left rolled
12 px to the left, so true matches have disparity 12 and zero row offset and pass the
StereoFrontEnd post-filter (/root/reference/src/StereoFrontEnd.cc:41-44).
"""
from __future__ import annotations

import numpy as np

_DEFAULT_SHAPES = {(480, 640): 400, (376, 1241): 1500, (480, 752): 500, (720, 1280): 3000}


def default_n_shapes(h: int, w: int) -> int:
    if (h, w) in _DEFAULT_SHAPES:
        return _DEFAULT_SHAPES[(h, w)]
    return max(8, int(400 * (h * w) / (480 * 640)))


def synth_image(h: int, w: int, seed: int, n_shapes: int | None = None) -> np.ndarray:
    """One u8 grayscale image [h, w]."""
    import cv2

    if n_shapes is None:
        n_shapes = default_n_shapes(h, w)
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128, np.uint8)
    for _ in range(n_shapes):
        nv = int(rng.integers(3, 7))
        cx = rng.uniform(0, w)
        cy = rng.uniform(0, h)
        pts = np.stack([cx + rng.uniform(-40, 40, nv), cy + rng.uniform(-40, 40, nv)], 1)
        cv2.fillPoly(img, [pts.astype(np.int32)], int(rng.integers(0, 256)))
    img = cv2.GaussianBlur(img, (3, 3), 0.8)
    noisy = img.astype(np.float32) + rng.normal(0.0, 2.0, img.shape).astype(np.float32)
    return np.clip(np.rint(noisy), 0, 255).astype(np.uint8)


def synth_pair(h: int, w: int, seed: int, n_shapes: int | None = None, disparity: int = 12):
    """(left, right) u8 images; right = left shifted `disparity` px to the left."""
    left = synth_image(h, w, seed, n_shapes)
    right = np.ascontiguousarray(np.roll(left, -disparity, axis=1))
    return left, right
