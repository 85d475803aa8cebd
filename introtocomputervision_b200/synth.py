"""Deterministic synthetic stereo pairs: textured left image, piecewise-constant
shift field, small integer noise on the right image.  numpy only; used by tests and bench.py."""
from __future__ import annotations

import numpy as np


def make_pair(rows: int, cols: int, n_disp: int, seed: int):
    """Returns (left_u8, right_u8, true_shift) with right[y, x] = left[y, min(W-1, x + s(y, x))] + noise,
    so the left-referenced (L->R) disparity is -s... in the reference's convention the left image
    searched in the right image over [-range, 0] finds right[x + d] = left[x], i.e. d = -s where the
    scene was shifted by s.  The shift field is piecewise constant on 64x64 blocks."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (rows, cols + 2), dtype=np.int32)
    # 3-tap horizontal + vertical box low-pass keeps texture but avoids pure white noise
    sm = (base[:, :-2] + base[:, 1:-1] + base[:, 2:]) // 3
    up = np.vstack([sm[:1], sm[:-1]])
    dn = np.vstack([sm[1:], sm[-1:]])
    left = ((sm + up + dn) // 3).astype(np.uint8)
    yy, xx = np.mgrid[0:rows, 0:cols]
    shift = (((xx // 64) + (yy // 64)) * 7) % max(1, n_disp)
    src = np.minimum(cols - 1, xx + shift)
    right = left[yy, src].astype(np.int32) + rng.integers(-2, 3, (rows, cols))
    right = np.clip(right, 0, 255).astype(np.uint8)
    return left, right, shift.astype(np.int32)


def noisy_variant(img_u8: np.ndarray, seed: int, sigma: float = 10.0) -> np.ndarray:
    """float32 image + N(0, sigma) noise, unclipped (the reference's addNoise, main.cpp:140-153)."""
    rng = np.random.default_rng(seed)
    return (img_u8.astype(np.float32) + rng.normal(0.0, sigma, img_u8.shape).astype(np.float32)).astype(np.float32)


def contrast_variant(img_u8: np.ndarray, gain: float = 1.1) -> np.ndarray:
    """`left * 1.1f` in float32 (main.cpp:191-193)."""
    return (img_u8.astype(np.float32) * np.float32(gain)).astype(np.float32)
