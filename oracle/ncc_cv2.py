"""TEST INFRASTRUCTURE ONLY — the executed-OpenCV NCC oracle.

A line-for-line Python mirror of ``serial::disparityNCorr``
(/root/reference/ProblemSets/ps2_cpp/lib/DisparityNCorr.cpp:27-68) around the real
``cv2.matchTemplate(..., TM_CCORR_NORMED)`` / ``cv2.minMaxLoc`` (OpenCV 4.13 here; the reference
pins 3.4.1 — the TM_CCORR_NORMED code path is the same).  Slow (one matchTemplate call per pixel);
used to pin oracle/stereo_oracle.c's restatement and to generate tests/golden/ fixtures.
"""
from __future__ import annotations

import numpy as np


def ncorr_cv2(left: np.ndarray, right: np.ndarray, window_rad: int, min_disp: int, max_disp: int):
    import cv2

    cv2.setNumThreads(1)
    left = np.ascontiguousarray(left, np.float32)
    right = np.ascontiguousarray(right, np.float32)
    R = int(window_rad)
    lp = cv2.copyMakeBorder(left, R, R, R, R, cv2.BORDER_REPLICATE)      # :28-31
    rp = cv2.copyMakeBorder(right, R, R, R, R, cv2.BORDER_REPLICATE)
    rows, cols = left.shape
    w = 2 * R + 1
    disp = np.zeros((rows, cols), np.int32)
    score = np.zeros((rows, cols), np.float32)
    for y in range(R, lp.shape[0] - R):                                     # :44
        for x in range(R, lp.shape[1] - R):                                 # :45
            templ = lp[y - R:y - R + w, x - R:x - R + w].copy()             # :47-48
            start_x = max(0, x + min_disp - R)                              # :50
            end_x = min(lp.shape[1], x + max_disp + 1 + R)                  # :51
            search = rp[y - R:y - R + w, start_x:end_x].copy()              # :52-53
            result = cv2.matchTemplate(search, templ, cv2.TM_CCORR_NORMED)  # :60
            _, max_val, _, max_loc = cv2.minMaxLoc(result)                  # :62-64
            d = max_loc[0] - (result.shape[1] - 1 if (min_disp <= 0 and max_disp <= 0) else 0)   # :67
            disp[y - R, x - R] = d
            score[y - R, x - R] = result[0, max_loc[0]]
    return disp, score
