"""Host-side mirror of the reference's ps2 disparity interface on top of the C ABI.

Reference surface reproduced (names, argument meaning, error behaviour):

* ``cuda::disparitySSD(left, right, windowRad, minDisparity, maxDisparity, disparity)``
  — ProblemSets/ps2_cpp/include/DisparitySSD.h:18-23
* ``cuda::disparityNCorr(...)`` — ProblemSets/ps2_cpp/include/DisparityNCorr.h:19-24
* ``disparitySSDPair`` / ``disparityNCorrPair`` — ProblemSets/ps2_cpp/src/main.cpp:21-48, 51-78
* ``Config::DisparitySSD{_windowRadius, _disparityRange}`` — ProblemSets/ps2_cpp/include/Config.h:40-46

Inputs are numpy arrays (the stand-in for ``cv::Mat``): float32 ``CV_32FC1`` like the reference
asserts (DisparitySSD.cu:150), or uint8.  The returned disparity is int8 (``CV_8SC1``) by default,
exactly like the reference; pass ``dtype=np.int16``/``np.int32`` for searches beyond 127.
All computation happens in libstereo_b200.so on a B200; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _capi
from ._capi import COST_NCORR, COST_SSD


class StereoError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = _capi.last_error()
        name = _capi.lib().stereo_status_string(status).decode()
        super().__init__(f"{where}: {name} ({status}): {msg}")


def _check(status: int, where: str) -> None:
    if status != _capi.STEREO_OK:
        raise StereoError(status, where)


@dataclass
class DisparityConfig:
    """``Config::DisparitySSD`` (include/Config.h:40-46): YAML keys ``window_radius``, ``disparity_range``."""
    window_radius: int
    disparity_range: int


_ELEM = {np.dtype(np.int8): 1, np.dtype(np.int16): 2, np.dtype(np.int32): 4}


class Context:
    """Owns the device-side state for one GPU (``stereo_ctx``)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(_capi.lib().stereo_ctx_create(int(device), C.byref(self._h)), "stereo_ctx_create")
        self.device = int(device)

    def close(self) -> None:
        if self._h:
            _capi.lib().stereo_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- introspection -------------------------------------------------------------------------
    @property
    def handle(self) -> C.c_void_p:
        return self._h

    @property
    def last_path(self) -> int:
        return int(_capi.lib().stereo_ctx_last_path(self._h))

    @property
    def last_kernel_ms(self) -> float:
        return float(_capi.lib().stereo_ctx_last_kernel_ms(self._h))

    @property
    def last_launches(self) -> int:
        return int(_capi.lib().stereo_ctx_last_launches(self._h))

    def last_hot_kernel_ms(self):
        """(milliseconds, launches measured) of the last call's hot kernels; (-1, 0) if none ran."""
        n = C.c_int(0)
        ms = float(_capi.lib().stereo_ctx_last_hot_kernel_ms(self._h, C.byref(n)))
        return ms, int(n.value)

    @property
    def last_hot_jobs(self) -> int:
        """Directions (one disparity map each) the measured hot launches of the last call covered."""
        return int(_capi.lib().stereo_ctx_last_hot_jobs(self._h))

    def force_path(self, path: int) -> None:
        _check(_capi.lib().stereo_ctx_force_path(self._h, int(path)), "stereo_ctx_force_path")

    def set_pipe_bands(self, bands: int) -> None:
        """Row bands per image pair in the pipelined host entry points (0 = automatic)."""
        _check(_capi.lib().stereo_ctx_set_pipe_bands(self._h, int(bands)), "stereo_ctx_set_pipe_bands")

    @property
    def last_fused_pairs(self) -> int:
        """Image pairs of the last call whose two maps came out of one cost volume."""
        return int(_capi.lib().stereo_ctx_last_fused_pairs(self._h))

    def set_fuse_pairs(self, on: bool) -> None:
        """Pair calls: both maps from one cost volume where possible (default on; results identical either way)."""
        _check(_capi.lib().stereo_ctx_set_fuse_pairs(self._h, int(bool(on))), "stereo_ctx_set_fuse_pairs")

    def set_host_threads(self, threads: int) -> None:
        """Host threads that convert CV_32FC1 host images to u8 before the upload (0 = automatic, -1 = off)."""
        _check(_capi.lib().stereo_ctx_set_host_threads(self._h, int(threads)), "stereo_ctx_set_host_threads")

    @property
    def host_threads(self) -> int:
        return int(_capi.lib().stereo_ctx_host_threads(self._h))

    def synchronize(self, stream: int = 0) -> None:
        _check(_capi.lib().stereo_ctx_synchronize(self._h, C.c_void_p(stream)), "stereo_ctx_synchronize")

    # -- host-buffer entry points ----------------------------------------------------------------
    def disparity(self, cost: int, ref: np.ndarray, tgt: np.ndarray, window_rad: int, min_disp: int, max_disp: int,
                  dtype=np.int8, return_best: bool = False):
        ref, tgt, fn = _prep_pair(ref, tgt, "stereo_disparity_f32_host", "stereo_disparity_u8_host")
        rows, cols = ref.shape
        dt = np.dtype(dtype)
        disp = np.empty((rows, cols), dt)
        best = np.empty((rows, cols), np.int32 if cost == COST_SSD else np.float32) if return_best else None
        st = fn(self._h, int(cost), ref.ctypes.data, ref.strides[0], tgt.ctypes.data, tgt.strides[0], rows, cols,
                int(window_rad), int(min_disp), int(max_disp), disp.ctypes.data, disp.strides[0], _ELEM[dt],
                best.ctypes.data if return_best else None, best.strides[0] if return_best else 0)
        _check(st, fn.__name__)
        return (disp, best) if return_best else disp

    def disparity_refgpu(self, cost: int, ref: np.ndarray, tgt: np.ndarray, window_rad: int, min_disp: int, max_disp: int,
                         return_best: bool = False):
        """The function the reference's GPU kernels compute (``stereo_disparity_refgpu_f32_host``): int8 map,
        -1 where no candidate passed the kernels' threshold."""
        ref, tgt = np.ascontiguousarray(ref, np.float32), np.ascontiguousarray(tgt, np.float32)
        if ref.ndim != 2 or ref.shape != tgt.shape:
            raise ValueError("left/right must be 2-D arrays of equal shape")
        rows, cols = ref.shape
        disp = np.empty((rows, cols), np.int8)
        best = np.empty((rows, cols), np.float32) if return_best else None
        st = _capi.lib().stereo_disparity_refgpu_f32_host(
            self._h, int(cost), ref.ctypes.data, ref.strides[0], tgt.ctypes.data, tgt.strides[0], rows, cols, int(window_rad),
            int(min_disp), int(max_disp), disp.ctypes.data, disp.strides[0], best.ctypes.data if return_best else None,
            best.strides[0] if return_best else 0)
        _check(st, "stereo_disparity_refgpu_f32_host")
        return (disp, best) if return_best else disp

    def disparity_pair(self, cost: int, left: np.ndarray, right: np.ndarray, window_rad: int, disparity_range: int,
                       dtype=np.int8) -> Tuple[np.ndarray, np.ndarray]:
        left, right, fn = _prep_pair(left, right, "stereo_disparity_pair_f32_host", "stereo_disparity_pair_u8_host")
        rows, cols = left.shape
        dt = np.dtype(dtype)
        dl = np.empty((rows, cols), dt)
        dr = np.empty((rows, cols), dt)
        st = fn(self._h, int(cost), left.ctypes.data, left.strides[0], right.ctypes.data, right.strides[0], rows, cols,
                int(window_rad), int(disparity_range), dl.ctypes.data, dr.ctypes.data, dl.strides[0], _ELEM[dt])
        _check(st, fn.__name__)
        return dl, dr

    def disparity_pair_batch(self, cost: int, left: np.ndarray, right: np.ndarray, window_rad: int,
                             disparity_range: int, dtype=np.int8) -> Tuple[np.ndarray, np.ndarray]:
        """left/right: arrays of shape (n_pairs, rows, cols), both uint8 or both float32 (CV_32FC1 images, the
        type the reference's entry points take): a batch of ``disparitySSDPair`` / ``disparityNCorrPair`` calls."""
        left, right = np.asarray(left), np.asarray(right)
        if left.dtype == np.float32 and right.dtype == np.float32:
            kind, name = np.float32, "stereo_disparity_pair_batch_f32_host"
        elif left.dtype == np.uint8 and right.dtype == np.uint8:
            kind, name = np.uint8, "stereo_disparity_pair_batch_u8_host"
        else:
            raise TypeError("batch images must both be float32 (CV_32FC1) or both uint8")
        left = np.ascontiguousarray(left, kind)
        right = np.ascontiguousarray(right, kind)
        if left.ndim != 3 or left.shape != right.shape:
            raise ValueError("batch inputs must be (n, rows, cols) arrays of equal shape")
        n, rows, cols = left.shape
        dt = np.dtype(dtype)
        dl = np.empty((n, rows, cols), dt)
        dr = np.empty((n, rows, cols), dt)
        st = getattr(_capi.lib(), name)(
            self._h, int(cost), n, left.ctypes.data, right.ctypes.data, left.strides[1], left.strides[0], rows, cols,
            int(window_rad), int(disparity_range), dl.ctypes.data, dr.ctypes.data, dl.strides[1], dl.strides[0], _ELEM[dt])
        _check(st, name)
        return dl, dr


def _prep_pair(a: np.ndarray, b: np.ndarray, f32_name: str, u8_name: str):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.ndim != 2 or a.shape != b.shape:
        # the reference asserts equal sizes in its callers (main.cpp:28)
        raise ValueError("left/right must be 2-D arrays of equal shape")
    if a.dtype == np.uint8 and b.dtype == np.uint8:
        kind, name = np.uint8, u8_name
    elif a.dtype == np.float32 and b.dtype == np.float32:
        kind, name = np.float32, f32_name
    else:
        # the reference: assert(left.type() == CV_32FC1 && right.type() == CV_32FC1) (DisparitySSD.cu:150)
        raise TypeError("images must both be float32 (CV_32FC1) or both uint8")
    # cv::Mat layout only: unit column stride and a non-negative row step of at least one row (views such as
    # img[::-1] or img[:, ::2] are copied; a negative step would wrap in the C ABI's size_t)
    def _mat(v):
        ok = v.strides[1] == v.itemsize and v.strides[0] >= v.shape[1] * v.itemsize
        return v if ok else np.ascontiguousarray(v, kind)
    a, b = _mat(a), _mat(b)
    return a, b, getattr(_capi.lib(), name)


class MultiGpu:
    """Several B200s in one process (``stereo_mgpu_*``): batches sharded by pair, one pair by row bands; host arrays in,
    host arrays out.  ``devices=None``: every visible sm_100 device; an ordinal may repeat (several contexts on one GPU)."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        if devices is None:
            st = _capi.lib().stereo_mgpu_create(None, 0, C.byref(self._h))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            st = _capi.lib().stereo_mgpu_create(arr, len(devices), C.byref(self._h))
        _check(st, "stereo_mgpu_create")

    @property
    def device_count(self) -> int:
        return int(_capi.lib().stereo_mgpu_device_count(self._h))

    def close(self) -> None:
        if self._h:
            _capi.lib().stereo_mgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def disparity_pair_bands(self, cost: int, left: np.ndarray, right: np.ndarray, window_rad: int, disparity_range: int,
                             dtype=np.int8) -> Tuple[np.ndarray, np.ndarray]:
        left, right, fn = _prep_pair(left, right, "stereo_mgpu_disparity_pair_bands_f32_host", "stereo_mgpu_disparity_pair_bands_u8_host")
        rows, cols = left.shape
        dt = np.dtype(dtype)
        dl, dr = np.empty((rows, cols), dt), np.empty((rows, cols), dt)
        st = fn(self._h, int(cost), left.ctypes.data, left.strides[0], right.ctypes.data, right.strides[0], rows, cols,
                int(window_rad), int(disparity_range), dl.ctypes.data, dr.ctypes.data, dl.strides[0], _ELEM[dt])
        _check(st, fn.__name__)
        return dl, dr

    def disparity_pair_batch(self, cost: int, left: np.ndarray, right: np.ndarray, window_rad: int, disparity_range: int,
                             dtype=np.int8) -> Tuple[np.ndarray, np.ndarray]:
        left, right = np.asarray(left), np.asarray(right)
        if left.dtype == np.float32 and right.dtype == np.float32:
            kind, name = np.float32, "stereo_mgpu_disparity_pair_batch_f32_host"
        elif left.dtype == np.uint8 and right.dtype == np.uint8:
            kind, name = np.uint8, "stereo_mgpu_disparity_pair_batch_u8_host"
        else:
            raise TypeError("batch images must both be float32 (CV_32FC1) or both uint8")
        left, right = np.ascontiguousarray(left, kind), np.ascontiguousarray(right, kind)
        if left.ndim != 3 or left.shape != right.shape:
            raise ValueError("batch inputs must be (n, rows, cols) arrays of equal shape")
        n, rows, cols = left.shape
        dt = np.dtype(dtype)
        dl, dr = np.empty((n, rows, cols), dt), np.empty((n, rows, cols), dt)
        st = getattr(_capi.lib(), name)(
            self._h, int(cost), n, left.ctypes.data, right.ctypes.data, left.strides[1], left.strides[0], rows, cols,
            int(window_rad), int(disparity_range), dl.ctypes.data, dr.ctypes.data, dl.strides[1], dl.strides[0], _ELEM[dt])
        _check(st, name)
        return dl, dr


_default: Optional[Context] = None


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context(0)
    return _default


# ---- the reference's four free functions ------------------------------------------------------------

def disparitySSD(left, right, windowRad: int, minDisparity: int, maxDisparity: int, dtype=np.int8,
                 ctx: Optional[Context] = None) -> np.ndarray:
    """``cuda::disparitySSD`` (DisparitySSD.h:18-23).  ``left`` is the reference image."""
    return (ctx or default_context()).disparity(COST_SSD, left, right, windowRad, minDisparity, maxDisparity, dtype)


def disparityNCorr(left, right, windowRad: int, minDisparity: int, maxDisparity: int, dtype=np.int8,
                   ctx: Optional[Context] = None) -> np.ndarray:
    """``cuda::disparityNCorr`` (DisparityNCorr.h:19-24)."""
    return (ctx or default_context()).disparity(COST_NCORR, left, right, windowRad, minDisparity, maxDisparity, dtype)


def disparitySSDPair(left, right, config: DisparityConfig, dtype=np.int8, ctx: Optional[Context] = None):
    """``disparitySSDPair`` (main.cpp:21-48): (left-referenced map, right-referenced map)."""
    return (ctx or default_context()).disparity_pair(COST_SSD, left, right, config.window_radius,
                                                     config.disparity_range, dtype)


def disparityNCorrPair(left, right, config: DisparityConfig, dtype=np.int8, ctx: Optional[Context] = None):
    """``disparityNCorrPair`` (main.cpp:51-78)."""
    return (ctx or default_context()).disparity_pair(COST_NCORR, left, right, config.window_radius,
                                                     config.disparity_range, dtype)
