"""ctypes bindings for the oracle libraries (TEST INFRASTRUCTURE ONLY).

* ``libstereo_oracle.so``  — C restatement (``stereo_oracle.c``), always buildable with gcc.
* ``_ref/libref_ssd.so``   — the reference's own ``serial::disparitySSD``
  (``/root/reference/ProblemSets/ps2_cpp/lib/DisparitySSD.cpp``) compiled in place by
  ``oracle/Makefile``; exists wherever it was built (it travels to the GPU box as a binary).
* ``_ref/libref_ncc.so``   — the reference's own ``serial::disparityNCorr``
  (``lib/DisparityNCorr.cpp``) compiled in place likewise; its ``cv::matchTemplate`` / ``cv::minMaxLoc``
  are the shim's restatement of OpenCV's TM_CCORR_NORMED arithmetic (``cvshim/opencv2/imgproc/imgproc.hpp``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "libstereo_oracle.so"
_REF = _HERE / "_ref" / "libref_ssd.so"
_REF_O0 = _HERE / "_ref" / "libref_ssd_O0.so"
_REF_NCC = _HERE / "_ref" / "libref_ncc.so"


class OracleError(RuntimeError):
    pass


def build(force: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is present, oracle/_ref."""
    need = force or not _LIB.exists() or _LIB.stat().st_mtime < (_HERE / "stereo_oracle.c").stat().st_mtime
    if need:
        subprocess.run(["make", "-C", str(_HERE), "libstereo_oracle.so"], check=True, capture_output=True)
    ref_src = Path("/root/reference/ProblemSets/ps2_cpp/lib/DisparitySSD.cpp")
    if ref_src.exists() and (force or not _REF.exists() or not _REF_NCC.exists()):
        subprocess.run(["make", "-C", str(_HERE), "ref"], check=True, capture_output=True)


_lib = None
_ref = {}


def _load():
    global _lib
    if _lib is None:
        if not _LIB.exists():
            build()
        _lib = C.CDLL(str(_LIB))
        fp, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        for name, last in (("oracle_disparity_ssd", i32p), ("oracle_disparity_ssd_fast", i32p),
                           ("oracle_disparity_ncorr", fp), ("oracle_disparity_ncorr_fast", fp)):
            fn = getattr(_lib, name)
            fn.restype = C.c_int
            fn.argtypes = [fp, C.c_size_t, fp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, last]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_num_threads.argtypes = [C.c_int]
    return _lib


def have_ref() -> bool:
    return _REF.exists()


def have_ref_ncc() -> bool:
    return _REF_NCC.exists()


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2:
        raise OracleError("images must be 2-D")
    return a


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _run(name, ref, tgt, R, dmin, dmax, want_aux, aux_dtype):
    lib = _load()
    ref, tgt = _f32(ref), _f32(tgt)
    if ref.shape != tgt.shape:
        raise OracleError("shape mismatch")
    rows, cols = ref.shape
    disp = np.empty((rows, cols), np.int32)
    aux = np.empty((rows, cols), aux_dtype) if want_aux else None
    auxp = aux.ctypes.data_as(C.POINTER(C.c_float if aux_dtype == np.float32 else C.c_int32)) if want_aux else None
    st = getattr(lib, name)(_fp(ref), cols, _fp(tgt), cols, rows, cols, int(R), int(dmin), int(dmax),
                            disp.ctypes.data_as(C.POINTER(C.c_int32)), auxp)
    if st != 0:
        raise OracleError(f"{name} failed with status {st}")
    return (disp, aux) if want_aux else disp


def ssd(ref, tgt, window_rad, min_disp, max_disp, return_cost=False):
    """Literal restatement of serial::disparitySSD (DisparitySSD.cpp:19-59); int32 disparities."""
    return _run("oracle_disparity_ssd", ref, tgt, window_rad, min_disp, max_disp, return_cost, np.int32)


def ssd_fast(ref, tgt, window_rad, min_disp, max_disp, return_cost=False):
    """Same results as :func:`ssd` in O(rows*cols*D)."""
    return _run("oracle_disparity_ssd_fast", ref, tgt, window_rad, min_disp, max_disp, return_cost, np.int32)


def ncorr(ref, tgt, window_rad, min_disp, max_disp, return_score=False):
    """Restatement of serial::disparityNCorr (DisparityNCorr.cpp:27-68) + TM_CCORR_NORMED."""
    return _run("oracle_disparity_ncorr", ref, tgt, window_rad, min_disp, max_disp, return_score, np.float32)


def ncorr_fast(ref, tgt, window_rad, min_disp, max_disp, return_score=False):
    """Same results as :func:`ncorr` in O(rows*cols*D) for integer-valued images (running sums in double)."""
    return _run("oracle_disparity_ncorr_fast", ref, tgt, window_rad, min_disp, max_disp, return_score, np.float32)


def refgpu(cost, ref, tgt, window_rad, min_disp, max_disp, return_best=False):
    """What the reference's GPU kernels compute (DisparitySSD.cu:27-141 / DisparityNCorr.cu:28-175):
    cost 0 = SSD, 1 = NCC; int32 disparities (-1 where nothing won) and, on request, the kernels' running-best map."""
    lib = _load()
    ref, tgt = _f32(ref), _f32(tgt)
    if ref.shape != tgt.shape:
        raise OracleError("shape mismatch")
    rows, cols = ref.shape
    disp = np.empty((rows, cols), np.int32)
    best = np.empty((rows, cols), np.float32)
    fn = lib.oracle_disparity_refgpu
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_float), C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                   C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    st = fn(int(cost), _fp(ref), cols, _fp(tgt), cols, rows, cols, int(window_rad), int(min_disp), int(max_disp),
            disp.ctypes.data_as(C.POINTER(C.c_int32)), best.ctypes.data_as(C.POINTER(C.c_float)))
    if st != 0:
        raise OracleError(f"oracle_disparity_refgpu failed with status {st}")
    return (disp, best) if return_best else disp


def narrow_i8(disp):
    """The reference's ``disparity.at<char>(...) = int`` store (DisparitySSD.cpp:59)."""
    return (np.asarray(disp).astype(np.int64) & 0xFF).astype(np.uint8).view(np.int8)


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def set_num_threads(n: int) -> None:
    _load().oracle_set_num_threads(int(n))


def ref_ssd(ref, tgt, window_rad, min_disp, max_disp, opt: str = "O2"):
    """The reference's own compiled serial::disparitySSD; returns int8 (CV_8SC1) like the reference."""
    path = _REF if opt == "O2" else _REF_O0
    if not path.exists():
        raise OracleError(f"{path} not built (reference sources absent?)")
    lib = _ref.get(opt)
    if lib is None:
        lib = C.CDLL(str(path))
        lib.ref_serial_disparity_ssd.restype = C.c_int
        lib.ref_serial_disparity_ssd.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int,
                                                 C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int8)]
        _ref[opt] = lib
    ref, tgt = _f32(ref), _f32(tgt)
    rows, cols = ref.shape
    out = np.empty((rows, cols), np.int8)
    st = lib.ref_serial_disparity_ssd(_fp(ref), _fp(tgt), rows, cols, int(window_rad), int(min_disp), int(max_disp),
                                      out.ctypes.data_as(C.POINTER(C.c_int8)))
    if st != 0:
        raise OracleError(f"ref_serial_disparity_ssd failed with status {st}")
    return out


def ref_ncorr(ref, tgt, window_rad, min_disp, max_disp):
    """The reference's own compiled serial::disparityNCorr (DisparityNCorr.cpp:12-71) over the shim's
    matchTemplate; returns int8 (CV_8SC1) like the reference.  O(rows*cols*D*w^2): small cases only."""
    if not _REF_NCC.exists():
        raise OracleError(f"{_REF_NCC} not built (reference sources absent?)")
    lib = _ref.get("ncc")
    if lib is None:
        lib = C.CDLL(str(_REF_NCC))
        lib.ref_serial_disparity_ncorr.restype = C.c_int
        lib.ref_serial_disparity_ncorr.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int,
                                                   C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int8)]
        _ref["ncc"] = lib
    ref, tgt = _f32(ref), _f32(tgt)
    if ref.shape != tgt.shape:
        raise OracleError("shape mismatch")
    rows, cols = ref.shape
    out = np.empty((rows, cols), np.int8)
    st = lib.ref_serial_disparity_ncorr(_fp(ref), _fp(tgt), rows, cols, int(window_rad), int(min_disp), int(max_disp),
                                        out.ctypes.data_as(C.POINTER(C.c_int8)))
    if st != 0:
        raise OracleError(f"ref_serial_disparity_ncorr failed with status {st}")
    return out
