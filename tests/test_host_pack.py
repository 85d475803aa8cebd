"""Host-side float32 -> u8 packing of the pipelined CV_32FC1 entry points (csrc/host_pack.cpp): pure host code, checked on
the CPU.  Both code paths: the AVX-512 row (non-temporal stores) and the portable one (STEREO_NO_AVX512=1 in a child)."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]

CHILD = r"""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, %r)
from introtocomputervision_b200 import _capi
lib = _capi.lib()
rng = np.random.default_rng(7)

def pack(a, threads, dst_pad=0):
    rows, cols = a.shape
    out = np.full((rows, cols + dst_pad), 0xEE, np.uint8)
    ok = C.c_int(-1)
    st = lib.stereo_host_pack_f32_u8(a.ctypes.data, a.strides[0], out.ctypes.data, out.strides[0], rows, cols, threads, C.byref(ok))
    assert st == 0
    return out[:, :cols], ok.value, out[:, cols:]

for rows, cols in ((1, 1), (3, 63), (5, 64), (7, 65), (33, 200), (64, 1280), (9, 4099)):
    for threads in (1, 3, 8):
        img = rng.integers(0, 256, (rows, cols)).astype(np.float32)
        got, ok, pad = pack(img, threads, dst_pad=5)
        assert ok == 1 and np.array_equal(got, img.astype(np.uint8)), (rows, cols, threads)
        assert np.all(pad == 0xEE)                              # nothing written past a row
        view = np.ascontiguousarray(np.pad(img, ((0, 0), (3, 2))))[:, 3:3 + cols]       # strided, unaligned source rows
        got, ok, _ = pack(view, threads)
        assert ok == 1 and np.array_equal(got, img.astype(np.uint8))
        for bad in (0.5, -1.0, 256.0, 255.0001, np.nan, np.inf, -np.inf, 1e10, -0.25):
            b = img.copy()
            b[rng.integers(0, rows), rng.integers(0, cols)] = bad
            assert pack(b, threads)[1] == 0, (rows, cols, threads, bad)
        z = img.copy(); z[0, 0] = -0.0                          # minus zero is 0
        assert pack(z, threads)[1] == 1
print("pack ok")
"""


@pytest.mark.parametrize("no_avx512", [False, True])
def test_host_pack(no_avx512):
    env = dict(os.environ)
    if no_avx512:
        env["STEREO_NO_AVX512"] = "1"
    res = subprocess.run([sys.executable, "-c", CHILD % str(ROOT)], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0 and "pack ok" in res.stdout, res.stdout + res.stderr


@pytest.mark.parametrize("threads", [2, 5, 16])
def test_host_pool_many_dispatches(threads):
    """The polling worker pool: thousands of back-to-back dispatches on one pool (every task exactly once), including
    dispatches that find the workers asleep."""
    sys.path.insert(0, str(ROOT))
    from introtocomputervision_b200 import _capi
    assert _capi.lib().stereo_host_pool_selftest(threads, 3000) == 0, _capi.last_error()
