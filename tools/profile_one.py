"""One device-resident call of the hot path, for ncu captures:  python tools/profile_one.py [workload] [reps] [ssd|ncc]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import introtocomputervision_b200 as sb  # noqa: E402
from introtocomputervision_b200 import _capi, synth  # noqa: E402
from bench import WORKLOADS  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "4k_d256_w11"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cost = sb.COST_NCORR if (len(sys.argv) > 3 and sys.argv[3] == "ncc") else sb.COST_SSD
rows, cols, nd, R = wl["rows"], wl["cols"], wl["ndisp"], wl["R"]
L, Rt, _ = synth.make_pair(rows, cols, nd, wl["seed"])
ctx = sb.Context(0)
dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(Rt).cuda()
out = torch.empty((rows, cols), dtype=torch.int16, device="cuda")
st = torch.cuda.Stream()
for _ in range(reps):
    rc = _capi.lib().stereo_disparity_u8_device(ctx.handle, cost, dl.data_ptr(), cols, dr.data_ptr(), cols, rows, cols, R,
                                                -(nd - 1), 0, out.data_ptr(), cols * 2, 2, None, 0, C.c_void_p(st.cuda_stream))
    assert rc == 0, _capi.last_error()
    ctx.synchronize(st.cuda_stream)
    print("kernel_ms", ctx.last_kernel_ms, "hot", ctx.last_hot_kernel_ms())
