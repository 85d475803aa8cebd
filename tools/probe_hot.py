#!/usr/bin/env python
"""Device-resident timing probe: whole-call and hot-kernel milliseconds for a list of problem shapes.

    python tools/probe_hot.py [--queued] rows,cols,ndisp,R,cost,pairs[,f32|noisy] ...   -> one JSON line per shape

--queued times the call with its launches already queued behind a running kernel (device time without the host's
launch cost); the default starts from an idle device, as a caller sees it.

u8 shapes go through stereo_disparity_pair_batch_u8_device, f32 / noisy ones (CV_32FC1 device images, 8-bit-valued or
with Gaussian noise) through stereo_disparity_pair_f32_device, pair by pair."""
import ctypes as C
import json
import statistics
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import introtocomputervision_b200 as sb
    from introtocomputervision_b200 import _capi, synth
    lib = _capi.lib()
    ctx = sb.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    queued = "--queued" in sys.argv        # leave the L2 flush in flight: the call's launches queue up behind it
    for spec in [a for a in sys.argv[1:] if not a.startswith("--")]:
        parts = spec.split(",")
        rows, cols, nd, R = (int(v) for v in parts[:4])
        cost_name, B = parts[4], int(parts[5])
        kind = parts[6] if len(parts) > 6 else "u8"
        cost = sb.COST_SSD if cost_name == "ssd" else sb.COST_NCORR
        L, Rt, _ = synth.make_pair(rows, cols, nd, 7)
        elem_dtype, elem = (torch.int8, 1) if nd <= 128 else (torch.int16, 2)
        if kind == "u8":
            d_l = torch.from_numpy(np.stack([L] * B)).to(dev)
            d_r = torch.from_numpy(np.stack([Rt] * B)).to(dev)
        else:
            Lf, Rf = (synth.noisy_variant(L, 1), synth.noisy_variant(Rt, 2)) if kind == "noisy" else (L.astype(np.float32), Rt.astype(np.float32))
            d_l = torch.from_numpy(np.stack([Lf] * B)).to(dev)
            d_r = torch.from_numpy(np.stack([Rf] * B)).to(dev)
        d_out = torch.empty((2, B, rows, cols), dtype=elem_dtype, device=dev)

        def call():
            if kind == "u8":
                rc = lib.stereo_disparity_pair_batch_u8_device(ctx.handle, cost, B, d_l.data_ptr(), d_r.data_ptr(), cols, rows * cols, rows, cols, R, nd - 1,
                                                               d_out[0].data_ptr(), d_out[1].data_ptr(), cols * elem, rows * cols * elem, elem, sp)
                if rc != 0:
                    raise RuntimeError(_capi.last_error())
            else:
                for i in range(B):
                    rc = lib.stereo_disparity_pair_f32_device(ctx.handle, cost, d_l[i].data_ptr(), cols * 4, d_r[i].data_ptr(), cols * 4, rows, cols, R, nd - 1,
                                                              d_out[0, i].data_ptr(), d_out[1, i].data_ptr(), cols * elem, elem, sp)
                    if rc != 0:
                        raise RuntimeError(_capi.last_error())

        for _ in range(3):
            call()
        torch.cuda.synchronize()
        ms_call, ms_hot = [], []
        for _ in range(15):
            flush.zero_()
            if queued:
                flush.zero_()
            else:
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            call()
            e1.record(stream)
            torch.cuda.synchronize()
            ms_call.append(e0.elapsed_time(e1))
            h, n = ctx.last_hot_kernel_ms()
            ms_hot.append(h)
        units = B * 2 * rows * cols * nd
        ms, hot = statistics.median(ms_call), statistics.median(ms_hot)
        print(json.dumps({"shape": spec, "path": ctx.last_path, "fused_pairs": ctx.last_fused_pairs, "launches": ctx.last_launches,
                          "ms_call": round(ms, 4), "ms_hot": round(hot, 4), "Tpixd_call": round(units / ms / 1e9, 3),
                          "Tpixd_hot": round(units / hot / 1e9, 3) if hot > 0 else None}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
