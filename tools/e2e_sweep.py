#!/usr/bin/env python
"""Times the host-buffer batch entry points (the bench's `e2e` path) over pairs-per-call and bands-per-pair
settings on one GPU.  A tuning aid for pipe_bands() / pipe_chunk_pairs() in csrc/stereo_b200.cu; prints one JSON
line per setting.  Not a bench line: wall clock around synchronous calls, no L2 discipline needed (every call
re-uploads its inputs from pinned host memory)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import introtocomputervision_b200 as sb
    from introtocomputervision_b200 import _capi, synth
    from bench import WORKLOADS

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="4k_d256_w11", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--bands", type=int, nargs="+", default=[0, 1, 2, 4, 8])
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--cost", default="ssd")
    ap.add_argument("--kinds", nargs="+", default=["f32", "u8"])
    ap.add_argument("--threads", type=int, nargs="+", default=[0], help="host packing threads (0: the library's default)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rows, cols, nd, R = wl["rows"], wl["cols"], wl["ndisp"], wl["R"]
    lib = _capi.lib()
    ctx = sb.Context(0)
    cost = sb.COST_SSD if args.cost == "ssd" else sb.COST_NCORR
    elem_dtype, elem = (torch.int8, 1) if nd <= 128 else (torch.int16, 2)
    Bmax = max(args.pairs)
    Ls, Rs = [], []
    for i in range(Bmax):
        L, Rt, _ = synth.make_pair(rows, cols, nd, wl["seed"] + i)
        Ls.append(L), Rs.append(Rt)
    hu = (torch.from_numpy(np.stack(Ls)).pin_memory(), torch.from_numpy(np.stack(Rs)).pin_memory())
    hf = (hu[0].to(torch.float32).pin_memory(), hu[1].to(torch.float32).pin_memory())
    h_dl = torch.empty((Bmax, rows, cols), dtype=elem_dtype).pin_memory()
    h_dr = torch.empty((Bmax, rows, cols), dtype=elem_dtype).pin_memory()
    for kind in args.kinds:
        src, px, fn = (hf, 4, lib.stereo_disparity_pair_batch_f32_host) if kind == "f32" else (hu, 1, lib.stereo_disparity_pair_batch_u8_host)
        for B in args.pairs:
            for bands, threads in [(b, t) for b in args.bands for t in (args.threads if kind == "f32" else args.threads[:1])]:
                ctx.set_pipe_bands(bands)
                lib.stereo_ctx_set_host_threads(ctx.handle, threads)

                def call():
                    rc = fn(ctx.handle, cost, B, src[0].data_ptr(), src[1].data_ptr(), cols * px, rows * cols * px, rows, cols, R,
                            nd - 1, h_dl.data_ptr(), h_dr.data_ptr(), cols * elem, rows * cols * elem, elem)
                    if rc != 0:
                        raise RuntimeError(_capi.last_error())

                call(), call()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    call()
                dt = (time.perf_counter() - t0) / args.steps
                print(json.dumps({"workload": args.workload, "cost": args.cost, "kind": kind, "pairs_per_call": B, "bands": bands, "host_threads": lib.stereo_ctx_host_threads(ctx.handle),
                                  "ms_per_call": round(dt * 1e3, 3), "ms_per_pair": round(dt * 1e3 / B, 3),
                                  "Mpix_disp_per_s": round(B * 2 * rows * cols * nd / dt / 1e6, 1),
                                  "launches": ctx.last_launches}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
