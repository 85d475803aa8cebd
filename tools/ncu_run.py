#!/usr/bin/env python
"""A short run of the device pair entry point for ncu (one process, one GPU): `--reps` calls of
stereo_disparity_pair_batch_u8_device on `--pairs` synthetic pairs of a bench workload."""
import argparse
import sys
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
    ap.add_argument("--pairs", type=int, default=2)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--cost", default="ssd")
    ap.add_argument("--no-fuse", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rows, cols, nd, R = wl["rows"], wl["cols"], wl["ndisp"], wl["R"]
    ctx = sb.Context(0)
    if args.no_fuse:
        ctx.set_fuse_pairs(False)
    cost = sb.COST_SSD if args.cost == "ssd" else sb.COST_NCORR
    elem_dtype, elem = (torch.int8, 1) if nd <= 128 else (torch.int16, 2)
    B = args.pairs
    Ls, Rs = [], []
    for i in range(B):
        L, Rt, _ = synth.make_pair(rows, cols, nd, wl["seed"] + i)
        Ls.append(L), Rs.append(Rt)
    dl = torch.from_numpy(np.stack(Ls)).cuda()
    dr = torch.from_numpy(np.stack(Rs)).cuda()
    out = torch.empty((2, B, rows, cols), dtype=elem_dtype, device="cuda")
    for _ in range(args.reps):
        rc = _capi.lib().stereo_disparity_pair_batch_u8_device(
            ctx.handle, cost, B, dl.data_ptr(), dr.data_ptr(), cols, rows * cols, rows, cols, R, nd - 1,
            out[0].data_ptr(), out[1].data_ptr(), cols * elem, rows * cols * elem, elem, None)
        assert rc == 0, _capi.last_error()
    ctx.synchronize()
    print("fused pairs in the last call:", ctx.last_fused_pairs, "launches:", ctx.last_launches)
    ctx.close()


if __name__ == "__main__":
    main()
