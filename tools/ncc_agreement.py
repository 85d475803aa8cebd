"""Diagnostic: agreement of the packed NCC path with the oracle on assorted images (GPU box)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import introtocomputervision_b200 as sb  # noqa: E402
from introtocomputervision_b200 import synth  # noqa: E402
import oracle  # noqa: E402

ctx = sb.Context(0)


def report(name, L, Rt, R, dmin, dmax):
    d_ref, s_ref = oracle.ncorr_fast(L.astype(np.float32), Rt.astype(np.float32), R, dmin, dmax, return_score=True)
    d, s = ctx.disparity(sb.COST_NCORR, L, Rt, R, dmin, dmax, dtype=np.int32, return_best=True)
    bad = d != d_ref
    rel = np.abs(s - s_ref) / np.maximum(np.abs(s_ref), 1e-30)
    print(f"{name:34s} R={R} [{dmin},{dmax}] path={ctx.last_path} mismatches={int(bad.sum())}/{bad.size} "
          f"({100 * bad.mean():.4f}%) max score rel err={rel.max():.2e}", flush=True)
    if bad.any():
        ys, xs = np.nonzero(bad)
        for y, x in list(zip(ys, xs))[:4]:
            print(f"     (y={y},x={x}) ours d={d[y, x]} s={s[y, x]:.9f}  oracle d={d_ref[y, x]} s={s_ref[y, x]:.9f}")


L, Rt, _ = synth.make_pair(70, 130, 20, 9)
report("band-test pair L->R", L, Rt, 4, -19, 0)
report("band-test pair R->L", Rt, L, 4, 0, 19)
L, Rt, _ = synth.make_pair(511, 640, 96, 11)
report("pair1 stand-in", L, Rt, 7, -95, 0)
report("pair1 stand-in R->L", Rt, L, 7, 0, 95)
L, Rt, _ = synth.make_pair(540, 1920, 128, 1001)
report("1080p half", L, Rt, 4, -127, 0)
yy, xx = np.mgrid[0:200, 0:640]
rng = np.random.default_rng(5)
for amp in (1, 3, 10):
    Ls = np.clip(60 + 0.3 * xx + 0.2 * yy + rng.integers(-amp, amp + 1, xx.shape), 0, 255).astype(np.uint8)
    Rs = np.clip(60 + 0.3 * (xx + 7) + 0.2 * yy + rng.integers(-amp, amp + 1, xx.shape), 0, 255).astype(np.uint8)
    report(f"smooth gradient + noise amp {amp}", Ls, Rs, 4, -40, 0)
    report(f"smooth gradient + noise amp {amp}", Ls, Rs, 7, -40, 0)
dark = (rng.integers(0, 4, (100, 300))).astype(np.uint8)
report("dark 0..3 noise", dark, np.roll(dark, 5, axis=1), 3, -20, 0)
