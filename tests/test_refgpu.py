"""Reference-GPU-semantics mode: stereo_disparity_refgpu_* against the C restatement of the
reference's GPU kernels (oracle.refgpu), plus CPU-only sanity checks of that restatement."""
import numpy as np
import pytest

import oracle
from introtocomputervision_b200 import synth


def test_refgpu_oracle_known_answers():
    # right(x) = left(x + 4): the L->R interior disparity is -4 in both semantics; the GPU semantics mark nothing -1 there
    rng = np.random.default_rng(3)
    left = rng.integers(0, 256, (50, 120)).astype(np.float32)
    right = np.empty_like(left)
    right[:, :] = left[:, np.minimum(np.arange(120) + 4, 119)]
    d = oracle.refgpu(0, left, right, 3, -10, 0)
    assert np.all(d[10:40, 20:100] == -4)
    d, best = oracle.refgpu(1, left, right, 3, -10, 0, return_best=True)
    assert np.all(d[10:40, 20:100] == -4) and np.all(best[10:40, 20:100] > 0.999)
    # the 5e6 threshold: windows whose every candidate costs more keep -1 (DisparitySSD.cu:16,88,177)
    a = np.zeros((45, 64), np.float32)
    b = np.full((45, 64), 255, np.float32)
    assert np.all(oracle.refgpu(0, a, b, 7, -3, 0) == -1)                  # 15 x 14 x 255^2 = 13.7e6 > 5e6
    assert np.all(oracle.refgpu(0, a, b, 2, -3, 0) == -3)                  # 5 x 4 x 255^2 = 1.3e6: first candidate wins, ties keep it
    # NCC of an all-zero window is 0/0 = NaN and never beats 0: -1 (DisparityNCorr.cu:16,108)
    assert np.all(oracle.refgpu(1, a, b, 2, -3, 0) == -1)
    # window = (2R+1) x 2R: R = 0 sums nothing, cost 0 < 5e6 for the first candidate everywhere
    assert np.all(oracle.refgpu(0, left, right, 0, -5, 0) == -5)


def test_refgpu_semantics_differ_from_the_cpu_semantics():
    L, Rt, _ = synth.make_pair(60, 200, 20, 5)
    Lf, Rf = L.astype(np.float32), Rt.astype(np.float32)
    g = oracle.refgpu(0, Lf, Rf, 4, -19, 0)
    c = oracle.ssd_fast(Lf, Rf, 4, -19, 0)
    interior = np.mean(g[8:-8, 30:-10] == c[8:-8, 30:-10])
    assert 0.5 < interior < 1.0            # mostly the same answers in the interior, not the same function


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,R,dmin,dmax", [(50, 120, 3, -10, 0), (128, 128, 6, -3, 0), (128, 128, 6, 0, 3), (97, 333, 7, -95, 0),
                                                    (41, 65, 2, 0, 40), (80, 64, 0, -5, 5), (40, 200, 5, -126, 0), (7, 9, 4, -2, 2)])
def test_refgpu_kernels_equal_the_restatement(ctx, rows, cols, R, dmin, dmax):
    import introtocomputervision_b200 as sb
    L, Rt, _ = synth.make_pair(rows, cols, max(2, min(cols // 2, 30)), 500 + rows + cols)
    for Lf, Rf in ((L.astype(np.float32), Rt.astype(np.float32)),
                   (synth.noisy_variant(L, 1), synth.noisy_variant(Rt, 2)),
                   (synth.contrast_variant(L), synth.contrast_variant(Rt))):
        for cost in (sb.COST_SSD, sb.COST_NCORR):
            d_ref, b_ref = oracle.refgpu(cost, Lf, Rf, R, dmin, dmax, return_best=True)
            d, b = ctx.disparity_refgpu(cost, Lf, Rf, R, dmin, dmax, return_best=True)
            assert ctx.last_path == sb.PATH_REFGPU
            bad = np.argwhere(d.astype(np.int32) != d_ref)
            assert bad.size == 0, f"cost {cost}: differs at {bad[:5].tolist()} (of {len(bad)})"
            assert np.array_equal(b.view(np.uint32), b_ref.view(np.uint32)), f"cost {cost}: running-best maps differ"


@pytest.mark.gpu
def test_refgpu_rejects_ranges_beyond_char(ctx):
    import introtocomputervision_b200 as sb
    img = np.zeros((8, 8), np.float32)
    with pytest.raises(sb.StereoError):
        ctx.disparity_refgpu(sb.COST_SSD, img, img, 1, -200, 0)
