"""The ps2 executable (apps/ps2): YAML subset, PNG I/O and the preprocessing/post-processing around the
matcher, pinned against executed OpenCV (python cv2); and, on the GPU box, the whole five-problem run
against oracle + cv2 on synthetic stand-ins of the bundled pairs (their pixels are Git-LFS stubs)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from introtocomputervision_b200 import synth

cv2 = pytest.importorskip("cv2")
ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "apps" / "ps2" / "ps2"

YAML = """---
# Problem Set 2 configuration (same keys as the reference's config/ps2.yaml)
images:
  pair0-L: {d}/pair0-L.png
  pair0-R: {d}/pair0-R.png   # trailing comment
  pair1-L: {d}/pair1-L.png
  pair1-R: {d}/pair1-R.png
  pair2-L: {d}/pair2-L.png
  pair2-R: {d}/pair2-R.png

output_dir: {d}/ps2_output
use_gpu_disparity: true
opencv_gray_shift: 15

problem_1_ssd:
  window_radius: 6
  disparity_range: 3
problem_2_ssd:
  window_radius: 7
  disparity_range: {rng}
problem_3_ssd:
  window_radius: 7
  disparity_range: {rng}
problem_4_ncorr:
  window_radius: 7
  disparity_range: {rng}
problem_5_ncorr:
  window_radius: 7
  disparity_range: {rng5}
...
"""


def build_exe():
    subprocess.run(["make", "-C", str(EXE.parent)], check=True, capture_output=True)
    return EXE


def colourise(gray, seed):
    """A 3-channel image whose OpenCV 'RGB2GRAY applied to BGR data' is close to `gray` but not trivially so."""
    rng = np.random.default_rng(seed)
    img = np.repeat(gray[:, :, None], 3, axis=2).astype(np.int32) + rng.integers(-6, 7, gray.shape + (3,))
    return np.clip(img, 0, 255).astype(np.uint8)


def test_selftest_against_opencv(tmp_path):
    exe = build_exe()
    d = tmp_path
    (d / "in.yaml").write_text(YAML.format(d="/data", rng=95, rng5=80))
    rng = np.random.default_rng(1)
    gray = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    colour = rng.integers(0, 256, (41, 67, 3), dtype=np.uint8)
    cv2.imwrite(str(d / "gray.png"), gray)
    cv2.imwrite(str(d / "colour.png"), colour)
    disp = rng.integers(-95, 1, (41, 67)).astype(np.int8)
    with open(d / "disp.bin", "wb") as f:
        f.write(np.array(disp.shape, np.int32).tobytes())
        f.write(disp.tobytes())
    res = subprocess.run([str(exe), "--selftest", str(d)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # YAML subset
    kv = dict(line.split("=", 1) for line in (d / "yaml.txt").read_text().splitlines())
    assert kv["images.pair0-R"] == "/data/pair0-R.png" and kv["output_dir"] == "/data/ps2_output"
    assert kv["use_gpu_disparity"] == "true" and kv["problem_4_ncorr.window_radius"] == "7"
    assert kv["problem_5_ncorr.disparity_range"] == "80" and len(kv) == 6 + 3 + 10
    # imread(IMREAD_UNCHANGED): gray stays 1 channel, colour comes back as B,G,R
    assert np.array_equal(np.fromfile(d / "gray.raw", np.uint8).reshape(gray.shape), cv2.imread(str(d / "gray.png"), cv2.IMREAD_UNCHANGED))
    assert np.array_equal(np.fromfile(d / "colour.raw", np.uint8).reshape(colour.shape), cv2.imread(str(d / "colour.png"), cv2.IMREAD_UNCHANGED))
    assert np.array_equal(cv2.imread(str(d / "gray_out.png"), cv2.IMREAD_UNCHANGED), gray)          # imwrite round trip
    # cvtColor(COLOR_RGB2GRAY) on BGR data + convertTo(CV_32FC1)   (main.cpp:114-117)
    bgr = cv2.imread(str(d / "colour.png"), cv2.IMREAD_UNCHANGED)
    g = cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY).astype(np.float32)
    assert np.array_equal(np.fromfile(d / "rgb2gray.f32", np.float32).reshape(g.shape), g)      # OpenCV >= 3.4.2 coefficients
    p = bgr.astype(np.int64)                                                                    # OpenCV 3.4.1 (the reference's pin)
    g14 = ((p[..., 0] * 4899 + p[..., 1] * 9617 + p[..., 2] * 1868 + (1 << 13)) >> 14).astype(np.float32)
    got14 = np.fromfile(d / "rgb2gray14.f32", np.float32).reshape(g.shape)
    assert np.array_equal(got14, g14) and np.abs(got14 - g).max() <= 1
    # addNoise: cv::randn with the never-seeded default RNG state (cv2.setRNGSeed(0) == state 0xffffffff)
    cv2.setRNGSeed(0)
    n1 = g + cv2.randn(np.empty(g.shape, np.float32), 0, 10)
    n2 = g + cv2.randn(np.empty(g.shape, np.float32), 0, 10)
    assert np.array_equal(np.fromfile(d / "noisy1.f32", np.float32).reshape(g.shape), n1)
    assert np.array_equal(np.fromfile(d / "noisy2.f32", np.float32).reshape(g.shape), n2)
    assert np.array_equal(np.fromfile(d / "contrast.f32", np.float32).reshape(g.shape), g * np.float32(1.1))
    # normalize(NORM_MINMAX -> CV_8UC1) and the inverted copy
    n8 = cv2.normalize(disp, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1)
    assert np.array_equal(np.fromfile(d / "norm.u8", np.uint8).reshape(disp.shape), n8)
    assert np.array_equal(np.fromfile(d / "inv.u8", np.uint8).reshape(disp.shape), 255 - n8)


def test_bad_config_exits_like_the_reference(tmp_path):
    exe = build_exe()
    (tmp_path / "bad.yaml").write_text("images:\n  pair0-L: /nonexistent.png\n")
    res = subprocess.run([str(exe), str(tmp_path / "bad.yaml")], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode != 0 and "Configuration load failed!" in res.stdout


@pytest.mark.gpu
def test_full_run_matches_oracle_pipeline(tmp_path):
    exe = build_exe()
    d = tmp_path
    rng_p, rng5 = 31, 24                     # smaller ranges than the reference's 95 / 80 keep the oracle quick
    L0, R0, _ = synth.make_pair(128, 128, 4, 10)           # pair0 stand-in: 128 x 128 (ps2_cpu.log:6), single channel
    L1, R1, _ = synth.make_pair(96, 160, rng_p + 1, 11)    # pair1 / pair2 stand-ins: colour files
    L2, R2, _ = synth.make_pair(80, 144, rng5 + 1, 12)
    cv2.imwrite(str(d / "pair0-L.png"), L0), cv2.imwrite(str(d / "pair0-R.png"), R0)
    for name, img, seed in (("pair1-L", L1, 1), ("pair1-R", R1, 2), ("pair2-L", L2, 3), ("pair2-R", R2, 4)):
        cv2.imwrite(str(d / f"{name}.png"), colourise(img, seed))
    (d / "ps2.yaml").write_text(YAML.format(d=str(d), rng=rng_p, rng5=rng5))
    res = subprocess.run([str(exe), str(d / "ps2.yaml")], capture_output=True, text=True, cwd=d)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Problem 5 runtime" in res.stdout and "Total runtime" in res.stdout
    log = (d / "ps2.log").read_text()
    assert "GPU warmup done" in log and "disparityNCorrKernel execution took" in log

    def gray(name):
        return cv2.cvtColor(cv2.imread(str(d / f"{name}.png"), cv2.IMREAD_UNCHANGED), cv2.COLOR_RGB2GRAY).astype(np.float32)

    def expect(stem, fn, l, r, R, rng):
        dl = oracle.narrow_i8(fn(l, r, R, -rng, 0))
        dr = oracle.narrow_i8(fn(r, l, R, 0, rng))
        for suffix, disp in (("1", dl), ("2", dr)):
            want = cv2.normalize(disp, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1)
            got = cv2.imread(str(d / "ps2_output" / f"{stem}-{suffix}.png"), cv2.IMREAD_UNCHANGED)
            agree = float(np.mean(got == want))
            assert agree >= (1.0 if fn is oracle.ssd else 0.999), (stem, suffix, agree)
            if suffix == "1" and not stem.startswith("ps2-1"):
                inv = cv2.imread(str(d / "ps2_output" / f"{stem}-1-inverted.png"), cv2.IMREAD_UNCHANGED)
                assert np.array_equal(inv, 255 - got)

    expect("ps2-1-a", oracle.ssd, L0.astype(np.float32), R0.astype(np.float32), 6, 3)
    l1, r1, l2, r2 = gray("pair1-L"), gray("pair1-R"), gray("pair2-L"), gray("pair2-R")
    expect("ps2-2-a", oracle.ssd, l1, r1, 7, rng_p)
    cv2.setRNGSeed(0)                                       # the RNG runs on from problem 3 into problem 4
    noisy = lambda a: a + cv2.randn(np.empty(a.shape, np.float32), 0, 10)
    l3, r3 = noisy(l1), noisy(r1)
    expect("ps2-3-a", oracle.ssd, l3, r3, 7, rng_p)
    expect("ps2-3-b", oracle.ssd, l1 * np.float32(1.1), r1 * np.float32(1.1), 7, rng_p)
    expect("ps2-4-a", oracle.ncorr, l1, r1, 7, rng_p)
    l4, r4 = noisy(l1), noisy(r1)
    expect("ps2-4-b", oracle.ncorr, l4, r4, 7, rng_p)
    expect("ps2-4-c", oracle.ncorr, l1 * np.float32(1.1), r1 * np.float32(1.1), 7, rng_p)
    expect("ps2-5-a", oracle.ncorr, l2, r2, 7, rng5)
