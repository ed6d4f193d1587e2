"""CPU: pin the dense-stereo oracle (oracle/sgbm_restate.py) against live cv2.StereoSGBM 4.13.0 with the
reference's constants (visual_odometry.cpp:163-164) and against the committed golden disparity image."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import sgbm_restate as G

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cv(left, right, **kw):
    a = dict(minDisparity=0, numDisparities=96, blockSize=9, P1=8 * 81, P2=32 * 81, disp12MaxDiff=1, preFilterCap=63,
             uniquenessRatio=10, speckleWindowSize=100, speckleRange=32)
    a.update(kw)
    return cv2.StereoSGBM_create(**a).compute(left, right)


def _crop(pkg, seed, h, w, y0=100, x0=300):
    left, right, _ = pkg.synth.synth_pair(seed)
    return np.ascontiguousarray(left[y0:y0 + h, x0:x0 + w]), np.ascontiguousarray(right[y0:y0 + h, x0:x0 + w])


@pytest.mark.parametrize("seed,h,w", [(0, 64, 360), (1, 40, 250), (2, 9, 140), (3, 1, 120), (4, 2, 101)])
def test_sgbm_matches_cv2_on_synthetic_crops(pkg, seed, h, w):
    left, right = _crop(pkg, seed, h, w)
    out = G.sgbm_compute(left, right)
    ref = _cv(left, right)
    assert out.dtype == np.int16 and np.array_equal(out, ref)
    if h >= 40:
        assert (ref != -16).mean() > 0.3  # the crop really carries disparities, not only the invalid value


def test_sgbm_saturating_noise(pkg):
    # independent uniform noise drives the summed cost S into int16 saturation (32767)
    rng = np.random.default_rng(5)
    left = rng.integers(0, 256, (37, 131), dtype=np.uint8)
    right = rng.integers(0, 256, (37, 131), dtype=np.uint8)
    st = {}
    out = G.sgbm_compute(left, right, stages=st)
    assert st["S"].max() == 32767
    assert np.array_equal(out, _cv(left, right))


@pytest.mark.parametrize("kw", [dict(speckleWindowSize=0), dict(uniquenessRatio=0, disp12MaxDiff=100000, speckleWindowSize=0),
                                dict(uniquenessRatio=25), dict(disp12MaxDiff=3, speckleWindowSize=30, speckleRange=2)])
def test_sgbm_stage_switches(pkg, kw):
    # switching the post-passes off one at a time isolates WTA / uniqueness / LR check / speckle filter
    left, right = _crop(pkg, 6, 48, 300)
    p = G.Params(disp12_max_diff=kw.get("disp12MaxDiff", 1), uniqueness_ratio=kw.get("uniquenessRatio", 10),
                 speckle_window_size=kw.get("speckleWindowSize", 100), speckle_range=kw.get("speckleRange", 32))
    assert np.array_equal(G.sgbm_compute(left, right, p), _cv(left, right, **kw))


def test_sgbm_golden(pkg):
    left, right = _crop(pkg, 0, 64, 360)
    g = np.load(os.path.join(GOLD, "sgbm_pair0_crop64x360.npz"))["disp16"]
    assert np.array_equal(G.sgbm_compute(left, right), g)


def test_median_and_speckle_vs_cv2():
    rng = np.random.default_rng(0)
    d = (rng.integers(0, 40, (50, 70)) * 16).astype(np.int16)
    d[rng.random(d.shape) < 0.3] = -16
    assert np.array_equal(G.median3(d), cv2.medianBlur(d, 3))
    for size, diff in ((100, 512), (5, 16), (20, 64)):
        ref = d.copy()
        cv2.filterSpeckles(ref, -16, size, diff)
        assert np.array_equal(G.filter_speckles(d, -16, size, diff), ref)


def test_find_3d_truncates_like_mat_at():
    disp = np.full((10, 20), -1.0, np.float32)
    disp[3, 7] = 8.0
    kp = np.array([[7.9, 3.9], [8.0, 3.0]], np.float32)
    X = G.find_3d(kp, disp, 718.856, 718.856, 607.1928, 185.2157, 0.573)
    assert np.isclose(X[0, 2], 718.856 * 0.573 / 8.0) and X[1, 2] < 0  # invalid (-1) gives a negative depth


def test_sgbm_random_parameters_and_sizes(pkg):
    """seeded sweep over image sizes and the free parameters (penalties, uniqueness, LR tolerance, prefilter cap,
    speckle filter): the restatement must equal live cv2 everywhere, not only at the reference's constants"""
    rng = np.random.default_rng(2024)
    for trial in range(14):
        h = int(rng.integers(3, 48))
        w = int(rng.integers(104, 260))
        left, right = _crop(pkg, int(rng.integers(0, 10)), h, w, y0=int(rng.integers(0, 300)), x0=int(rng.integers(0, 900)))
        if trial % 4 == 3:  # decorrelate the pair now and then: many invalid / speckle pixels
            right = np.ascontiguousarray(np.roll(right, int(rng.integers(1, 7)), axis=0))
        P1 = int(rng.integers(1, 900))
        kw = dict(P1=P1, P2=P1 + int(rng.integers(1, 2500)), disp12MaxDiff=int(rng.integers(1, 5)),
                  preFilterCap=int(rng.integers(1, 64)), uniquenessRatio=int(rng.integers(0, 40)),
                  speckleWindowSize=int(rng.choice([0, 10, 50, 100, 200])), speckleRange=int(rng.integers(1, 40)))
        p = G.Params(P1=kw["P1"], P2=kw["P2"], disp12_max_diff=kw["disp12MaxDiff"], pre_filter_cap=kw["preFilterCap"],
                     uniqueness_ratio=kw["uniquenessRatio"], speckle_window_size=kw["speckleWindowSize"],
                     speckle_range=kw["speckleRange"])
        assert np.array_equal(G.sgbm_compute(left, right, p), _cv(left, right, **kw)), (trial, h, w, kw)
