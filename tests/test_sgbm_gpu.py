"""GPU: dense stereo (K18-K23) through the C-ABI against the oracle (oracle/sgbm_restate.py, pinned to cv2 4.13.0
StereoSGBM), stage by stage and end to end; the reference call is visual_odometry.cpp:163-168."""
import os

import numpy as np
import pytest

from oracle import sgbm_restate as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _crop(pkg, seed, h, w, y0=100, x0=300):
    left, right, _ = pkg.synth.synth_pair(seed)
    return np.ascontiguousarray(left[y0:y0 + h, x0:x0 + w]), np.ascontiguousarray(right[y0:y0 + h, x0:x0 + w])


@pytest.mark.parametrize("seed,h,w", [(0, 64, 360), (1, 40, 250), (2, 9, 140), (3, 1, 120), (4, 2, 101), (5, 53, 333)])
def test_sgbm_stages_and_result(pkg, gpu_ctx, seed, h, w):
    left, right = _crop(pkg, seed, h, w)
    st = {}
    ref = G.sgbm_compute(left, right, stages=st)
    S, parts = G.aggregate(st["C"], G.Params(), return_parts=True)
    # vertical sweep only: the three path volumes are still intact
    gpu_ctx.sgbm_debug_stop_after(1)
    try:
        gpu_ctx.sgbm_compute(left, right)
        C = gpu_ctx.sgbm_debug_volume(0, 0, h, w)
        assert np.array_equal(C.astype(np.int32) + 2592, st["C"]), "cost volume"
        for k in (1, 2, 3):
            L = gpu_ctx.sgbm_debug_volume(0, k, h, w)
            assert np.array_equal(L.astype(np.int32), parts[f"L{k}"]), f"path volume {k}"
    finally:
        gpu_ctx.sgbm_debug_stop_after(0)
    out, outf = gpu_ctx.sgbm_compute(left, right, want_float=True)
    assert np.array_equal(gpu_ctx.sgbm_debug_volume(0, 1, h, w).astype(np.int32), parts["S4"]), "S4"
    assert np.array_equal(gpu_ctx.sgbm_debug_volume(0, 4, h, w), st["raw"]), "raw disparity"
    assert np.array_equal(gpu_ctx.sgbm_debug_volume(0, 5, h, w), st["med"]), "median"
    assert out.dtype == np.int16 and np.array_equal(out, ref)
    assert np.array_equal(outf, G.disparity_float(ref))


def test_sgbm_golden_cv2(pkg, gpu_ctx):
    left, right = _crop(pkg, 0, 64, 360)
    g = np.load(os.path.join(GOLD, "sgbm_pair0_crop64x360.npz"))["disp16"]
    assert np.array_equal(gpu_ctx.sgbm_compute(left, right), g)


def test_sgbm_live_cv2_full_size_batch(pkg, gpu_ctx):
    cv2 = pytest.importorskip("cv2")
    sg = cv2.StereoSGBM_create(0, 96, 9, 8 * 81, 32 * 81, 1, 63, 10, 100, 32)
    pairs = [pkg.synth.synth_pair(s)[:2] for s in (0, 7, 9)]
    L = np.stack([p[0] for p in pairs])
    R = np.stack([p[1] for p in pairs])
    out = gpu_ctx.sgbm_compute(L, R)
    for i in range(len(pairs)):
        ref = sg.compute(L[i], R[i])
        assert np.array_equal(out[i], ref), f"pair {i}: {(out[i] != ref).sum()} px differ"
        assert (ref != -16).mean() > 0.8


def test_sgbm_saturating_noise_and_params(gpu_ctx):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    left = rng.integers(0, 256, (37, 131), dtype=np.uint8)
    right = rng.integers(0, 256, (37, 131), dtype=np.uint8)
    assert np.array_equal(gpu_ctx.sgbm_compute(left, right),
                          cv2.StereoSGBM_create(0, 96, 9, 648, 2592, 1, 63, 10, 100, 32).compute(left, right))
    # free parameters: penalties, uniqueness, LR tolerance, prefilter cap, speckle filter
    for kw in (dict(P1=100, P2=1000), dict(uniquenessRatio=0, disp12MaxDiff=100000, speckleWindowSize=0),
               dict(uniquenessRatio=25, preFilterCap=31), dict(disp12MaxDiff=3, speckleWindowSize=30, speckleRange=2)):
        a = dict(minDisparity=0, numDisparities=96, blockSize=9, P1=648, P2=2592, disp12MaxDiff=1, preFilterCap=63,
                 uniquenessRatio=10, speckleWindowSize=100, speckleRange=32)
        a.update(kw)
        ref = cv2.StereoSGBM_create(**a).compute(left, right)
        p = gpu_ctx.sgbm_params(P1=a["P1"], P2=a["P2"], disp12_max_diff=a["disp12MaxDiff"], pre_filter_cap=a["preFilterCap"],
                                uniqueness_ratio=a["uniquenessRatio"], speckle_window_size=a["speckleWindowSize"],
                                speckle_range=a["speckleRange"])
        assert np.array_equal(gpu_ctx.sgbm_compute(left, right, p), ref), kw


def test_sgbm_rejects_what_opencv_rejects(pkg, gpu_ctx):
    left = np.zeros((20, 100), np.uint8)  # width - 96 <= 4: OpenCV raises (stereosgbm.cpp:511)
    with pytest.raises(pkg.ffi.VslamError):
        gpu_ctx.sgbm_compute(left, left)
    with pytest.raises(pkg.ffi.VslamError):
        gpu_ctx.sgbm_compute(np.zeros((20, 200), np.uint8), np.zeros((20, 200), np.uint8),
                             gpu_ctx.sgbm_params(num_disparities=64))


def test_sgbm_device_entry_matches_host_entry(pkg, gpu_ctx):
    import torch
    left, right, _ = pkg.synth.synth_pair(2)
    h, w = left.shape
    dl, dr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    d16 = torch.empty((h, w), dtype=torch.int16, device="cuda")
    df = torch.empty((h, w), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.sgbm_compute_dev(dl, dr, 1, w, h, w, w * h, d16, df)
    gpu_ctx.synchronize()
    host = gpu_ctx.sgbm_compute(left, right)
    assert np.array_equal(d16.cpu().numpy(), host)
    assert np.array_equal(df.cpu().numpy(), host.astype(np.float32) / 16)


def test_sgbm_batch_larger_than_one_chunk_and_row_pitch(pkg, gpu_ctx):
    """11 pairs cross the 8-pair scratch chunk; the device entry reads rows at a pitch wider than the image"""
    import torch
    cv2 = pytest.importorskip("cv2")
    sg = cv2.StereoSGBM_create(0, 96, 9, 8 * 81, 32 * 81, 1, 63, 10, 100, 32)
    n, h, w, pitch = 11, 30, 150, 160
    L = np.stack([_crop(pkg, s, h, w, y0=20 * (s % 5), x0=100 + 40 * s)[0] for s in range(n)])
    R = np.stack([_crop(pkg, s, h, w, y0=20 * (s % 5), x0=100 + 40 * s)[1] for s in range(n)])
    ref = np.stack([sg.compute(L[i], R[i]) for i in range(n)])
    assert np.array_equal(gpu_ctx.sgbm_compute(L, R), ref)
    padl = np.full((n, h, pitch), 255, np.uint8)
    padr = np.full((n, h, pitch), 255, np.uint8)
    padl[:, :, :w], padr[:, :, :w] = L, R
    dl, dr = torch.from_numpy(padl).cuda(), torch.from_numpy(padr).cuda()
    d16 = torch.empty((n, h, w), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.sgbm_compute_dev(dl, dr, n, w, h, pitch, pitch * h, d16)
    gpu_ctx.synchronize()
    assert np.array_equal(d16.cpu().numpy(), ref)


def test_sgbm_two_stream_chunks_equal_single_stream(pkg, gpu_ctx):
    """device batches alternate chunks between two streams on disjoint scratch slots; results do not change"""
    import torch
    n, h, w = 6, 40, 260
    L = np.stack([_crop(pkg, s, h, w, y0=30 * s, x0=200 + 50 * s)[0] for s in range(n)])
    R = np.stack([_crop(pkg, s, h, w, y0=30 * s, x0=200 + 50 * s)[1] for s in range(n)])
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    got = {}
    try:
        for on in (True, False):
            gpu_ctx.set_concurrency(on)
            d16 = torch.empty((n, h, w), dtype=torch.int16, device="cuda")
            torch.cuda.synchronize()
            gpu_ctx.sgbm_compute_dev(dl, dr, n, w, h, w, w * h, d16)
            gpu_ctx.synchronize()
            got[on] = d16.cpu().numpy()
    finally:
        gpu_ctx.set_concurrency(True)
    assert np.array_equal(got[True], got[False])
    assert np.array_equal(got[True], np.stack([G.sgbm_compute(L[i], R[i]) for i in range(n)]))


def test_sgbm_random_parameters_and_sizes(pkg, gpu_ctx):
    """the seeded parameter / size sweep of tests/test_oracle_sgbm.py through the C-ABI, against live cv2"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2024)
    for trial in range(14):
        h = int(rng.integers(3, 48))
        w = int(rng.integers(104, 260))
        left, right = _crop(pkg, int(rng.integers(0, 10)), h, w, y0=int(rng.integers(0, 300)), x0=int(rng.integers(0, 900)))
        if trial % 4 == 3:
            right = np.ascontiguousarray(np.roll(right, int(rng.integers(1, 7)), axis=0))
        P1 = int(rng.integers(1, 900))
        kw = dict(P1=P1, P2=P1 + int(rng.integers(1, 2500)), disp12MaxDiff=int(rng.integers(1, 5)),
                  preFilterCap=int(rng.integers(1, 64)), uniquenessRatio=int(rng.integers(0, 40)),
                  speckleWindowSize=int(rng.choice([0, 10, 50, 100, 200])), speckleRange=int(rng.integers(1, 40)))
        ref = cv2.StereoSGBM_create(minDisparity=0, numDisparities=96, blockSize=9, **kw).compute(left, right)
        p = gpu_ctx.sgbm_params(P1=kw["P1"], P2=kw["P2"], disp12_max_diff=kw["disp12MaxDiff"], pre_filter_cap=kw["preFilterCap"],
                                uniqueness_ratio=kw["uniquenessRatio"], speckle_window_size=kw["speckleWindowSize"],
                                speckle_range=kw["speckleRange"])
        assert np.array_equal(gpu_ctx.sgbm_compute(left, right, p), ref), (trial, h, w, kw)
