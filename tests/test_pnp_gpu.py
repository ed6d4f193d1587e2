"""GPU parity: K12 PnP-RANSAC vs live cv2.solvePnPRansac -- inlier lists index for index on any input, pose within 1e-4
(north star; measured ~1e-9) -- and per RANSAC sample against the oracle; K7 stand-alone ANMS vs the oracle."""
import numpy as np
import pytest

from oracle import pnp_oracle as P
from oracle import vo_restate as V
from pnp_scenes import CASES, garbage, scene

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4   # north star: poses within 1e-4 relative


def _check_vs_cv2(g, ok, rvec, tvec, inl):
    import cv2
    if not ok:
        assert len(g["inliers"]) == 0
        return
    assert np.array_equal(g["inliers"], inl.ravel())            # index for index, no tolerance
    Rc, _ = cv2.Rodrigues(rvec)
    assert np.abs(g["T_c_w"][:, :3] - Rc).max() < REL_TOL
    assert np.abs(g["tvec"] - tvec.ravel()).max() / np.linalg.norm(tvec) < REL_TOL
    assert np.abs(g["rvec"] - rvec.ravel()).max() < REL_TOL
    assert np.allclose(g["T_c_w"][:, :3] @ g["T_c_w"][:, :3].T, np.eye(3), atol=1e-9)


@pytest.mark.parametrize("case", CASES)
def test_pnp_inliers_index_exact_vs_cv2(pkg, gpu_ctx, case):
    """noisy scenes (0.3 .. 2 px), 0 .. 60 % gross outliers, n = 6 .. 2000: same inlier list as cv2, same pose"""
    import cv2
    seed, n, outl, noise = case
    for rep in range(3):
        pw, uv, K, R, t, good = scene(pkg, seed + 100 * rep, n, outl, noise)
        ok, rvec, tvec, inl = cv2.solvePnPRansac(pw, uv, K, None, iterationsCount=100, reprojectionError=4.0, confidence=0.99)
        g = gpu_ctx.pnp_ransac(pw, uv, K, 100, 4.0, 0.99)
        _check_vs_cv2(g, ok, rvec, tvec, inl)
        if ok and n >= 60 and outl <= 0.3:   # and against ground truth
            assert np.abs(g["T_c_w"][:, :3] - R).max() < 5e-3 and np.abs(g["tvec"] - t).max() < 0.1


def test_pnp_per_sample_models_and_counts_vs_oracle(pkg, gpu_ctx):
    """every RANSAC sample: the EPnP model (OpenCV's arithmetic, rounding-determined null-space basis included) and its
    inlier count equal the oracle's; the loop stops after the same number of iterations"""
    pw, uv, K, *_ = scene(pkg, 4, 500, 0.5, 1.0)
    gpu_ctx.pnp_ransac(pw, uv, K, 100, 4.0, 0.99)
    models, counts, executed = gpu_ctx.pnp_debug(100)
    o = P.solve_pnp_ransac(pw, uv, K)
    assert executed == o["n_iters"]
    worst = 0.0
    for it in range(100):   # the kernel evaluates all 100 drawn samples, OpenCV's loop only the first `executed`
        idx = P.draw_samples(len(pw), 100)[it]
        rv, tv, _ = P.epnp(pw[idx], uv[idx], K)
        Rm = P.rodrigues_to_mat(rv)
        worst = max(worst, np.abs(models[it, :9].reshape(3, 3) - Rm).max(), np.abs(models[it, 9:] - tv).max() / max(1.0, np.abs(tv).max()))
        assert counts[it] == int((P.errors(pw, uv, K, rv, tv) <= np.float32(16.0)).sum())
        if it < executed:
            assert np.array_equal(idx, o["trace"][it, :5]) and counts[it] == o["trace"][it, 5]
    # acos / sin / cos of the model's Rodrigues round trip are the only non-IEEE-exact operations on the device
    assert worst < 1e-13, worst


def test_pnp_other_parameters_and_garbage(pkg, gpu_ctx):
    import cv2
    pw, uv, K, *_ = scene(pkg, 21, 300, 0.35, 0.6)
    for iters, thr, conf in [(20, 2.0, 0.9), (300, 8.0, 0.999), (1, 4.0, 0.99), (100, 1.0, 0.5)]:
        ok, rvec, tvec, inl = cv2.solvePnPRansac(pw, uv, K, None, iterationsCount=iters, reprojectionError=thr, confidence=conf)
        _check_vs_cv2(gpu_ctx.pnp_ransac(pw, uv, K, iters, thr, conf), ok, rvec, tvec, inl)
    for s in range(5):      # pure garbage: whatever cv2 decides, index for index
        gx, gu = garbage(s)
        ok, rvec, tvec, inl = cv2.solvePnPRansac(gx, gu, K, None, iterationsCount=100, reprojectionError=4.0, confidence=0.99)
        g = gpu_ctx.pnp_ransac(gx, gu, K)
        assert np.array_equal(g["inliers"], inl.ravel() if ok else np.zeros(0, np.int32))


def test_pnp_small_and_degenerate_inputs(pkg, gpu_ctx):
    import cv2
    K = pkg.synth.kitti_K()
    g = gpu_ctx.pnp_ransac(np.zeros((3, 3), np.float32), np.zeros((3, 2), np.float32), K)
    assert len(g["inliers"]) == 0
    pw, uv, *_ = scene(pkg, 30, 5, 0.0, 0.1)   # n == 5: OpenCV returns the EPnP pose itself, all five as inliers
    ok, rvec, tvec, inl = cv2.solvePnPRansac(pw, uv, K, None, iterationsCount=100, reprojectionError=4.0, confidence=0.99)
    g = gpu_ctx.pnp_ransac(pw, uv, K)
    assert ok and np.array_equal(g["inliers"], inl.ravel()) and len(inl) == 5
    assert np.abs(g["rvec"] - rvec.ravel()).max() < 1e-9 and np.abs(g["tvec"] - tvec.ravel()).max() < 1e-9
    same = np.tile(pw[:1], (20, 1)), np.tile(uv[:1], (20, 1))   # all points identical: rank-0 control-point PCA
    ok, rvec, tvec, inl = cv2.solvePnPRansac(same[0], same[1], K, None, iterationsCount=100, reprojectionError=4.0, confidence=0.99)
    g = gpu_ctx.pnp_ransac(same[0], same[1], K)
    assert np.array_equal(g["inliers"], inl.ravel() if ok else np.zeros(0, np.int32))


def test_anms_standalone_vs_oracle(pkg, gpu_ctx):
    rng = np.random.default_rng(4)
    n = 2500
    kp = np.zeros(n, dtype=pkg.KEYPOINT_DTYPE)
    kp["x"] = rng.uniform(0, 1241, n).astype(np.float32)
    kp["y"] = rng.uniform(0, 376, n).astype(np.float32)
    kp["response"] = rng.uniform(1e-6, 1e-2, n).astype(np.float32)
    kp["response"][::9] = kp["response"][4]
    keep = gpu_ctx.anms(kp, 500)
    ref = V.anms(np.stack([kp["x"], kp["y"]], 1), kp["response"], 500)
    assert np.array_equal(np.sort(keep), np.sort(ref)) and len(keep) >= 500
    assert np.array_equal(gpu_ctx.anms(kp[:100], 500), np.arange(100))
