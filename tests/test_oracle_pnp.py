"""CPU: oracle/pnp_oracle.c (restatement of cv::solvePnPRansac, visual_odometry.cpp:277) pinned against live cv2 4.13.0,
and the device EPnP source compiled for the host against both."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pnp_oracle as P
from pnp_scenes import CASES, garbage, scene

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jacobi_svd_bit_exact_vs_cv2():
    """OpenCV's own one-sided Jacobi (with its private hypot) below the LAPACK hand-over size: every bit of w, U, Vt --
    including the null-space columns of U on a rank-deficient M^T M, which rounding alone decides."""
    rng = np.random.default_rng(0)
    mats = [rng.normal(size=(6, 4)) for _ in range(30)] + [rng.normal(size=(6, 5)) for _ in range(10)]
    mats += [rng.normal(size=(3, 3)) for _ in range(30)]
    for _ in range(40):
        M = rng.normal(size=(10, 12))
        mats.append(M.T @ M)
    mats.append(np.zeros((3, 3)))      # zero singular values: OpenCV's random-vector completion of U
    mats.append(np.diag([2.0, 0.0, 0.0]))
    for A in mats:
        w, u, vt = cv2.SVDecomp(A)
        w2, u2, vt2 = P.svd(A)
        assert np.array_equal(w.ravel(), w2) and np.array_equal(u, u2) and np.array_equal(vt, vt2)


def test_rodrigues_and_errors_bit_exact_vs_cv2(pkg):
    rng = np.random.default_rng(5)
    K = pkg.synth.kitti_K()
    n = 5000
    pw = np.stack([rng.uniform(-15, 15, n), rng.uniform(-3, 3, n), rng.uniform(6, 45, n)], 1).astype(np.float32)
    uv = rng.uniform(0, 1200, (n, 2)).astype(np.float32)
    for _ in range(10):
        rvec, tvec = rng.normal(0, 0.3, 3), rng.normal(0, 1, 3)
        proj = cv2.projectPoints(pw, rvec, tvec, K, None)[0].reshape(-1, 2)
        d = uv - proj
        assert np.array_equal((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32), P.errors(pw, uv, K, rvec, tvec))
        Rm = cv2.Rodrigues(rvec)[0]
        assert np.array_equal(Rm, P.rodrigues_to_mat(rvec))
        assert np.array_equal(cv2.Rodrigues(Rm)[0].ravel(), P.rodrigues_to_vec(Rm))


def test_epnp_minimal_solver_bit_exact_vs_cv2(pkg):
    pw, uv, K, *_ = scene(pkg, 1, 400, 0.2)
    rng = np.random.default_rng(2)
    for _ in range(200):
        idx = rng.choice(len(pw), 5, replace=False)
        ok, rv, tv = cv2.solvePnP(pw[idx], uv[idx], K, None, flags=cv2.SOLVEPNP_EPNP)
        r2, t2, _ = P.epnp(pw[idx], uv[idx], K)
        assert np.array_equal(rv.ravel(), r2) and np.array_equal(tv.ravel(), t2)


@pytest.mark.parametrize("case", CASES)
def test_ransac_inliers_index_exact_vs_cv2(pkg, case):
    for rep in range(3):
        seed, n, outl, noise = case
        pw, uv, K, *_ = scene(pkg, seed + 100 * rep, n, outl, noise)
        ok, rvec, tvec, inl = cv2.solvePnPRansac(pw, uv, K, None, iterationsCount=100, reprojectionError=4.0, confidence=0.99)
        o = P.solve_pnp_ransac(pw, uv, K)
        assert ok == o["ok"]
        if not ok:
            continue
        assert np.array_equal(inl.ravel(), o["inliers"])
        assert np.abs(rvec.ravel() - o["rvec"]).max() < 1e-6
        assert np.abs(tvec.ravel() - o["tvec"]).max() / np.linalg.norm(tvec) < 1e-6


def test_ransac_other_parameters_and_garbage(pkg):
    pw, uv, K, *_ = scene(pkg, 21, 300, 0.35, 0.6)
    for iters, thr, conf in [(20, 2.0, 0.9), (300, 8.0, 0.999), (1, 4.0, 0.99), (100, 1.0, 0.5)]:
        ok, rvec, tvec, inl = cv2.solvePnPRansac(pw, uv, K, None, iterationsCount=iters, reprojectionError=thr, confidence=conf)
        o = P.solve_pnp_ransac(pw, uv, K, iters, thr, conf)
        assert ok == o["ok"]
        if ok:
            assert np.array_equal(inl.ravel(), o["inliers"])
    for s in range(5):
        gx, gu = garbage(s)
        ok, rvec, tvec, inl = cv2.solvePnPRansac(gx, gu, K, None, iterationsCount=100, reprojectionError=4.0, confidence=0.99)
        o = P.solve_pnp_ransac(gx, gu, K)
        assert ok == o["ok"]
        if ok:
            assert np.array_equal(inl.ravel(), o["inliers"])


@pytest.fixture(scope="module")
def epnp_host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("epnp") / "epnp_host.so")
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                    os.path.join(ROOT, "tests", "native", "epnp_host.cpp")], check=True)
    return C.CDLL(so)


def test_device_epnp_source_on_host_equals_oracle(pkg, epnp_host_lib):
    """csrc/epnp.cuh compiled for the CPU (qualifiers defined away): the device port itself is bit-identical to the
    oracle and therefore to cv2 -- what remains GPU-specific is IEEE fp64 add/mul/div/sqrt, checked in test_pnp_gpu."""
    pw, uv, K, *_ = scene(pkg, 4, 500, 0.5, 1.0)
    Kc = np.ascontiguousarray(K).reshape(9)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(3)
    for _ in range(200):
        idx = rng.choice(len(pw), 5, replace=False).astype(np.int32)
        R, t, rv, R2 = np.zeros(9), np.zeros(3), np.zeros(3), np.zeros(9)
        epnp_host_lib.epnp_host(p(pw), p(uv), p(idx), p(Kc), p(R), p(t), p(rv), p(R2))
        r_o, t_o, R_o = P.epnp(pw[idx], uv[idx], K)
        assert np.array_equal(rv, r_o) and np.array_equal(t, t_o) and np.array_equal(R.reshape(3, 3), R_o)
        assert np.array_equal(R2.reshape(3, 3), P.rodrigues_to_mat(r_o))
