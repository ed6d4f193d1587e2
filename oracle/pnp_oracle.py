"""ORACLE (test infrastructure only): ctypes loader for oracle/pnp_oracle.c (see its header for provenance).

solve_pnp_ransac() restates cv::solvePnPRansac(obj f32, img f32, K, noDist, rvec, tvec, false, iters, thr, conf,
inliers) as VO::motion_estimation calls it (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:277).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "_build", "libpnp_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "pnp_oracle.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def svd(A):
    """cv::SVD::compute through OpenCV's Jacobi path: returns (w, U, Vt) like cv2.SVDecomp (m >= n)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    m, n = A.shape
    w = np.zeros(n); Ut = np.zeros((n, m)); Vt = np.zeros((n, n))
    lib().pnp_oracle_svd(_p(A), m, n, _p(w), _p(Ut), _p(Vt))
    return w, Ut.T.copy(), Vt


def epnp(obj, img, K, debug=False):
    """cv2.solvePnP(obj, img, K, None, flags=cv2.SOLVEPNP_EPNP) -> (rvec, tvec, R[, Ut, d])"""
    obj = np.ascontiguousarray(obj, dtype=np.float32).reshape(-1, 3)
    img = np.ascontiguousarray(img, dtype=np.float32).reshape(-1, 2)
    Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    assert len(obj) == len(img) and 4 <= len(obj) <= 16
    rvec = np.zeros(3); tvec = np.zeros(3); R = np.zeros(9); dbg = np.zeros(156)
    lib().pnp_oracle_epnp(_p(obj), _p(img), len(obj), _p(Kc), _p(rvec), _p(tvec), _p(R), _p(dbg))
    if debug:
        return rvec, tvec, R.reshape(3, 3), dbg[:144].reshape(12, 12), dbg[144:]
    return rvec, tvec, R.reshape(3, 3)


def errors(obj, img, K, rvec, tvec):
    """PnPRansacCallback::computeError: float32 squared reprojection distances."""
    obj = np.ascontiguousarray(obj, dtype=np.float32).reshape(-1, 3)
    img = np.ascontiguousarray(img, dtype=np.float32).reshape(-1, 2)
    Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    rv = np.ascontiguousarray(rvec, dtype=np.float64).reshape(3); tv = np.ascontiguousarray(tvec, dtype=np.float64).reshape(3)
    err = np.zeros(len(obj), dtype=np.float32)
    lib().pnp_oracle_errors(_p(obj), _p(img), len(obj), _p(Kc), _p(rv), _p(tv), _p(err))
    return err


def rodrigues_to_vec(R):
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(9); r = np.zeros(3)
    lib().pnp_oracle_rodrigues_to_vec(_p(R), _p(r))
    return r


def rodrigues_to_mat(r):
    r = np.ascontiguousarray(r, dtype=np.float64).reshape(3); R = np.zeros(9)
    lib().pnp_oracle_rodrigues_to_mat(_p(r), _p(R))
    return R.reshape(3, 3)


def draw_samples(n, iters):
    """The sample stream of RANSACPointSetRegistrator::run: cv::RNG((uint64)-1), 5 distinct indices per iteration."""
    state = (1 << 64) - 1
    out = np.zeros((iters, 5), dtype=np.int32)
    for it in range(iters):
        for i in range(5):
            while True:
                state = ((state & 0xFFFFFFFF) * 4164903690 + (state >> 32)) & ((1 << 64) - 1)
                c = (state & 0xFFFFFFFF) % n
                if c not in out[it, :i]:
                    out[it, i] = c
                    break
    return out


def solve_pnp_ransac(obj, img, K, iters=100, reproj_err=4.0, confidence=0.99, refit=True, trace_cap=512):
    """Returns dict(ok, rvec, tvec, T_c_w (3x4), inliers (ascending int32), n_iters, trace, ransac_rvec, ransac_tvec)."""
    obj = np.ascontiguousarray(obj, dtype=np.float32).reshape(-1, 3)
    img = np.ascontiguousarray(img, dtype=np.float32).reshape(-1, 2)
    Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    n = len(obj)
    if n == 5:   # solvePnPRansac: model_points == npoints -> solvePnP(EPNP) on all points, every point an inlier
        rvec, tvec, _ = epnp(obj, img, K)
        T = np.concatenate([rodrigues_to_mat(rvec), tvec.reshape(3, 1)], axis=1)
        return dict(ok=True, n_iters=0, trace=np.zeros((0, 8), np.int32), ransac_rvec=rvec, ransac_tvec=tvec,
                    inliers=np.arange(5, dtype=np.int32), rvec=rvec, tvec=tvec, T_c_w=T)
    mask = np.zeros(max(n, 1), dtype=np.uint8)
    rvec = np.zeros(3); tvec = np.zeros(3); T = np.zeros(12)
    n_good = C.c_int(0)
    trace = np.zeros((trace_cap, 8), dtype=np.int32)
    n_it = lib().pnp_oracle_ransac(_p(obj), _p(img), n, _p(Kc), int(iters), C.c_double(reproj_err), C.c_double(confidence),
                                   _p(mask), _p(rvec), _p(tvec), C.byref(n_good), _p(trace), trace_cap)
    ok = n_good.value > 0
    out = dict(ok=ok, n_iters=n_it, trace=trace[:min(n_it, trace_cap)], ransac_rvec=rvec.copy(), ransac_tvec=tvec.copy(),
               inliers=np.flatnonzero(mask[:n]).astype(np.int32) if ok else np.zeros(0, np.int32))
    if ok and refit:
        lib().pnp_oracle_refit(_p(obj), _p(img), n, _p(mask), _p(Kc), _p(rvec), _p(tvec), _p(T))
        out["T_c_w"] = T.reshape(3, 4).copy()
    out["rvec"], out["tvec"] = rvec, tvec
    return out
