"""ORACLE (test infrastructure only): ctypes loader for oracle/ba_oracle.c (see its header for provenance)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "_build", "libba_oracle.so")
_lib = None


class Options(C.Structure):
    _fields_ = [("huber_delta", C.c_double), ("chi2_th", C.c_double), ("num_iterations", C.c_int),
                ("pose_only", C.c_int), ("max_trials", C.c_int), ("tau", C.c_double)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int), ("trials", C.c_int), ("accepted", C.c_int), ("chi2_initial", C.c_double),
                ("chi2_final", C.c_double), ("lambda_final", C.c_double), ("chi2_threshold", C.c_double),
                ("n_inlier_obs", C.c_int), ("n_outlier_obs", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "ba_oracle.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(_LIB)
        _lib.ba_oracle_chi2.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def optimize(poses, points, obs_pose, obs_point, obs_uv, K, num_iterations=10, pose_only=False, huber_delta=5.991,
             chi2_th=5.991, max_trials=10, tau=1e-5, trace_cap=256):
    """optimize_map / optimize_pose_only restatement.  Returns a dict (inputs are not modified)."""
    poses = np.array(poses, dtype=np.float64, order="C").reshape(-1, 12).copy()
    points = np.array(points, dtype=np.float64, order="C").reshape(-1, 3).copy()
    op = np.ascontiguousarray(obs_pose, dtype=np.int32)
    ol = np.ascontiguousarray(obs_point, dtype=np.int32)
    uv = np.ascontiguousarray(obs_uv, dtype=np.float64).reshape(-1, 2)
    Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    opt = Options(huber_delta, chi2_th, num_iterations, int(pose_only), max_trials, tau)
    res = Result()
    chi2 = np.zeros(max(len(op), 1), dtype=np.float64)
    inl = np.ones(max(len(points), 1), dtype=np.uint8)
    trace = np.zeros((trace_cap, 4), dtype=np.float64)
    lib().ba_oracle_optimize(len(poses), _p(poses), len(points), _p(points), len(op), _p(op), _p(ol), _p(uv), _p(Kc),
                             C.byref(opt), C.byref(res), _p(chi2), _p(inl), _p(trace), trace_cap)
    return dict(poses=poses, points=points, chi2_per_obs=chi2[:len(op)], point_inlier=inl[:len(points)].astype(bool),
                iterations=res.iterations, trials=res.trials, accepted=res.accepted, chi2_initial=res.chi2_initial,
                chi2_final=res.chi2_final, lambda_final=res.lambda_final, chi2_threshold=res.chi2_threshold,
                n_inlier_obs=res.n_inlier_obs, n_outlier_obs=res.n_outlier_obs, trace=trace[:res.trials])


def chi2(poses, points, obs_pose, obs_point, obs_uv, K, huber_delta=5.991):
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 12)
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    op = np.ascontiguousarray(obs_pose, dtype=np.int32)
    ol = np.ascontiguousarray(obs_point, dtype=np.int32)
    uv = np.ascontiguousarray(obs_uv, dtype=np.float64).reshape(-1, 2)
    Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    return float(lib().ba_oracle_chi2(len(poses), _p(poses), len(points), _p(points), len(op), _p(op), _p(ol), _p(uv),
                                      _p(Kc), C.c_double(huber_delta)))


def dense_system(poses, points, obs_pose, obs_point, obs_uv, K, huber_delta=5.991):
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 12)
    points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    op = np.ascontiguousarray(obs_pose, dtype=np.int32)
    ol = np.ascontiguousarray(obs_point, dtype=np.int32)
    uv = np.ascontiguousarray(obs_uv, dtype=np.float64).reshape(-1, 2)
    Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    N = 6 * len(poses) + 3 * len(points)
    H = np.zeros((N, N)); b = np.zeros(N)
    lib().ba_oracle_dense_system(len(poses), _p(poses), len(points), _p(points), len(op), _p(op), _p(ol), _p(uv),
                                 _p(Kc), C.c_double(huber_delta), _p(H), _p(b))
    return H, b


def se3_exp(xi):
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    R = np.zeros(9); t = np.zeros(3)
    lib().ba_se3_exp(_p(xi), _p(R), _p(t))
    return R.reshape(3, 3), t


def residual_and_jacobians(T, p, K, z, pose_only=False):
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(12); p = np.ascontiguousarray(p, dtype=np.float64)
    Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9); z = np.ascontiguousarray(z, dtype=np.float64)
    e = np.zeros(2); pc = np.zeros(3); A = np.zeros(12); B = np.zeros(6)
    lib().ba_residual(_p(T), _p(p), _p(Kc), _p(z), _p(e), _p(pc))
    lib().ba_jacobians(_p(T), _p(pc), _p(Kc), int(pose_only), _p(A), _p(B))
    return e, A.reshape(2, 6), B.reshape(2, 3)
