"""CPU: the trajectory evaluation tool on the reference's pose-file format (map.cpp:168-204)."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("eval_traj", os.path.join(ROOT, "tools", "eval_traj.py"))
E = importlib.util.module_from_spec(spec)
spec.loader.exec_module(E)


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _gt(n=1200):
    T = np.zeros((n, 3, 4))
    p = np.zeros(3)
    for i in range(n):
        R = _rot_y(0.002 * i)
        T[i, :, :3] = R
        T[i, :, 3] = p
        p = p + R @ np.array([0.0, 0.0, 0.8])
    return T


def _write_est(path, ids, T):
    with open(path, "w") as f:
        for i, t in zip(ids, T):
            f.write(str(int(i)) + " " + " ".join(repr(float(v)) for v in t.reshape(-1)) + "\n")


def test_perfect_estimate_has_zero_error(tmp_path):
    gt = _gt()
    np.savetxt(tmp_path / "gt.txt", gt.reshape(len(gt), 12))
    ids = np.arange(0, len(gt), 7)
    _write_est(tmp_path / "est.txt", ids, gt[ids])
    r = E.evaluate(str(tmp_path / "est.txt"), str(tmp_path / "gt.txt"))
    assert r["frames"] == len(ids) and r["ate_rmse_m"] < 1e-9
    assert r["kitti_t_err_percent"] < 1e-6 and r["kitti_r_err_deg_per_m"] < 1e-6


def test_scale_drift_shows_up_as_translation_percent(tmp_path):
    gt = _gt()
    np.savetxt(tmp_path / "gt.txt", gt.reshape(len(gt), 12))
    est = gt.copy()
    est[:, :, 3] *= 1.02  # 2 % scale error
    ids = np.arange(len(gt))
    _write_est(tmp_path / "est.txt", ids, est)
    r = E.evaluate(str(tmp_path / "est.txt"), str(tmp_path / "gt.txt"))
    assert 1.5 < r["kitti_t_err_percent"] < 2.5
    assert r["ate_rmse_m"] > 1.0


def test_alignment_removes_a_rigid_offset_and_duplicates_keep_last(tmp_path):
    gt = _gt(400)
    np.savetxt(tmp_path / "gt.txt", gt.reshape(len(gt), 12))
    Rz = np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    est = gt.copy()
    est[:, :, :3] = Rz @ gt[:, :, :3]
    est[:, :, 3] = gt[:, :, 3] @ Rz.T + np.array([5.0, -2.0, 1.0])
    ids = np.arange(0, 400, 3)
    rows_ids = np.concatenate([[ids[4]], ids])  # a stale duplicate of frame ids[4] first
    rows_T = np.concatenate([np.zeros((1, 3, 4)), est[ids]])
    _write_est(tmp_path / "est.txt", rows_ids, rows_T)
    assert E.evaluate(str(tmp_path / "est.txt"), str(tmp_path / "gt.txt"))["ate_rmse_m"] > 1.0
    r = E.evaluate(str(tmp_path / "est.txt"), str(tmp_path / "gt.txt"), align=True)
    assert r["frames"] == len(ids) and r["ate_rmse_m"] < 1e-6
