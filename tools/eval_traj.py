"""Trajectory evaluation for the pose file the reference writes (map.cpp:168-204: one row per keyframe,
`frame_id r00 r01 r02 x r10 r11 r12 y r20 r21 r22 z` of T_w_c) against KITTI odometry ground truth
(`poses/XX.txt`: 12 numbers per frame, same row-major 3x4 layout without the id).

    python tools/eval_traj.py estimated_traj.txt poses/00.txt [--align]

Reports ATE-RMSE over the frames present in the estimate (optionally after a rigid Umeyama alignment) and the KITTI
relative errors (translation %, rotation deg/m) over all sub-sequences of 100..800 m that start and end on an
estimated frame.  KITTI PNGs are converted for run_vslam with `python tools/eval_traj.py --kitti-to-pgm <seq_dir> <out_dir>`.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np


def load_estimate(path: str):
    """-> (frame ids [n], T_w_c [n,3,4]); duplicate ids keep the last row (a keyframe can be written twice)."""
    a = np.loadtxt(path, ndmin=2)
    if a.shape[1] != 13:
        raise ValueError(f"{path}: expected 13 columns (frame id + 12), got {a.shape[1]}")
    ids = a[:, 0].astype(np.int64)
    last = {int(i): k for k, i in enumerate(ids)}
    keep = np.array(sorted(last.values()))
    order = np.argsort(ids[keep], kind="stable")
    keep = keep[order]
    return ids[keep], a[keep, 1:].reshape(-1, 3, 4)


def load_kitti_poses(path: str):
    a = np.loadtxt(path, ndmin=2)
    if a.shape[1] != 12:
        raise ValueError(f"{path}: expected 12 columns, got {a.shape[1]}")
    return a.reshape(-1, 3, 4)


def umeyama_rigid(src: np.ndarray, dst: np.ndarray):
    """Least-squares R, t with dst ~ R src + t (no scale: stereo is metric)."""
    mu_s, mu_d = src.mean(0), dst.mean(0)
    H = (dst - mu_d).T @ (src - mu_s)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt))])
    R = U @ D @ Vt
    return R, mu_d - R @ mu_s


def ate_rmse(est_t: np.ndarray, gt_t: np.ndarray, align: bool = False) -> float:
    if align and len(est_t) >= 3:
        R, t = umeyama_rigid(est_t, gt_t)
        est_t = est_t @ R.T + t
    return float(np.sqrt(((est_t - gt_t) ** 2).sum(1).mean()))


def _to44(T):
    M = np.tile(np.eye(4), (len(T), 1, 1))
    M[:, :3, :] = T
    return M


def kitti_relative_errors(ids, est, gt, lengths=(100, 200, 300, 400, 500, 600, 700, 800)):
    """KITTI devkit metric restricted to the estimated frames: for every start frame and every length, the first
    estimated frame at least that far along the ground-truth path; error of the relative motion."""
    gt44, est44 = _to44(gt[ids]), _to44(est)
    step = np.linalg.norm(np.diff(gt[:, :, 3], axis=0), axis=1)
    dist_all = np.concatenate([[0.0], np.cumsum(step)])
    dist = dist_all[ids]
    t_errs, r_errs = [], []
    for i in range(len(ids)):
        for L in lengths:
            j = np.searchsorted(dist, dist[i] + L)
            if j >= len(ids):
                continue
            d_gt = np.linalg.inv(gt44[i]) @ gt44[j]
            d_est = np.linalg.inv(est44[i]) @ est44[j]
            e = np.linalg.inv(d_est) @ d_gt
            length = dist[j] - dist[i]
            t_errs.append(np.linalg.norm(e[:3, 3]) / length)
            r_errs.append(np.arccos(np.clip((np.trace(e[:3, :3]) - 1) / 2, -1, 1)) / length)
    if not t_errs:
        return None, None
    return float(np.mean(t_errs) * 100.0), float(np.degrees(np.mean(r_errs)))


def evaluate(est_path: str, gt_path: str, align: bool = False) -> dict:
    ids, est = load_estimate(est_path)
    gt = load_kitti_poses(gt_path)
    ok = ids < len(gt)
    ids, est = ids[ok], est[ok]
    t_pct, r_deg_m = kitti_relative_errors(ids, est, gt)
    return {"frames": int(len(ids)), "path_length_m": float(np.linalg.norm(np.diff(gt[:ids.max() + 1, :, 3], axis=0), axis=1).sum()),
            "ate_rmse_m": ate_rmse(est[:, :, 3], gt[ids][:, :, 3], align), "aligned": bool(align),
            "kitti_t_err_percent": t_pct, "kitti_r_err_deg_per_m": r_deg_m}


def kitti_to_pgm(seq_dir: str, out_dir: str, n: int | None = None):
    """KITTI `sequences/XX/image_{0,1}/%06d.png` -> the binary PGMs VO::read_img loads (visual_odometry.cpp:37-68 reads
    the PNGs with cv::imread; the drop-in's host layer has no image codec)."""
    import cv2
    for cam in ("image_0", "image_1"):
        os.makedirs(os.path.join(out_dir, cam), exist_ok=True)
        names = sorted(f for f in os.listdir(os.path.join(seq_dir, cam)) if f.endswith(".png"))
        for f in names[:n]:
            img = cv2.imread(os.path.join(seq_dir, cam, f), cv2.IMREAD_GRAYSCALE)
            with open(os.path.join(out_dir, cam, f[:-4] + ".pgm"), "wb") as o:
                o.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
                o.write(img.tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("estimate", nargs="?")
    ap.add_argument("ground_truth", nargs="?")
    ap.add_argument("--align", action="store_true")
    ap.add_argument("--kitti-to-pgm", nargs=2, metavar=("SEQ_DIR", "OUT_DIR"))
    ap.add_argument("--frames", type=int, default=None)
    a = ap.parse_args()
    if a.kitti_to_pgm:
        kitti_to_pgm(a.kitti_to_pgm[0], a.kitti_to_pgm[1], a.frames)
        return
    if not a.estimate or not a.ground_truth:
        ap.error("estimate and ground_truth are required")
    import json
    print(json.dumps(evaluate(a.estimate, a.ground_truth, a.align)))


if __name__ == "__main__":
    main()
