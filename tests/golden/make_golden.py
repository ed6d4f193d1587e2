"""Regenerate the committed golden vectors from the LIVE reference libraries (cv2 4.13.0) and the C BA oracle.

    python tests/golden/make_golden.py

The reference repo has no tests or fixtures of its own; these vectors pin the oracle and the CUDA path
to what cv2 4.13.0 (the OpenCV calls the reference makes) produced in the build container.  Inputs are regenerated
from seeds by stereo-visual-slam_b200/synth.py, so only outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
import vslam_b200_loader  # noqa: E402

pkg = vslam_b200_loader.pkg
from oracle import ba_oracle, orb_restate as R, vo_restate as V  # noqa: E402


def cv_orb(img, n):
    kps, desc = cv2.ORB_create(n).detectAndCompute(img, None)
    o = np.array([k.octave for k in kps]); r = np.array([k.response for k in kps], np.float32)
    pt = np.array([k.pt for k in kps], np.float32); a = np.array([k.angle for k in kps], np.float32)
    sz = np.array([k.size for k in kps], np.float32)
    sc = R.level_scales()
    xl = np.rint(pt[:, 0] / sc[o]).astype(np.int32); yl = np.rint(pt[:, 1] / sc[o]).astype(np.int32)
    order = R.canonical_order(o, r, yl, xl)
    return dict(octave=o[order].astype(np.int32), response=r[order], pt=pt[order], angle=a[order], size=sz[order],
                desc=desc[order])


def sgbm_crop(s, seed, h, w, y0=100, x0=300):
    left, right, _ = s.synth_pair(seed)
    return np.ascontiguousarray(left[y0:y0 + h, x0:x0 + w]), np.ascontiguousarray(right[y0:y0 + h, x0:x0 + w])


def main():
    s = pkg.synth
    left, right, _ = s.synth_pair(0)
    g = cv_orb(left, 500)
    np.savez_compressed(os.path.join(HERE, "orb_seed0_n500.npz"), **g)
    small = s.synth_canvas(5, 400, 240)
    g = cv_orb(small, 300)
    np.savez_compressed(os.path.join(HERE, "orb_canvas5_400x240_n300.npz"), **g)
    # matching: cv2 BFMatcher cross-check on seeded descriptors with ties
    q = s.synth_descriptors(10, 400, dup_frac=0.3)
    t = np.concatenate([q[::2], s.synth_descriptors(11, 300, dup_frac=0.3)])
    m = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
    np.savez_compressed(os.path.join(HERE, "match_ties_400x500.npz"),
                        queryIdx=np.array([x.queryIdx for x in m], np.int32),
                        trainIdx=np.array([x.trainIdx for x in m], np.int32),
                        distance=np.array([x.distance for x in m], np.float32))
    # triangulation: cv2.triangulatePoints
    rng = np.random.default_rng(0)
    P1, P2 = V.stereo_projection_matrices(s.FX, s.FY, s.CX, s.CY, s.BASELINE_M)
    xl = np.stack([rng.uniform(100, 1100, 64), rng.uniform(40, 330, 64)], 1).astype(np.float32)
    xr = xl.copy(); xr[:, 0] -= rng.uniform(1.5, 60, 64).astype(np.float32); xr[:, 1] += rng.normal(0, 0.4, 64).astype(np.float32)
    X = cv2.triangulatePoints(P1, P2, xl.T.astype(np.float64), xr.T.astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "triangulate_64.npz"), xl=xl, xr=xr, xyz=(X[:3] / X[3]).T)
    # BA: the C oracle on a small window (pins the restatement against regressions; the reference pins nothing)
    p = s.synth_ba_problem(11, 5, 120, outlier_frac=0.05)
    r = ba_oracle.optimize(p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"], num_iterations=10)
    np.savez_compressed(os.path.join(HERE, "ba_seed11_K5_L120_it10.npz"), poses=r["poses"], points=r["points"],
                        chi2=np.array([r["chi2_initial"], r["chi2_final"]]), trials=r["trials"],
                        lambda_final=r["lambda_final"], point_inlier=r["point_inlier"], trace=r["trace"])
    # dense stereo: cv2.StereoSGBM with the reference's constants (visual_odometry.cpp:163-164) on crops of pair 0
    sg = cv2.StereoSGBM_create(0, 96, 9, 8 * 81, 32 * 81, 1, 63, 10, 100, 32)
    crop = sgbm_crop(s, 0, 64, 360)
    np.savez_compressed(os.path.join(HERE, "sgbm_pair0_crop64x360.npz"), disp16=sg.compute(*crop))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
