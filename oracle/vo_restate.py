"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the reference's own VO stage code
(/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp) and of the OpenCV calls it makes:

* `anms`                 -- VO::adaptive_non_maximal_suppresion, visual_odometry.cpp:96-157
* `bf_match_crosscheck`  -- cv::BFMatcher(NORM_HAMMING, crossCheck=true)::match, visual_odometry.cpp:24,225
* `match_gate`           -- distance gate of VO::feature_matching, visual_odometry.cpp:228-246
* `triangulate_dlt`      -- the role of disparity_map + Frame::find_3d + set_ref_3d_position
                            (visual_odometry.cpp:159-217, types_def.cpp:9-18) in the north star's sparse form;
                            arithmetic oracle = cv2.triangulatePoints (SVD of the 4x4 DLT system)
* `feature_detection`    -- visual_odometry.cpp:70-94 (detect -> ANMS -> compute), canonical order

Pinned by tests/test_oracle_vo.py against live cv2 4.13.0 (BFMatcher, triangulatePoints) and the
committed fixtures in tests/golden/.  The reference itself has no tests.
"""
from __future__ import annotations

import numpy as np

from . import orb_restate as R

_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.uint16)


def hamming_matrix(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """D[i][j] = popcount(q[i] ^ t[j]) over 32 bytes (core/src/batch_distance.cpp, NORM_HAMMING)."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    D = np.zeros((len(q), len(t)), dtype=np.uint16)
    for b in range(q.shape[1]):
        D += _POP8[q[:, b][:, None] ^ t[:, b][None, :]]
    return D


def bf_match_crosscheck(q: np.ndarray, t: np.ndarray):
    """Strict mutual nearest neighbour, first-minimum tie-breaking, ascending queryIdx.

    Returns (queryIdx, trainIdx, distance) int32 arrays."""
    if len(q) == 0 or len(t) == 0:
        z = np.zeros(0, dtype=np.int32)
        return z, z.copy(), z.copy()
    D = hamming_matrix(q, t)
    fw = D.argmin(axis=1)  # lowest j on ties
    bw = D.argmin(axis=0)  # lowest i on ties
    qi = np.arange(len(q))
    keep = bw[fw] == qi
    qi = qi[keep]
    ti = fw[keep]
    return qi.astype(np.int32), ti.astype(np.int32), D[qi, ti].astype(np.int32)


def match_gate(qi, ti, dist, frame_gap: float):
    """keep distance <= max(2*min_dist, 30*frame_gap) (visual_odometry.cpp:239-246).

    The reference dereferences minmax_element of an empty vector when there is no match (UB,
    visual_odometry.cpp:229-242); the restatement returns the empty set."""
    if len(dist) == 0:
        return qi, ti, dist
    thr = max(2.0 * float(dist.min()), 30.0 * float(frame_gap))
    k = dist.astype(np.float64) <= thr
    return qi[k], ti[k], dist[k]


def anms(pt: np.ndarray, response: np.ndarray, num: int, c_robust=np.float32(1.11)) -> np.ndarray:
    """Indices (into the input) kept by ANMS, in response-descending stable order.

    radius_i = min_j { |pt_i - pt_j| : response_j > response_i * 1.11f }, float32 difference then double
    norm (cv::norm(Point2f) -> sqrt((double)dx*dx + (double)dy*dy)); strongest keeps DBL_MAX; keep
    radius >= (num-th largest radius).  No-op if size < num (visual_odometry.cpp:100)."""
    n = len(response)
    if n < num:
        return np.arange(n)
    order = np.argsort(-response.astype(np.float64), kind="stable")
    p = pt[order].astype(np.float32)
    r = response[order].astype(np.float32)
    thr = (r * np.float32(c_robust)).astype(np.float32)
    rad = np.full(n, np.finfo(np.float64).max, dtype=np.float64)
    for i in range(n):
        m = r[:i] > thr[i]
        if m.any():
            d = (p[i][None, :] - p[:i][m]).astype(np.float32).astype(np.float64)
            rad[i] = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).min()
    final = np.sort(rad)[::-1][num - 1]
    keep = rad >= final
    return order[keep]


def feature_detection(img: np.ndarray, pattern: np.ndarray, nfeatures: int = 3000, anms_keep: int = 500):
    """VO::feature_detection (visual_odometry.cpp:70-94): ORB detect(nfeatures) -> ANMS(anms_keep) ->
    ORB compute.  Output in canonical order (octave asc, response desc, y, x) -- what the stable
    by-octave regrouping inside cv::ORB::compute yields from the response-sorted ANMS output."""
    kp = R.orb_detect(img, nfeatures)
    keep = anms(kp["pt"], kp["response"], anms_keep)
    keep = np.sort(keep)  # canonical order is preserved by index order
    sub = {k: (v[keep] if k != "levels" else v) for k, v in kp.items()}
    desc = R.orb_compute(kp["levels"], sub, pattern)
    return sub, desc


def triangulate_dlt(xl: np.ndarray, xr: np.ndarray, P1: np.ndarray, P2: np.ndarray) -> np.ndarray:
    """Per-match DLT as cv::triangulatePoints: A = [x*P[2]-P[0]; y*P[2]-P[1]] for both views, X = last
    right-singular vector of the 4x4 A.  Returns camera-frame (n,3) float64 (X/W)."""
    n = len(xl)
    out = np.zeros((n, 3), dtype=np.float64)
    for i in range(n):
        A = np.zeros((4, 4))
        A[0] = xl[i, 0] * P1[2] - P1[0]
        A[1] = xl[i, 1] * P1[2] - P1[1]
        A[2] = xr[i, 0] * P2[2] - P2[0]
        A[3] = xr[i, 1] * P2[2] - P2[1]
        _, _, vt = np.linalg.svd(A)
        X = vt[3]
        out[i] = X[:3] / X[3]
    return out


def stereo_projection_matrices(fx, fy, cx, cy, b):
    """Rectified KITTI-style pair: P1 = K[I|0], P2 = K[I|(-b,0,0)] (types_def.hpp:53-54 constants)."""
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    P1 = K @ np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = K @ np.hstack([np.eye(3), np.array([[-b], [0.0], [0.0]])])
    return P1, P2


def depth_gates(p_cam: np.ndarray, T_c_w: np.ndarray):
    """set_ref_3d_position gates (visual_odometry.cpp:194,201): usable 10<Z<400, reliable Z<40;
    world = T_c_w^-1 * p (types_def.cpp:17), narrowed to float32 like cv::Point3f."""
    R_ = T_c_w[:, :3]
    t_ = T_c_w[:, 3]
    z = p_cam[:, 2]
    usable = (z > 10) & (z < 400)
    reliable = usable & (z < 40)
    world = (p_cam - t_[None, :]) @ R_  # R^T (p - t)
    return world.astype(np.float32), usable, reliable
