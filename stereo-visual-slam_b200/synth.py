"""Seeded synthetic workloads for the stereo-VO + BA hot path.

KITTI is not on disk and there is no network, so every config of BASELINE.json is
driven by these generators.  They are pure numpy (no cv2, no torch) so the tests,
the oracle and bench.py all see byte-identical inputs for a given seed.

Intrinsics are the reference's KITTI-00 constants
(/root/reference/include/stereo_visual_slam_main/types_def.hpp:53-54,
 /root/reference/src/run_vslam.cpp:34-35).
"""
from __future__ import annotations

import numpy as np

W, H = 1241, 376
FX = FY = 718.856
CX, CY = 607.1928, 185.2157
BASELINE_M = 0.573


def kitti_K() -> np.ndarray:
    return np.array([[FX, 0.0, CX], [0.0, FY, CY], [0.0, 0.0, 1.0]], dtype=np.float64)


def stereo_projection_matrices(fx: float = FX, fy: float = FY, cx: float = CX, cy: float = CY, b: float = BASELINE_M):
    """Rectified KITTI-style pair: P1 = K [I | 0], P2 = K [I | (-b, 0, 0)] (types_def.hpp:53-54 constants)."""
    K = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])
    P1 = K @ np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = K @ np.hstack([np.eye(3), np.array([[-b], [0.0], [0.0]])])
    return P1, P2


def _gauss_blur_sep(img: np.ndarray, sigma: float) -> np.ndarray:
    r = int(np.ceil(3 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    p = np.pad(img.astype(np.float64), ((r, r), (r, r)), mode="reflect")
    tmp = np.zeros((p.shape[0], img.shape[1]), dtype=np.float64)
    for i, kv in enumerate(k):
        tmp += kv * p[:, i:i + img.shape[1]]
    out = np.zeros(img.shape, dtype=np.float64)
    for i, kv in enumerate(k):
        out += kv * tmp[i:i + img.shape[0], :]
    return out


def synth_canvas(seed: int, w: int = W + 200, h: int = H, n_rect: int = 400) -> np.ndarray:
    """Textured canvas: blurred uniform noise stretched to 0..255 plus filled rectangles."""
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    b = _gauss_blur_sep(noise, 1.5)
    b = (b - b.min()) / (b.max() - b.min()) * 255.0
    canvas = np.rint(b).astype(np.uint8)
    for _ in range(n_rect):
        rw = int(rng.integers(5, 40))
        rh = int(rng.integers(5, 40))
        x0 = int(rng.integers(0, w - rw))
        y0 = int(rng.integers(0, h - rh))
        g = int(rng.integers(0, 255))
        canvas[y0:y0 + rh, x0:x0 + rw] = g
    return canvas


def synth_canvas_street(seed: int, w: int = W + 200, h: int = H) -> np.ndarray:
    """Canvas with the corner statistics of a street scene: smooth shading, soft-edged flat objects, patches of fine
    texture on a fraction of the area and a little sensor noise.  About 3 % of its pixels pass FAST-9/16 at t = 20
    (KITTI frames: 1-3 %), against 17 % for `synth_canvas` -- ORB(2000) still fills its quota on it."""
    rng = np.random.default_rng(seed)
    base = _gauss_blur_sep(rng.integers(0, 256, size=(h, w), dtype=np.uint8), 12.0)
    img = (base - base.min()) / (base.max() - base.min()) * 120.0 + 60.0
    for _ in range(260):
        rw = int(rng.integers(8, 90))
        rh = int(rng.integers(8, 60))
        x0 = int(rng.integers(0, w - rw))
        y0 = int(rng.integers(0, h - rh))
        img[y0:y0 + rh, x0:x0 + rw] = img[y0:y0 + rh, x0:x0 + rw] * 0.3 + float(rng.integers(20, 235)) * 0.7
    fine = _gauss_blur_sep(rng.integers(0, 256, size=(h, w), dtype=np.uint8), 1.0)
    fine = (fine - fine.mean()) / fine.std()
    mask = _gauss_blur_sep((rng.random((h, w)) < 0.0008).astype(np.uint8) * 255, 10.0)
    mask = np.clip(mask / mask.max() * 3.0, 0.0, 1.0)
    img = _gauss_blur_sep(np.clip(img + fine * 22.0 * mask, 0, 255).astype(np.uint8), 0.7)
    img = img + rng.normal(0.0, 1.5, size=img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_pair(seed: int, w: int = W, h: int = H, texture: str = "dense"):
    """One rectified stereo pair (left, right, band_disparity[h]).

    left = canvas[:, 100:100+w]; right row y is the same canvas shifted by the
    disparity of y's horizontal band (8 bands, d in {4..80} px), i.e. a point at
    left column x appears at right column x - d.
    """
    canvas = synth_canvas(seed, w + 200, h) if texture == "dense" else synth_canvas_street(seed, w + 200, h)
    rng = np.random.default_rng(seed + 1_000_003)
    left = np.ascontiguousarray(canvas[:, 100:100 + w])
    n_band = 8
    disp = rng.integers(4, 81, size=n_band)
    edges = np.linspace(0, h, n_band + 1).astype(int)
    right = np.empty_like(left)
    d_row = np.empty(h, dtype=np.int32)
    for bnd in range(n_band):
        d = int(disp[bnd])
        y0, y1 = edges[bnd], edges[bnd + 1]
        right[y0:y1] = canvas[y0:y1, 100 + d:100 + d + w]
        d_row[y0:y1] = d
    return left, np.ascontiguousarray(right), d_row


def synth_descriptors(seed: int, n: int, dup_frac: float = 0.0) -> np.ndarray:
    """n random 256-bit descriptors (n x 32 u8); dup_frac rows duplicated to force ties."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    nd = int(n * dup_frac)
    if nd:
        src = rng.integers(0, n, size=nd)
        dst = rng.integers(0, n, size=nd)
        d[dst] = d[src]
    return d


def synth_descriptor_pair(seed: int, nq: int, nt: int, flip_bits: int = 20, match_frac: float = 0.7):
    """Query/train descriptor sets where match_frac of the queries have a noisy twin in train."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    m = int(min(nq, nt) * match_frac)
    qi = rng.permutation(nq)[:m]
    ti = rng.permutation(nt)[:m]
    tw = q[qi].copy()
    for r in range(m):
        bits = rng.integers(0, 256, size=int(rng.integers(0, flip_bits + 1)))
        for bpos in bits:
            tw[r, bpos >> 3] ^= np.uint8(1 << (bpos & 7))
    t[ti] = tw
    return q, t


# ----------------------------------------------------------------------------------------------
# SE3 helpers (Sophus conventions): tangent = [upsilon(3); omega(3)]
# ----------------------------------------------------------------------------------------------
def _hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def se3_exp(xi: np.ndarray):
    """Return (R, t) of Sophus::SE3d::exp(xi)."""
    ups, om = np.asarray(xi[:3], float), np.asarray(xi[3:], float)
    th = np.linalg.norm(om)
    Om = _hat(om)
    if th < 1e-10:
        R = np.eye(3) + Om + 0.5 * Om @ Om
        V = np.eye(3) + 0.5 * Om + Om @ Om / 6.0
    else:
        R = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th ** 2 * Om @ Om
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Om + (th - np.sin(th)) / th ** 3 * Om @ Om
    return R, V @ ups


def synth_ba_problem(seed: int, n_kf: int, n_lm: int, obs_per_lm=(2, 6), n_obs_exact: int | None = None,
                     pix_sigma: float = 0.5, pose_sigma: float = 0.02, lm_sigma: float = 0.1,
                     outlier_frac: float = 0.0):
    """Sliding-window BA problem (SURVEY.md §8d cfg 3 / cfg 5).

    Returns dict with poses (n_kf x 12 row-major [R|t] of T_c_w, float64), points (n_lm x 3 float64,
    narrowed through float32 like Landmark::pt_3d_), obs_pose/obs_point (int32), obs_uv (float64,
    narrowed through float32 like cv::KeyPoint::pt), K (3x3), plus the ground truth.
    Observations are grouped by landmark (landmark-major), contiguous keyframe runs.
    """
    rng = np.random.default_rng(seed)
    K = kitti_K()
    # ground-truth trajectory: +0.8 m/frame in Z, 0.5 deg/frame yaw
    Rs, ts = [], []
    for i in range(n_kf):
        yaw = np.deg2rad(0.5) * i
        R_wc = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        c = np.array([0.05 * i, 0.0, 0.8 * i])
        R_cw = R_wc.T
        Rs.append(R_cw)
        ts.append(-R_cw @ c)
    lo, hi = obs_per_lm
    if n_obs_exact is not None:
        m_fixed = n_obs_exact // n_lm
        assert m_fixed * n_lm == n_obs_exact and m_fixed <= n_kf
    pts, op, ol, uv = [], [], [], []
    l = 0
    while l < n_lm:
        m = m_fixed if n_obs_exact is not None else int(rng.integers(lo, hi + 1))
        m = min(m, n_kf)
        k0 = int(rng.integers(0, n_kf - m + 1))
        # a point in front of the middle keyframe of its run
        kc = k0 + m // 2
        pc = np.array([rng.uniform(-20, 20), rng.uniform(-3, 3), rng.uniform(5, 40)])
        pw = Rs[kc].T @ (pc - ts[kc])
        ok = True
        uvs = []
        for k in range(k0, k0 + m):
            q = Rs[k] @ pw + ts[k]
            if q[2] < 1.0:
                ok = False
                break
            u = FX * q[0] / q[2] + CX
            v = FY * q[1] / q[2] + CY
            uvs.append((u, v))
        if not ok:
            continue
        for j, k in enumerate(range(k0, k0 + m)):
            n = rng.normal(0, pix_sigma, 2)
            if outlier_frac > 0 and rng.uniform() < outlier_frac:
                n = rng.uniform(-60, 60, 2)
            op.append(k)
            ol.append(l)
            uv.append((uvs[j][0] + n[0], uvs[j][1] + n[1]))
        pts.append(pw)
        l += 1
    pts = np.array(pts)
    poses_gt = np.zeros((n_kf, 12))
    poses0 = np.zeros((n_kf, 12))
    for k in range(n_kf):
        T = np.hstack([Rs[k], ts[k][:, None]])
        poses_gt[k] = T.reshape(-1)
        dR, dt = se3_exp(rng.normal(0, pose_sigma, 6))
        T0 = np.hstack([dR @ Rs[k], (dR @ ts[k] + dt)[:, None]])
        poses0[k] = T0.reshape(-1)
    pts0 = (pts + rng.normal(0, lm_sigma, pts.shape)).astype(np.float32).astype(np.float64)
    uv = np.array(uv).astype(np.float32).astype(np.float64)
    return dict(K=K, poses=poses0, points=pts0, obs_pose=np.array(op, dtype=np.int32),
                obs_point=np.array(ol, dtype=np.int32), obs_uv=uv, poses_gt=poses_gt, points_gt=pts)


def synth_sequence(seed: int, n_frames: int, w: int = W, h: int = H):
    """Stereo sequence of a camera translating along +x past 8 fronto-parallel bands (piecewise-planar scene).

    Band disparities are multiples of 4 px in [8, 40] (Z = fx*b/d in [10.3, 51.5] m, inside the reference's usable
    depth gate) and the camera moves b/4 per frame, so every image is an exact integer crop of one canvas: a world
    point of band d at left column x in frame 0 sits at x - i*d/4 in frame i and at x - i*d/4 - d in the right image.
    Returns (lefts, rights, T_w_c translations [n,3]); rotation is identity throughout.
    """
    rng = np.random.default_rng(seed + 77)
    n_band = 8
    disp = 4 * rng.integers(2, 11, size=n_band)
    max_shift = int(disp.max()) * (n_frames // 4 + 2)
    canvas = synth_canvas(seed, w + 200 + max_shift, h, n_rect=400 + max_shift // 3)
    edges = np.linspace(0, h, n_band + 1).astype(int)
    lefts, rights = [], []
    for i in range(n_frames):
        L = np.empty((h, w), np.uint8)
        R = np.empty((h, w), np.uint8)
        for bnd in range(n_band):
            d = int(disp[bnd])
            y0, y1 = edges[bnd], edges[bnd + 1]
            off = 100 + (i * d) // 4
            L[y0:y1] = canvas[y0:y1, off:off + w]
            R[y0:y1] = canvas[y0:y1, off + d:off + d + w]
        lefts.append(L)
        rights.append(R)
    t = np.zeros((n_frames, 3))
    t[:, 0] = np.arange(n_frames) * BASELINE_M / 4.0
    return lefts, rights, t, disp


def write_pgm(path: str, img: np.ndarray):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())
