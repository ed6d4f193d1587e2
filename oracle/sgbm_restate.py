"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of the dense-stereo call the reference makes in VO::disparity_map
(/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:159-174):

    cv::StereoSGBM::create(0, 96, 9, 8*9*9, 32*9*9, 1, 63, 10, 100, 32)->compute(left, right, disp16)
    disp16.convertTo(disparity, CV_32F, 1/16)

The arithmetic lives in OpenCV (calib3d/src/stereosgbm.cpp, un-vendored; parity target = cv2 4.13.0 as installed):

* `prefilter`            -- the x-Sobel + clip table of calcPixelCostBT (both images, per row)
* `pixel_cost_bt`        -- Birchfield-Tomasi cost on the prefiltered and the raw row, raw cost >> 2
* `cost_volume`          -- 9x9 box sum with replicated borders in the [minX1, maxX1) x [0, H) domain, plus P2
* `aggregate`            -- MODE_SGBM single pass: L->R, up-left, up, up-right, then R->L (5 directions), int16 saturation
* `select_disparity`     -- winner-take-all, uniqueness ratio, right-image consistency map, sub-pixel parabola, LR check
* `median3`, `filter_speckles` -- the medianBlur(3) and filterSpeckles post-passes of StereoSGBMImpl::compute
* `sgbm_compute`         -- everything above = cv2.StereoSGBM.compute (CV_16S, disparity*16, invalid = (minD-1)*16)
* `find_3d`              -- Frame::find_3d, types_def.cpp:9-18 (float->int truncation of the keypoint position)

Every intermediate is exposed so that the CUDA kernels can be compared stage by stage.
Pinned by tests/test_oracle_sgbm.py against live cv2.StereoSGBM (bit-exact on the whole disparity image).
"""
from __future__ import annotations

import numpy as np

MAX_COST = 32767
DISP_SHIFT = 4
DISP_SCALE = 16


class Params:
    """The reference's constants (visual_odometry.cpp:163-164)."""

    def __init__(self, min_disparity=0, num_disparities=96, block_size=9, P1=8 * 81, P2=32 * 81, disp12_max_diff=1,
                 pre_filter_cap=63, uniqueness_ratio=10, speckle_window_size=100, speckle_range=32):
        self.minD = min_disparity
        self.D = num_disparities
        self.block = block_size
        self.P1 = P1 if P1 > 0 else 2
        self.P2 = max(P2 if P2 > 0 else 5, self.P1 + 1)
        self.disp12 = disp12_max_diff if disp12_max_diff > 0 else 1
        self.ftzero = max(pre_filter_cap, 15) | 1
        self.uniq = uniqueness_ratio if uniqueness_ratio >= 0 else 10
        self.speckle_window = speckle_window_size
        self.speckle_range = speckle_range
        assert self.minD == 0, "the reference uses minDisparity = 0; the restatement covers that case"


def prefilter(img: np.ndarray, ftzero: int):
    """Returns (sobel_clipped, raw) int32 H x W; the first and last column of BOTH planes hold tab[0] = ftzero
    (the clip table is indexed from its centre, so a zero gradient maps to ftzero)."""
    a = img.astype(np.int32)
    H, W = a.shape
    up = np.vstack([a[:1], a[:-1]])       # row y-1 (row 0 uses itself)
    dn = np.vstack([a[1:], a[-1:]])       # row y+1 (last row uses itself)
    sob = np.full((H, W), ftzero, np.int32)
    g = (a[:, 2:] - a[:, :-2]) * 2 + up[:, 2:] - up[:, :-2] + dn[:, 2:] - dn[:, :-2]
    sob[:, 1:-1] = np.clip(g, -ftzero, ftzero) + ftzero
    raw = np.full((H, W), ftzero, np.int32)
    raw[:, 1:-1] = a[:, 1:-1]
    return sob, raw


def _halfpix_minmax(v: np.ndarray):
    """v0/v1 = min/max of (v, (v+left)/2, (v+right)/2) with the image border falling back to v."""
    vl = v.copy()
    vr = v.copy()
    vl[:, 1:] = (v[:, 1:] + v[:, :-1]) // 2
    vr[:, :-1] = (v[:, :-1] + v[:, 1:]) // 2
    return np.minimum(np.minimum(vl, vr), v), np.maximum(np.maximum(vl, vr), v)


def pixel_cost_bt(left: np.ndarray, right: np.ndarray, p: Params) -> np.ndarray:
    """pixDiff[y, x - minX1, d] (int32) for x in [maxD, W), d in [0, D)."""
    H, W = left.shape
    D = p.D
    minX1 = D
    out = np.zeros((H, W - minX1, D), np.int32)
    xs = np.arange(minX1, W)
    for (l, r), shift in zip(zip(prefilter(left, p.ftzero), prefilter(right, p.ftzero)), (0, 2)):
        u0, u1 = _halfpix_minmax(l)
        v0, v1 = _halfpix_minmax(r)
        for d in range(D):
            u, uu0, uu1 = l[:, xs], u0[:, xs], u1[:, xs]
            v, vv0, vv1 = r[:, xs - d], v0[:, xs - d], v1[:, xs - d]
            c0 = np.maximum(0, np.maximum(u - vv1, vv0 - u))
            c1 = np.maximum(0, np.maximum(v - uu1, uu0 - v))
            out[:, :, d] += np.minimum(c0, c1) >> shift
    return out


def cost_volume(pix: np.ndarray, p: Params) -> np.ndarray:
    """C[y, x, d] = P2 + sum of pixDiff over the block window, borders replicated inside the computed domain."""
    H, W1, D = pix.shape
    r = p.block // 2
    ypad = np.pad(pix, ((r, r), (r, r), (0, 0)), mode="edge").astype(np.int64)
    cs = np.cumsum(ypad, axis=0)
    cs = np.concatenate([np.zeros((1, W1 + 2 * r, D), np.int64), cs], axis=0)
    v = cs[2 * r + 1:] - cs[:H]
    cs = np.cumsum(v, axis=1)
    cs = np.concatenate([np.zeros((H, 1, D), np.int64), cs], axis=1)
    h = cs[:, 2 * r + 1:] - cs[:, :W1]
    return (h + p.P2).astype(np.int32)


def _step(C, Lp, minp, P1, P2):
    """One SGM recurrence step. C (.., D) includes P2; Lp (.., D) predecessor costs; minp (..) their minimum."""
    big = np.full(Lp.shape[:-1] + (1,), MAX_COST, np.int32)
    lm = np.concatenate([big, Lp[..., :-1]], axis=-1) + P1
    lp = np.concatenate([Lp[..., 1:], big], axis=-1) + P1
    delta = (minp + P2)[..., None]
    return C + np.minimum(np.minimum(Lp, delta), np.minimum(lm, lp)) - delta


def aggregate(C: np.ndarray, p: Params, return_parts: bool = False):
    """S[y, x, d] (int32 holding the saturated int16 value) after all five directions."""
    H, W1, D = C.shape
    P1, P2 = p.P1, p.P2
    S = np.zeros((H, W1, D), np.int32)
    parts = {}
    # directions 1..3 come from the previous row: (x-1, y-1), (x, y-1), (x+1, y-1); outside the domain L = 0, min = 0
    Lv = np.zeros((3, H, W1, D), np.int32)
    prev = np.zeros((3, W1 + 2, D), np.int32)
    for y in range(H):
        pm = prev.min(axis=-1)
        for k, off in enumerate((0, 1, 2)):  # predecessor column x-1, x, x+1 in the padded array
            L = _step(C[y], prev[k, off:off + W1], pm[k, off:off + W1], P1, P2)
            Lv[k, y] = L
        prev[:, 1:-1] = Lv[:, y]
    # direction 0: left to right inside the row
    L0 = np.zeros((H, W1, D), np.int32)
    Lp = np.zeros((H, D), np.int32)
    for x in range(W1):
        Lp = _step(C[:, x], Lp, Lp.min(axis=-1), P1, P2)
        L0[:, x] = Lp
    S4 = np.minimum(L0 + Lv[0] + Lv[1] + Lv[2], MAX_COST)
    # fifth direction: right to left, added during the disparity pass
    Lr = np.zeros((H, W1, D), np.int32)
    Lp = np.zeros((H, D), np.int32)
    for x in range(W1 - 1, -1, -1):
        Lp = _step(C[:, x], Lp, Lp.min(axis=-1), P1, P2)
        Lr[:, x] = Lp
    S = np.minimum(S4 + Lr, MAX_COST)
    if return_parts:
        parts = {"L0": L0, "L1": Lv[0], "L2": Lv[1], "L3": Lv[2], "L4": Lr, "S4": S4}
        return S, parts
    return S


def select_disparity(S: np.ndarray, W: int, p: Params) -> np.ndarray:
    """WTA + uniqueness + disp2 map + sub-pixel + LR consistency -> int16 H x W (disparity*16, invalid = -16)."""
    H, W1, D = S.shape
    minX1 = D
    INVALID = (p.minD - 1) * DISP_SCALE
    disp = np.full((H, W), INVALID, np.int32)
    disp2 = np.full((H, W), INVALID, np.int32)
    disp2cost = np.full((H, W), MAX_COST, np.int32)
    best = S.argmin(axis=-1)                      # first minimum = smallest d
    minS = S.min(axis=-1)
    dd = np.arange(D)[None, None, :]
    notuniq = ((S * (100 - p.uniq) < (minS * 100)[..., None]) & (np.abs(best[..., None] - dd) > 1)).any(axis=-1)
    rows = np.arange(H)
    for x in range(W1 - 1, -1, -1):
        ok = ~notuniq[:, x]
        d = best[:, x]
        x2 = x + minX1 - d
        upd = ok & (disp2cost[rows, x2] > minS[:, x])
        disp2cost[rows[upd], x2[upd]] = minS[upd, x]
        disp2[rows[upd], x2[upd]] = d[upd]
    d = best
    inner = (d > 0) & (d < D - 1)
    dm = np.clip(d - 1, 0, D - 1)
    dp = np.clip(d + 1, 0, D - 1)
    Sm = np.take_along_axis(S, dm[..., None], -1)[..., 0]
    Sp_ = np.take_along_axis(S, dp[..., None], -1)[..., 0]
    denom2 = np.maximum(Sm + Sp_ - 2 * minS, 1)
    num = (Sm - Sp_) * DISP_SCALE + denom2
    # C integer division truncates toward zero
    q = np.where(num >= 0, num // (denom2 * 2), -((-num) // (denom2 * 2)))
    sub = np.where(inner, d * DISP_SCALE + q, d * DISP_SCALE)
    disp[:, minX1:] = np.where(notuniq, INVALID, sub)
    # LR check
    xs = np.arange(W)[None, :].repeat(H, 0)
    d1 = disp
    valid = d1 != INVALID
    _d = d1 >> DISP_SHIFT
    d_ = (d1 + DISP_SCALE - 1) >> DISP_SHIFT
    _x = xs - _d
    x_ = xs - d_
    r2 = rows[:, None].repeat(W, 1)
    in1 = (_x >= 0) & (_x < W)
    in2 = (x_ >= 0) & (x_ < W)
    a = disp2[r2, np.clip(_x, 0, W - 1)]
    b = disp2[r2, np.clip(x_, 0, W - 1)]
    bad = valid & in1 & (a >= p.minD) & (np.abs(a - _d) > p.disp12) & in2 & (b >= p.minD) & (np.abs(b - d_) > p.disp12)
    out = np.where(bad, INVALID, disp)
    return out.astype(np.int16)


def median3(d: np.ndarray) -> np.ndarray:
    """cv::medianBlur(ksize 3) on CV_16S, replicated border."""
    pd = np.pad(d, 1, mode="edge")
    H, W = d.shape
    st = np.stack([pd[i:i + H, j:j + W] for i in range(3) for j in range(3)], axis=0)
    return np.sort(st, axis=0)[4].astype(np.int16)


def filter_speckles(d: np.ndarray, new_val: int, max_size: int, max_diff: int) -> np.ndarray:
    """cv::filterSpeckles: 4-connected components under |a-b| <= max_diff among pixels != new_val; components of
    at most max_size pixels become new_val (the result does not depend on the scan order)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    H, W = d.shape
    a = d.astype(np.int32)
    ok = a != new_val
    idx = np.arange(H * W).reshape(H, W)
    eh = ok[:, :-1] & ok[:, 1:] & (np.abs(a[:, :-1] - a[:, 1:]) <= max_diff)
    ev = ok[:-1] & ok[1:] & (np.abs(a[:-1] - a[1:]) <= max_diff)
    src = np.concatenate([idx[:, :-1][eh], idx[:-1][ev]])
    dst = np.concatenate([idx[:, 1:][eh], idx[1:][ev]])
    g = coo_matrix((np.ones(len(src), np.int8), (src, dst)), shape=(H * W, H * W))
    _, lab = connected_components(g, directed=False)
    size = np.bincount(lab)
    small = (size[lab] <= max_size).reshape(H, W) & ok
    out = d.copy()
    out[small] = new_val
    return out


def sgbm_compute(left: np.ndarray, right: np.ndarray, p: Params | None = None, stages: dict | None = None) -> np.ndarray:
    """= cv2.StereoSGBM_create(0, 96, 9, 648, 2592, 1, 63, 10, 100, 32).compute(left, right)  (int16, H x W)."""
    p = p or Params()
    H, W = left.shape
    if W <= p.D:
        return np.full((H, W), (p.minD - 1) * DISP_SCALE, np.int16)
    pix = pixel_cost_bt(left, right, p)
    C = cost_volume(pix, p)
    S = aggregate(C, p)
    raw = select_disparity(S, W, p)
    med = median3(raw)
    out = med
    if p.speckle_window > 0:
        out = filter_speckles(med, (p.minD - 1) * DISP_SCALE, p.speckle_window, DISP_SCALE * p.speckle_range)
    if stages is not None:
        stages.update(pix=pix, C=C, S=S, raw=raw, med=med)
    return out


def disparity_float(disp16: np.ndarray) -> np.ndarray:
    """disparity_sgbm.convertTo(disparity, CV_32F, 1.0 / 16.0f)  (visual_odometry.cpp:168)."""
    return (disp16.astype(np.float64) * (1.0 / 16.0)).astype(np.float32)


def find_3d(kp_xy: np.ndarray, disparity: np.ndarray, fx, fy, cx, cy, b):
    """Frame::find_3d (types_def.cpp:9-18) in the camera frame: disparity.at<float>(kp.pt.y, kp.pt.x) truncates the
    float coordinates to int; depth = fx*b/disp (negative for the invalid value -1, rejected by the caller's gate)."""
    x = (kp_xy[:, 0].astype(np.float64) - cx) / fx
    y = (kp_xy[:, 1].astype(np.float64) - cy) / fy
    dv = disparity[kp_xy[:, 1].astype(np.int32), kp_xy[:, 0].astype(np.int32)].astype(np.float64)
    with np.errstate(divide="ignore"):
        depth = fx * b / dv
    return np.stack([x * depth, y * depth, depth], axis=1)
