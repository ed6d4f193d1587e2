"""ORACLE (test infrastructure only -- never imported by the product path).

numpy restatement of `cv::ORB::detectAndCompute` as used by the reference's
`VO::feature_detection` (/root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:70-94,
detector at :22,:31, `detect` at :80, `compute` at :85).  The arithmetic lives in OpenCV
(un-vendored dependency; README.md:72 pins 3.2, the only runnable build here is
opencv-python-headless 4.13.0 -- that build is the parity target).

Pinned by tests/test_oracle_orb.py against live cv2 4.13.0 (stage by stage and end to end) and by
the committed fixtures under tests/golden/.  Every stage is exposed separately so the CUDA
kernels can be compared intermediate by intermediate (pyramid levels, FAST score maps, Harris
responses, angles, blurred levels, descriptors).
"""
from __future__ import annotations

import numpy as np

NLEVELS = 8
EDGE = 31
PATCH = 31
HALF_PATCH = 15
FAST_T = 20
HARRIS_BLOCK = 7
HARRIS_K = np.float32(0.04)

# FAST ring, OpenCV order (features2d/src/fast_score.cpp makeOffsets, patternSize 16)
RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3),
        (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]

UMAX = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


# ----------------------------------------------------------------------------------------------
# geometry of the pyramid (orb.cpp: getScale, level sizes, per-level quotas)
# ----------------------------------------------------------------------------------------------
def level_scales(nlevels: int = NLEVELS) -> np.ndarray:
    sf = float(np.float32(1.2))
    return np.array([np.float32(sf ** l) for l in range(nlevels)], dtype=np.float32)


def cv_round(x: float) -> int:
    return int(np.rint(x))  # round-half-even, as cvRound (SSE cvtsd2si)


def level_sizes(w: int, h: int, nlevels: int = NLEVELS):
    sc = level_scales(nlevels)
    return [(cv_round(w / float(s)), cv_round(h / float(s))) for s in sc]


def level_quotas(nfeatures: int, nlevels: int = NLEVELS):
    sf = float(np.float32(1.2))
    factor = np.float32(1.0 / sf)
    nd = np.float32(nfeatures * (1.0 - float(factor)) / (1.0 - float(np.float32(float(factor) ** nlevels))))
    out, s = [], 0
    for _ in range(nlevels - 1):
        n = cv_round(float(nd))
        out.append(n)
        s += n
        nd = np.float32(nd * factor)
    out.append(max(nfeatures - s, 0))
    return out


# ----------------------------------------------------------------------------------------------
# K1: resize INTER_LINEAR_EXACT (imgproc/src/resize.cpp, fixed-point 8.8 path)
# ----------------------------------------------------------------------------------------------
def _lin_coeffs(dst: int, src: int):
    scale = 1.0 / (dst / src)  # OpenCV: inv_scale = dsize/ssize; scale = 1/inv_scale
    off = np.zeros(dst, dtype=np.int64)
    c0 = np.zeros(dst, dtype=np.int64)
    c1 = np.zeros(dst, dtype=np.int64)
    for v in range(dst):
        f = scale * (v + 0.5) - 0.5
        i = int(np.floor(f))
        if i < 0:
            off[v], c0[v], c1[v] = 0, 256, 0
        elif i >= src - 1:
            off[v], c0[v], c1[v] = src - 1, 256, 0
        else:
            a = int(np.rint((f - i) * 256.0))
            off[v], c0[v], c1[v] = i, 256 - a, a
    return off, c0, c1


def resize_linear_exact(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape
    ox, cx0, cx1 = _lin_coeffs(dw, sw)
    oy, cy0, cy1 = _lin_coeffs(dh, sh)
    s = src.astype(np.int64)
    ox1 = np.minimum(ox + 1, sw - 1)
    hh = cx0[None, :] * s[:, ox] + cx1[None, :] * s[:, ox1]  # 8.8 fixed point rows
    oy1 = np.minimum(oy + 1, sh - 1)
    out = (cy0[:, None] * hh[oy, :] + cy1[:, None] * hh[oy1, :] + 32768) >> 16
    return out.astype(np.uint8)


def build_pyramid(img: np.ndarray, nlevels: int = NLEVELS):
    h, w = img.shape
    sizes = level_sizes(w, h, nlevels)
    levels = [np.ascontiguousarray(img)]
    for l in range(1, nlevels):
        levels.append(resize_linear_exact(levels[l - 1], sizes[l][0], sizes[l][1]))
    return levels


# ----------------------------------------------------------------------------------------------
# K2: FAST-9/16 score + 3x3 NMS (features2d/src/fast.cpp, fast_score.cpp)
# ----------------------------------------------------------------------------------------------
def fast_score_map(img: np.ndarray, threshold: int = FAST_T) -> np.ndarray:
    """score map (int32): cornerScore-1 convention, 0 for non-corners; defined on 3<=x<w-3, 3<=y<h-3."""
    h, w = img.shape
    I = img.astype(np.int32)
    hc, wc = h - 6, w - 6
    c = I[3:h - 3, 3:w - 3]
    d = np.empty((16 + 9, hc, wc), dtype=np.int32)
    for k, (dx, dy) in enumerate(RING):
        d[k] = c - I[3 + dy:3 + dy + hc, 3 + dx:3 + dx + wc]
    d[16:25] = d[0:9]
    best = np.full((hc, wc), -10 ** 9, dtype=np.int32)
    for k in range(16):
        mn = d[k:k + 9].min(axis=0)
        mx = d[k:k + 9].max(axis=0)
        best = np.maximum(best, np.maximum(mn, -mx))
    score = np.zeros((h, w), dtype=np.int32)
    score[3:h - 3, 3:w - 3] = np.where(best > threshold, best - 1, 0)
    return score


def fast_nms(score: np.ndarray):
    """keypoints surviving 3x3 NMS (strictly greater than all 8 neighbours), row-major order."""
    h, w = score.shape
    s = score
    c = s[1:-1, 1:-1]
    keep = c > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            keep &= c > s[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
    # OpenCV only emits for 3<=x<w-3 and rows 3<=y<h-3 (score map is zero outside anyway)
    ys, xs = np.nonzero(keep)
    ys = ys + 1
    xs = xs + 1
    return xs.astype(np.int32), ys.astype(np.int32), s[ys, xs].astype(np.int32)


# ----------------------------------------------------------------------------------------------
# K4: Harris response (orb.cpp HarrisResponses, blockSize 7, k=0.04)
# ----------------------------------------------------------------------------------------------
def harris_responses(img: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    I = img.astype(np.int32)
    r = HARRIS_BLOCK // 2
    scale = np.float32(1.0) / (np.float32(4 * HARRIS_BLOCK) * np.float32(255.0))
    s4 = ((scale * scale) * scale) * scale
    out = np.empty(len(xs), dtype=np.float32)
    # vectorised over keypoints
    a = np.zeros(len(xs), dtype=np.int64)
    b = np.zeros(len(xs), dtype=np.int64)
    c = np.zeros(len(xs), dtype=np.int64)
    for dy in range(-r, HARRIS_BLOCK - r):
        for dx in range(-r, HARRIS_BLOCK - r):
            y = ys + dy
            x = xs + dx
            ix = (I[y, x + 1] - I[y, x - 1]) * 2 + (I[y - 1, x + 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y + 1, x - 1])
            iy = (I[y + 1, x] - I[y - 1, x]) * 2 + (I[y + 1, x - 1] - I[y - 1, x - 1]) + (I[y + 1, x + 1] - I[y - 1, x + 1])
            a += ix * ix
            b += iy * iy
            c += ix * iy
    af = a.astype(np.int32).astype(np.float32)
    bf = b.astype(np.int32).astype(np.float32)
    cf = c.astype(np.int32).astype(np.float32)
    t1 = af * bf
    t2 = cf * cf
    t3 = t1 - t2
    s = af + bf
    t4 = (HARRIS_K * s) * s
    out = (t3 - t4) * s4
    return out.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# K6: intensity-centroid angle (orb.cpp ICAngles + core fastAtan2)
# ----------------------------------------------------------------------------------------------
_F = np.float32
_ATAN_SCALE = _F(180.0 / np.pi)
_P1 = _F(0.9997878412794807) * _ATAN_SCALE
_P3 = _F(-0.3258083974640975) * _ATAN_SCALE
_P5 = _F(0.1555786518463281) * _ATAN_SCALE
_P7 = _F(-0.04432655554792128) * _ATAN_SCALE
_EPS = _F(2.220446049250313e-16)


def fast_atan2(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    y = np.asarray(y, dtype=np.float32)
    x = np.asarray(x, dtype=np.float32)
    ax, ay = np.abs(x), np.abs(y)
    big = ax >= ay
    num = np.where(big, ay, ax)
    den = np.where(big, ax, ay) + _EPS
    c = (num / den).astype(np.float32)
    c2 = (c * c).astype(np.float32)
    a = (((_P7 * c2 + _P5) * c2 + _P3) * c2 + _P1) * c
    a = a.astype(np.float32)
    a = np.where(big, a, _F(90.0) - a).astype(np.float32)
    a = np.where(x < 0, _F(180.0) - a, a).astype(np.float32)
    a = np.where(y < 0, _F(360.0) - a, a).astype(np.float32)
    return a


def ic_angles(img: np.ndarray, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    I = img.astype(np.int32)
    m10 = np.zeros(len(xs), dtype=np.int64)
    m01 = np.zeros(len(xs), dtype=np.int64)
    for u in range(-HALF_PATCH, HALF_PATCH + 1):
        m10 += u * I[ys, xs + u]
    for v in range(1, HALF_PATCH + 1):
        d = UMAX[v]
        vsum = np.zeros(len(xs), dtype=np.int64)
        for u in range(-d, d + 1):
            vp = I[ys + v, xs + u]
            vm = I[ys - v, xs + u]
            vsum += vp - vm
            m10 += u * (vp + vm)
        m01 += v * vsum
    return fast_atan2(m01.astype(np.float32), m10.astype(np.float32))


# ----------------------------------------------------------------------------------------------
# K8: descriptor blur -- generic float32 sepFilter2D, 7x7 sigma 2, REFLECT_101
# ----------------------------------------------------------------------------------------------
def gaussian_kernel7() -> np.ndarray:
    """cv::getGaussianKernel(7, 2, CV_32F): double exp, double normalise, narrow to float32."""
    x = np.arange(7, dtype=np.float64) - 3.0
    k = np.exp(-0.5 * (x / 2.0) ** 2)
    return (k / k.sum()).astype(np.float32)


def _fma32(a, b, c):
    """Correctly rounded float32 fma(a, b, c): exact product in float64, TwoSum, round-to-odd, narrow."""
    a64 = np.asarray(a, dtype=np.float32).astype(np.float64)
    b64 = np.asarray(b, dtype=np.float32).astype(np.float64)
    c64 = np.asarray(c, dtype=np.float32).astype(np.float64)
    p = a64 * b64  # exact: 24+24 bits
    s = p + c64
    bb = s - p
    e = (p - (s - bb)) + (c64 - bb)  # exact rounding error of s
    s, e = np.broadcast_arrays(s, e)
    s = s.copy()
    bits = s.view(np.int64)
    fix = (e != 0) & ((bits & 1) == 0)
    toward = np.where(e > 0, np.inf, -np.inf)
    s[fix] = np.nextafter(s[fix], toward[fix])
    return s.astype(np.float32)


def blur7(img: np.ndarray) -> np.ndarray:
    """7x7 sigma-2 blur exactly as cv2 4.13.0 (AVX2+FMA dispatch) computes `sepFilter2D(u8 -> u8)` with
    a float32 kernel, REFLECT_101 (found by exhaustive variant search against cv2, 0 mismatches on 15 Mpx):

    row pass    x <  32*floor(w/32): acc = k0*s[-3]; acc = fma(k_i, s[i-3], acc), i = 1..6   (vector body)
                x >= 32*floor(w/32): acc = k0*s[-3]; acc = acc + k_i*s[i-3]  (each op rounded; scalar tail)
    column pass acc = k3*r[0]; acc = fma(r[+i] + r[-i], k[3+i], acc), i = 1..3
    then cvRound (half-even) and saturate to u8.
    """
    k = gaussian_kernel7()
    h, w = img.shape
    p = np.pad(img.astype(np.float32), ((3, 3), (3, 3)), mode="reflect")
    fused = (k[0] * p[:, 0:w]).astype(np.float32)
    unfused = fused.copy()
    for i in range(1, 7):
        fused = _fma32(k[i], p[:, i:i + w], fused)
        unfused = (unfused + (k[i] * p[:, i:i + w]).astype(np.float32)).astype(np.float32)
    rows = fused
    tail = (w // 32) * 32
    rows[:, tail:] = unfused[:, tail:]
    out = (k[3] * rows[3:3 + h]).astype(np.float32)
    for i in (1, 2, 3):
        s = (rows[3 + i:3 + i + h] + rows[3 - i:3 - i + h]).astype(np.float32)
        out = _fma32(s, k[3 + i], out)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# K9: rotated BRIEF (orb.cpp computeOrbDescriptors, WTA_K = 2)
# ----------------------------------------------------------------------------------------------
def rbrief(blurred: np.ndarray, cx: np.ndarray, cy: np.ndarray, angle_deg: np.ndarray,
           pattern: np.ndarray) -> np.ndarray:
    """pattern: (256,4) int32 rows (x0,y0,x1,y1).  Returns (n,32) u8."""
    n = len(cx)
    ang = (angle_deg.astype(np.float32) * _F(np.pi / 180.0)).astype(np.float32)
    a = np.cos(ang.astype(np.float64)).astype(np.float32)
    b = np.sin(ang.astype(np.float64)).astype(np.float32)
    px = pattern.astype(np.float32)
    desc = np.zeros((n, 32), dtype=np.uint8)

    def sample(xc, yc):
        # x = px*a - py*b ; y = px*b + py*a   (float32, separate mul and add/sub, no FMA)
        X = ((xc[None, :] * a[:, None]).astype(np.float32) - (yc[None, :] * b[:, None]).astype(np.float32)).astype(np.float32)
        Y = ((xc[None, :] * b[:, None]).astype(np.float32) + (yc[None, :] * a[:, None]).astype(np.float32)).astype(np.float32)
        ix = np.rint(X).astype(np.int64)
        iy = np.rint(Y).astype(np.int64)
        return blurred[cy[:, None] + iy, cx[:, None] + ix]

    v0 = sample(px[:, 0], px[:, 1])
    v1 = sample(px[:, 2], px[:, 3])
    bits = (v0 < v1).astype(np.uint8)  # (n,256)
    for t in range(8):
        desc |= (bits[:, t::8] << t).astype(np.uint8)
    return desc


# ----------------------------------------------------------------------------------------------
# retainBest with ties kept (features2d/src/keypoint.cpp KeyPointsFilter::retainBest)
# ----------------------------------------------------------------------------------------------
def retain_best_mask(resp: np.ndarray, n: int) -> np.ndarray:
    if n <= 0:
        return np.zeros(len(resp), dtype=bool)
    if len(resp) <= n:
        return np.ones(len(resp), dtype=bool)
    cut = np.sort(resp)[::-1][n - 1]
    return resp >= cut


# ----------------------------------------------------------------------------------------------
# full detect (+ compute) in canonical order
# ----------------------------------------------------------------------------------------------
def canonical_order(octave, resp, y, x):
    """(octave asc, response desc, y asc, x asc): the canonical keypoint order (SURVEY §7 hard part 2)."""
    return np.lexsort((x, y, -resp.astype(np.float64), octave))


def orb_detect(img: np.ndarray, nfeatures: int, nlevels: int = NLEVELS, stages: dict | None = None):
    """Returns structured arrays: x,y (level coords, int32), octave, response (f32), angle (f32 deg),
    pt (n,2 float32, level-0 coords), size (f32); canonical order."""
    h, w = img.shape
    sc = level_scales(nlevels)
    levels = build_pyramid(img, nlevels)
    quotas = level_quotas(nfeatures, nlevels)
    X, Y, O, R, A = [], [], [], [], []
    if stages is not None:
        stages["levels"] = levels
        stages["fast"] = []
    for l, lv in enumerate(levels):
        lh, lw = lv.shape
        score = fast_score_map(lv)
        xs, ys, sc_fast = fast_nms(score)
        if stages is not None:
            stages["fast"].append((xs.copy(), ys.copy(), sc_fast.copy()))
        m = (xs >= EDGE) & (xs < lw - EDGE) & (ys >= EDGE) & (ys < lh - EDGE)
        xs, ys, sc_fast = xs[m], ys[m], sc_fast[m]
        m = retain_best_mask(sc_fast.astype(np.float32), 2 * quotas[l])
        xs, ys = xs[m], ys[m]
        hr = harris_responses(lv, xs, ys)
        m = retain_best_mask(hr, quotas[l])
        xs, ys, hr = xs[m], ys[m], hr[m]
        ang = ic_angles(lv, xs, ys)
        X.append(xs); Y.append(ys); O.append(np.full(len(xs), l, np.int32)); R.append(hr); A.append(ang)
    x = np.concatenate(X); y = np.concatenate(Y); o = np.concatenate(O)
    r = np.concatenate(R); a = np.concatenate(A)
    order = canonical_order(o, r, y, x)
    x, y, o, r, a = x[order], y[order], o[order], r[order], a[order]
    pt = np.stack([(x.astype(np.float32) * sc[o]).astype(np.float32),
                   (y.astype(np.float32) * sc[o]).astype(np.float32)], axis=1)
    size = (np.float32(PATCH) * sc[o]).astype(np.float32)
    return dict(x=x, y=y, octave=o, response=r, angle=a, pt=pt, size=size, levels=levels)


def orb_compute(levels, kp: dict, pattern: np.ndarray, stages: dict | None = None) -> np.ndarray:
    """descriptors for keypoints (any order, grouped internally by octave); returns (n,32) in kp order."""
    sc = level_scales(len(levels))
    n = len(kp["octave"])
    desc = np.zeros((n, 32), dtype=np.uint8)
    blurred = {}
    for l in np.unique(kp["octave"]):
        blurred[int(l)] = blur7(levels[int(l)])
        idx = np.nonzero(kp["octave"] == l)[0]
        inv = np.float32(1.0) / sc[int(l)]
        cx = np.rint((kp["pt"][idx, 0] * inv).astype(np.float32)).astype(np.int64)
        cy = np.rint((kp["pt"][idx, 1] * inv).astype(np.float32)).astype(np.int64)
        desc[idx] = rbrief(blurred[int(l)], cx, cy, kp["angle"][idx], pattern)
    if stages is not None:
        stages["blurred"] = blurred
    return desc


def orb_detect_and_compute(img: np.ndarray, nfeatures: int, pattern: np.ndarray):
    kp = orb_detect(img, nfeatures)
    desc = orb_compute(kp["levels"], kp, pattern)
    return kp, desc
