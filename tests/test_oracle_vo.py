"""CPU: pin the VO-stage oracle (oracle/vo_restate.py) against live cv2 4.13.0."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import vo_restate as V


def _cv_match(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
    return (np.array([x.queryIdx for x in m], np.int32), np.array([x.trainIdx for x in m], np.int32),
            np.array([x.distance for x in m], np.float32))


@pytest.mark.parametrize("nq,nt,seed", [(300, 500, 0), (500, 300, 1), (2000, 2000, 2), (1, 1, 3), (17, 1000, 4)])
def test_bf_crosscheck_random(pkg, nq, nt, seed):
    q, t = pkg.synth.synth_descriptor_pair(seed, nq, nt)
    qi, ti, d = V.bf_match_crosscheck(q, t)
    cq, ct, cd = _cv_match(q, t)
    assert np.array_equal(qi, cq) and np.array_equal(ti, ct) and np.array_equal(d.astype(np.float32), cd)


def test_bf_crosscheck_ties(pkg):
    # duplicated rows force exact distance ties in both directions
    q = pkg.synth.synth_descriptors(10, 400, dup_frac=0.3)
    t = np.concatenate([q[::2], pkg.synth.synth_descriptors(11, 300, dup_frac=0.3)])
    qi, ti, d = V.bf_match_crosscheck(q, t)
    cq, ct, cd = _cv_match(q, t)
    assert np.array_equal(qi, cq) and np.array_equal(ti, ct) and np.array_equal(d.astype(np.float32), cd)


def test_gate():
    qi = np.arange(5, dtype=np.int32)
    d = np.array([10, 20, 30, 31, 61], np.int32)
    a, b, c = V.match_gate(qi, qi, d, 1.0)
    assert c.tolist() == [10, 20, 30]
    a, b, c = V.match_gate(qi, qi, d + 10, 1.0)  # 2*min = 40 > 30
    assert c.tolist() == [20, 30, 40]
    a, b, c = V.match_gate(qi[:0], qi[:0], d[:0], 1.0)
    assert len(c) == 0


def test_triangulate_matches_cv2():
    rng = np.random.default_rng(0)
    P1, P2 = V.stereo_projection_matrices(718.856, 718.856, 607.1928, 185.2157, 0.573)
    n = 200
    xl = np.stack([rng.uniform(100, 1100, n), rng.uniform(40, 330, n)], 1).astype(np.float32)
    disp = rng.uniform(1.5, 60, n).astype(np.float32)
    xr = xl.copy()
    xr[:, 0] -= disp
    xr[:, 1] += rng.normal(0, 0.4, n).astype(np.float32)  # imperfect rectification / detection jitter
    ours = V.triangulate_dlt(xl.astype(np.float64), xr.astype(np.float64), P1, P2)
    X = cv2.triangulatePoints(P1, P2, xl.T.astype(np.float64), xr.T.astype(np.float64))
    ref = (X[:3] / X[3]).T
    assert np.allclose(ours, ref, rtol=1e-8, atol=1e-9)
    z = 718.856 * 0.573 / disp
    assert np.allclose(ours[:, 2], z, rtol=2e-3)


def test_anms_reference_loop():
    """literal transcription of visual_odometry.cpp:96-157 vs the vectorised restatement"""
    rng = np.random.default_rng(3)
    n, num = 700, 150
    pt = rng.uniform(0, 1000, (n, 2)).astype(np.float32)
    resp = rng.uniform(1e-6, 1e-3, n).astype(np.float32)
    resp[::7] = resp[3]  # ties
    keep = V.anms(pt, resp, num)
    order = np.argsort(-resp.astype(np.float64), kind="stable")
    p, r = pt[order], resp[order]
    rad = np.empty(n)
    for i in range(n):
        response = np.float32(r[i] * np.float32(1.11))
        radius = np.finfo(np.float64).max
        j = 0
        while j < i and r[j] > response:
            d = (p[i] - p[j]).astype(np.float32)
            radius = min(radius, float(np.sqrt(float(d[0]) * float(d[0]) + float(d[1]) * float(d[1]))))
            j += 1
        rad[i] = radius
    final = np.sort(rad)[::-1][num - 1]
    ref = order[rad >= final]
    assert np.array_equal(keep, ref)
    assert len(keep) >= num
    assert np.array_equal(V.anms(pt[:10], resp[:10], 500), np.arange(10))


def test_golden_match_and_triangulate(pkg):
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(gold, "match_ties_400x500.npz"))
    q = pkg.synth.synth_descriptors(10, 400, dup_frac=0.3)
    t = np.concatenate([q[::2], pkg.synth.synth_descriptors(11, 300, dup_frac=0.3)])
    qi, ti, d = V.bf_match_crosscheck(q, t)
    assert np.array_equal(qi, g["queryIdx"]) and np.array_equal(ti, g["trainIdx"]) and np.array_equal(d, g["distance"])
    g = np.load(os.path.join(gold, "triangulate_64.npz"))
    P1, P2 = V.stereo_projection_matrices(pkg.synth.FX, pkg.synth.FY, pkg.synth.CX, pkg.synth.CY, pkg.synth.BASELINE_M)
    assert np.allclose(V.triangulate_dlt(g["xl"].astype(np.float64), g["xr"].astype(np.float64), P1, P2), g["xyz"], rtol=1e-8)
