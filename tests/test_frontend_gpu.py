"""GPU parity: K11 triangulation and the batched stereo frontend (detect + match + triangulate) vs the oracle."""
import numpy as np
import pytest

from oracle import orb_restate as R
from oracle import vo_restate as V

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # north star: poses / landmarks within 1e-4 relative


def _cams(pkg):
    s = pkg.synth
    return V.stereo_projection_matrices(s.FX, s.FY, s.CX, s.CY, s.BASELINE_M)


def test_triangulate_vs_cv2_and_oracle(pkg, gpu_ctx):
    import cv2
    rng = np.random.default_rng(0)
    P1, P2 = _cams(pkg)
    n = 3000
    xl = np.stack([rng.uniform(50, 1200, n), rng.uniform(20, 350, n)], 1).astype(np.float32)
    disp = rng.uniform(0.8, 90, n).astype(np.float32)     # Z from 4.6 m to 515 m: both gates exercised
    xr = xl.copy()
    xr[:, 0] -= disp
    xr[:, 1] += rng.normal(0, 0.5, n).astype(np.float32)  # imperfect epipolar alignment
    T = np.hstack([pkg.synth.se3_exp(np.array([0.3, -0.1, 2.0, 0.02, -0.3, 0.01]))[0],
                   np.array([[0.5], [-0.2], [3.0]])])
    xyz, fl = gpu_ctx.triangulate(xl, xr, P1, P2, T)
    X = cv2.triangulatePoints(P1, P2, xl.T.astype(np.float64), xr.T.astype(np.float64))
    pc = (X[:3] / X[3]).T
    ref_w, usable, reliable = V.depth_gates(pc, T)
    assert np.allclose(V.triangulate_dlt(xl.astype(np.float64), xr.astype(np.float64), P1, P2), pc, rtol=1e-7, atol=1e-8)
    err = np.abs(xyz.astype(np.float64) - ref_w.astype(np.float64)).max(axis=1) / np.linalg.norm(ref_w, axis=1)
    assert err.max() < REL_TOL, err.max()
    # gates are exact except for depths within float rounding of a threshold
    z = pc[:, 2]
    safe = (np.abs(z - 10) > 1e-6) & (np.abs(z - 40) > 1e-6) & (np.abs(z - 400) > 1e-6)
    assert np.array_equal((fl & 1).astype(bool)[safe], usable[safe])
    assert np.array_equal((fl & 2).astype(bool)[safe], reliable[safe])
    assert usable.sum() > 100 and reliable.sum() > 100 and (~usable).sum() > 100
    # identity pose: world == camera frame
    xyz0, _ = gpu_ctx.triangulate(xl, xr, P1, P2, None)
    err0 = np.abs(xyz0 - pc.astype(np.float32)).max(axis=1) / np.linalg.norm(pc, axis=1)
    assert err0.max() < REL_TOL


def test_triangulate_empty(pkg, gpu_ctx):
    P1, P2 = _cams(pkg)
    xyz, fl = gpu_ctx.triangulate(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), P1, P2)
    assert xyz.shape == (0, 3) and fl.shape == (0,)


def _oracle_frontend(L, Rimg, pattern, P1, P2, nfeat):
    import cv2
    kl, dl = R.orb_detect_and_compute(L, nfeat, pattern)
    kr, dr = R.orb_detect_and_compute(Rimg, nfeat, pattern)
    qi, ti, d = V.bf_match_crosscheck(dl, dr)
    qi, ti, d = V.match_gate(qi, ti, d, 1.0)
    X = cv2.triangulatePoints(P1, P2, kl["pt"][qi].T.astype(np.float64), kr["pt"][ti].T.astype(np.float64))
    pc = (X[:3] / X[3]).T
    return kl, dl, kr, dr, qi, ti, d, pc


@pytest.mark.parametrize("seeds", [(0, 1, 2), (7,)])
def test_stereo_frontend_batch(pkg, gpu_ctx, pattern, seeds):
    P1, P2 = _cams(pkg)
    pairs = [pkg.synth.synth_pair(s) for s in seeds]
    left = np.stack([p[0] for p in pairs]); right = np.stack([p[1] for p in pairs])
    out = gpu_ctx.stereo_frontend(left, right, P1, P2, nfeatures=2000)
    b = len(seeds)
    for i in range(b):
        kl, dl, kr, dr, qi, ti, d, pc = _oracle_frontend(left[i], right[i], pattern, P1, P2, 2000)
        nl, nr, nm = out["n_kp"][i], out["n_kp"][b + i], out["n_matches"][i]
        assert nl == len(dl) and nr == len(dr)                      # keypoint counts bit-exact
        assert np.array_equal(out["desc"][i, :nl], dl) and np.array_equal(out["desc"][b + i, :nr], dr)
        assert np.array_equal(out["kp"]["x"][i, :nl], kl["pt"][:, 0])
        m = out["matches"][i, :nm]
        assert nm == len(qi)
        assert np.array_equal(m["queryIdx"], qi) and np.array_equal(m["trainIdx"], ti)   # match indices bit-exact
        assert np.array_equal(m["distance"], d.astype(np.float32))
        xyz = out["xyz"][i, :nm].astype(np.float64)
        ok = np.abs(pc[:, 2]) < 1e4     # DLT of (near-)zero disparity matches is ill-conditioned on both sides
        err = np.abs(xyz[ok] - pc[ok]).max(axis=1) / np.linalg.norm(pc[ok], axis=1)
        assert err.max() < REL_TOL
        # synthetic ground truth: band disparities are integers in 4..80 px -> depths 5..103 m, so roughly half of
        # the bands fall below the reference's 10 m "usable" gate; check the flags against the true band depth
        flags = out["flags"][i, :nm]
        dtrue = pairs[i][2][np.rint(kl["pt"][qi][:, 1]).astype(int).clip(0, 375)]
        ztrue = pkg.synth.FX * pkg.synth.BASELINE_M / dtrue
        correct = np.abs(xyz[:, 2] - ztrue) / ztrue < 0.05          # matches that found the true correspondence
        assert correct.sum() > 0.5 * nm
        zc = ztrue[correct]
        clear = (np.abs(zc - 10) > 1) & (np.abs(zc - 40) > 2)       # away from the gate thresholds
        assert np.array_equal((flags[correct] & 1).astype(bool)[clear], (zc > 10)[clear])
        assert np.array_equal((flags[correct] & 2).astype(bool)[clear], ((zc > 10) & (zc < 40))[clear])
        ref_usable = (pc[:, 2] > 10) & (pc[:, 2] < 400)
        safe = (np.abs(pc[:, 2] - 10) > 1e-6) & (np.abs(pc[:, 2] - 40) > 1e-6) & (np.abs(pc[:, 2] - 400) > 1e-6)
        assert np.array_equal((flags & 1).astype(bool)[safe], ref_usable[safe])


def test_stereo_frontend_chunked_pipeline_equals_per_pair(pkg, gpu_ctx_big):
    """The host-buffer call splits a batch into 16-pair chunks whose H2D / kernels / D2H overlap on three streams;
    the results must be those of the same pairs submitted one at a time (37 pairs: two full chunks + a ragged one)."""
    P1, P2 = _cams(pkg)
    base = [pkg.synth.synth_pair(s)[:2] for s in range(3)]
    b = 37
    left = np.stack([np.roll(base[i % 3][0], 5 * (i // 3), axis=0) for i in range(b)])
    right = np.stack([np.roll(base[i % 3][1], 5 * (i // 3), axis=0) for i in range(b)])
    T = np.tile(np.eye(4)[:3].reshape(1, 12), (b, 1))
    T[:, 3] = np.arange(b) * 0.25       # distinct poses: the pose of pair i must be applied to pair i
    out = gpu_ctx_big.stereo_frontend(left, right, P1, P2, T_c_w=T, nfeatures=1000)
    for i in (0, 1, 15, 16, 17, 31, 32, 36):
        one = gpu_ctx_big.stereo_frontend(left[i:i + 1], right[i:i + 1], P1, P2, T_c_w=T[i:i + 1], nfeatures=1000)
        nl, nr, nm = one["n_kp"][0], one["n_kp"][1], one["n_matches"][0]
        assert (out["n_kp"][i], out["n_kp"][b + i], out["n_matches"][i]) == (nl, nr, nm)
        assert nl > 900 and nm > 50
        assert np.array_equal(out["kp"][i, :nl], one["kp"][0, :nl])
        assert np.array_equal(out["kp"][b + i, :nr], one["kp"][1, :nr])
        assert np.array_equal(out["desc"][i, :nl], one["desc"][0, :nl])
        assert np.array_equal(out["desc"][b + i, :nr], one["desc"][1, :nr])
        assert np.array_equal(out["matches"][i, :nm], one["matches"][0, :nm])
        assert np.array_equal(out["xyz"][i, :nm], one["xyz"][0, :nm])
        assert np.array_equal(out["flags"][i, :nm], one["flags"][0, :nm])


def test_two_stream_chunking_equals_single_stream(pkg):
    """Host-buffer path (3 chunks: 32 + 32 + 8 pairs alternating between two compute streams on disjoint scratch) and
    device-resident path (2 half-batches) give bit-identical results with vslam_ctx_set_concurrency on and off."""
    import torch
    n, nfeat, h, w = 72, 400, 200, 640
    ctx = pkg.Context(device=0, max_images=2 * n, max_width=w, max_height=h, max_keypoints=512)
    try:
        P1, P2 = _cams(pkg)
        L = np.stack([np.ascontiguousarray(pkg.synth.synth_pair(100 + i)[0][60:60 + h, 40 + i:40 + i + w]) for i in range(n)])
        Rr = np.stack([np.ascontiguousarray(pkg.synth.synth_pair(100 + i)[1][60:60 + h, 40 + i:40 + i + w]) for i in range(n)])
        res = {}
        for on in (True, False):
            ctx.set_concurrency(on)
            o = ctx.stereo_frontend(L, Rr, P1, P2, nfeatures=nfeat)
            res[on] = {k: np.array(v, copy=True) for k, v in o.items() if isinstance(v, np.ndarray)}
        a, b = res[True], res[False]
        assert np.array_equal(a["n_kp"], b["n_kp"]) and np.array_equal(a["n_matches"], b["n_matches"])
        assert a["n_kp"].min() > 100 and a["n_matches"].min() > 10
        for i in range(2 * n):
            k = a["n_kp"][i]
            assert a["kp"][i, :k].tobytes() == b["kp"][i, :k].tobytes() and np.array_equal(a["desc"][i, :k], b["desc"][i, :k])
        for i in range(n):
            m = a["n_matches"][i]
            assert a["matches"][i, :m].tobytes() == b["matches"][i, :m].tobytes()
            assert np.array_equal(a["xyz"][i, :m], b["xyz"][i, :m]) and np.array_equal(a["flags"][i, :m], b["flags"][i, :m])
        # device-resident entry: 72 pairs >= 64 -> two half-batches on two streams
        dev = torch.device("cuda:0")
        dl, dr = torch.from_numpy(L).to(dev), torch.from_numpy(Rr).to(dev)
        cap = ctx.kp_cap
        outs = {}
        for on in (True, False):
            ctx.set_concurrency(on)
            d_kp = torch.zeros((2 * n, cap, 7), dtype=torch.int32, device=dev)
            d_desc = torch.zeros((2 * n, cap, 32), dtype=torch.uint8, device=dev)
            d_nkp = torch.zeros(2 * n, dtype=torch.int32, device=dev)
            d_m = torch.zeros((n, cap, 4), dtype=torch.int32, device=dev)
            d_nm = torch.zeros(n, dtype=torch.int32, device=dev)
            d_xyz = torch.zeros((n, cap, 3), dtype=torch.float32, device=dev)
            d_fl = torch.zeros((n, cap), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            ctx.stereo_frontend_dev(dl, dr, n, w, h, w, w * h, P1, P2, None, d_kp, d_desc, d_nkp, d_m, d_nm, d_xyz, d_fl,
                                    nfeatures=nfeat)
            ctx.synchronize()
            outs[on] = [t.cpu().numpy() for t in (d_nkp, d_nm, d_kp, d_desc, d_m, d_xyz)]
        assert np.array_equal(outs[True][0], a["n_kp"]) and np.array_equal(outs[True][1], a["n_matches"])
        for x, y in zip(outs[True], outs[False]):
            assert np.array_equal(x, y)
    finally:
        ctx.set_concurrency(True)
        ctx.close()


def test_repitched_ramped_host_batch_equals_device_path_on_unaligned_rows(pkg):
    """136 pairs with 641-byte rows through the host-buffer call -- re-pitched to 16-byte rows on the device
    (repitch_kernel) and cut into ramped chunks (16 + 32 + 32 + 32 + 24 pairs) -- against the device-resident call on the
    same images at their unaligned pitch (byte-load paths of the FAST / blur kernels): bit-identical results."""
    import torch
    n, nfeat, h, w = 136, 300, 200, 641
    ctx = pkg.Context(device=0, max_images=2 * n, max_width=w, max_height=h, max_keypoints=384)
    try:
        P1, P2 = _cams(pkg)
        base = [pkg.synth.synth_pair(200 + i) for i in range(6)]
        L = np.stack([np.ascontiguousarray(base[i % 6][0][40:40 + h, 30 + i:30 + i + w]) for i in range(n)])
        Rr = np.stack([np.ascontiguousarray(base[i % 6][1][40:40 + h, 30 + i:30 + i + w]) for i in range(n)])
        a = ctx.stereo_frontend(L, Rr, P1, P2, nfeatures=nfeat)
        a = {k: np.array(v, copy=True) for k, v in a.items() if isinstance(v, np.ndarray)}
        assert a["n_kp"].min() > 100 and a["n_matches"].min() > 5
        dev = torch.device("cuda:0")
        dl, dr = torch.from_numpy(L).to(dev), torch.from_numpy(Rr).to(dev)
        cap = ctx.kp_cap
        d_kp = torch.zeros((2 * n, cap, 7), dtype=torch.int32, device=dev)
        d_desc = torch.zeros((2 * n, cap, 32), dtype=torch.uint8, device=dev)
        d_nkp = torch.zeros(2 * n, dtype=torch.int32, device=dev)
        d_m = torch.zeros((n, cap, 4), dtype=torch.int32, device=dev)
        d_nm = torch.zeros(n, dtype=torch.int32, device=dev)
        d_xyz = torch.zeros((n, cap, 3), dtype=torch.float32, device=dev)
        d_fl = torch.zeros((n, cap), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ctx.stereo_frontend_dev(dl, dr, n, w, h, w, w * h, P1, P2, None, d_kp, d_desc, d_nkp, d_m, d_nm, d_xyz, d_fl,
                                nfeatures=nfeat)
        ctx.synchronize()
        nkp, nm = d_nkp.cpu().numpy(), d_nm.cpu().numpy()
        assert np.array_equal(nkp, a["n_kp"]) and np.array_equal(nm, a["n_matches"])
        kp, desc, m, xyz = d_kp.cpu().numpy(), d_desc.cpu().numpy(), d_m.cpu().numpy(), d_xyz.cpu().numpy()
        for i in range(2 * n):
            k = nkp[i]
            assert a["kp"][i, :k].tobytes() == kp[i, :k].tobytes() and np.array_equal(a["desc"][i, :k], desc[i, :k])
        for i in range(n):
            assert a["matches"][i, :nm[i]].tobytes() == m[i, :nm[i]].tobytes()
            assert np.array_equal(a["xyz"][i, :nm[i]], xyz[i, :nm[i]])
    finally:
        ctx.close()


def test_begin_end_on_two_contexts_equals_the_synchronous_call(pkg):
    """vslam_stereo_frontend_batch_begin / _end with two contexts alternating (two batches in flight) returns what
    vslam_stereo_frontend_batch returns; a second _begin on a busy context and an _end without a batch are refused."""
    n, nfeat, h, w = 40, 300, 200, 640
    P1, P2 = _cams(pkg)
    base = [pkg.synth.synth_pair(300 + i) for i in range(4)]
    batches = []
    for k in range(3):
        L = np.stack([np.ascontiguousarray(base[(i + k) % 4][0][30:30 + h, 20 + i:20 + i + w]) for i in range(n)])
        Rr = np.stack([np.ascontiguousarray(base[(i + k) % 4][1][30:30 + h, 20 + i:20 + i + w]) for i in range(n)])
        batches.append((L, Rr))
    ctxs = [pkg.Context(device=0, max_images=2 * n, max_width=w, max_height=h, max_keypoints=384) for _ in range(2)]
    try:
        ref = []
        for L, Rr in batches:
            o = ctxs[0].stereo_frontend(L, Rr, P1, P2, nfeatures=nfeat)
            ref.append({k: np.array(v, copy=True) for k, v in o.items()})
        outs = [ctxs[i & 1].alloc_frontend_outputs(n) for i in range(3)]
        ctxs[0].stereo_frontend_begin(batches[0][0], batches[0][1], P1, P2, outs[0], nfeatures=nfeat)
        with pytest.raises(pkg.ffi.VslamError):
            ctxs[0].stereo_frontend_begin(batches[1][0], batches[1][1], P1, P2, outs[1], nfeatures=nfeat)
        ctxs[1].stereo_frontend_begin(batches[1][0], batches[1][1], P1, P2, outs[1], nfeatures=nfeat)
        ctxs[0].stereo_frontend_end()
        ctxs[0].stereo_frontend_begin(batches[2][0], batches[2][1], P1, P2, outs[2], nfeatures=nfeat)
        ctxs[1].stereo_frontend_end()
        ctxs[0].stereo_frontend_end()
        with pytest.raises(pkg.ffi.VslamError):
            ctxs[0].stereo_frontend_end()
        for o, r in zip(outs, ref):
            assert np.array_equal(o["n_kp"], r["n_kp"]) and np.array_equal(o["n_matches"], r["n_matches"])
            assert r["n_kp"].min() > 100
            for i in range(2 * n):
                k = r["n_kp"][i]
                assert o["kp"][i, :k].tobytes() == r["kp"][i, :k].tobytes() and np.array_equal(o["desc"][i, :k], r["desc"][i, :k])
            for i in range(n):
                m = r["n_matches"][i]
                assert o["matches"][i, :m].tobytes() == r["matches"][i, :m].tobytes()
                assert np.array_equal(o["xyz"][i, :m], r["xyz"][i, :m]) and np.array_equal(o["flags"][i, :m], r["flags"][i, :m])
    finally:
        for c in ctxs:
            c.close()
