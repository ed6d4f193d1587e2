"""GPU parity: K1-K9 through the C-ABI vs the numpy oracle and live cv2 4.13.0 -- bit-exact."""
import numpy as np
import pytest

from oracle import orb_restate as R
from oracle import vo_restate as V

pytestmark = pytest.mark.gpu


def _cv_canonical(img, n):
    import cv2
    kps, desc = cv2.ORB_create(n).detectAndCompute(img, None)
    o = np.array([k.octave for k in kps]); r = np.array([k.response for k in kps], np.float32)
    pt = np.array([k.pt for k in kps], np.float32); a = np.array([k.angle for k in kps], np.float32)
    sz = np.array([k.size for k in kps], np.float32)
    sc = R.level_scales()
    xl = np.rint(pt[:, 0] / sc[o]).astype(np.int32); yl = np.rint(pt[:, 1] / sc[o]).astype(np.int32)
    order = R.canonical_order(o, r, yl, xl)
    return dict(octave=o[order], response=r[order], pt=pt[order], angle=a[order], size=sz[order]), desc[order]


def _assert_kp_equal(kp, desc, ref, ref_desc):
    assert len(kp) == len(ref["octave"])
    assert np.array_equal(kp["octave"], ref["octave"])
    assert np.array_equal(kp["x"], ref["pt"][:, 0]) and np.array_equal(kp["y"], ref["pt"][:, 1])
    assert np.array_equal(kp["response"], ref["response"])
    assert np.array_equal(kp["angle"], ref["angle"])
    assert np.array_equal(kp["size"], ref["size"])
    assert (kp["class_id"] == -1).all()
    assert np.array_equal(desc, ref_desc)


def test_intermediates_vs_oracle(pkg, gpu_ctx):
    """pyramid levels, blurred levels and FAST candidates (score, position) stage by stage"""
    L, _, _ = pkg.synth.synth_pair(0)
    gpu_ctx.orb_detect_compute(L, 2000, 0)
    levels = R.build_pyramid(L)
    for l in range(8):
        lv, bl, cand = gpu_ctx.orb_debug_level(0, l)
        if l > 0:
            assert np.array_equal(lv, levels[l]), f"pyramid level {l}"
        assert np.array_equal(bl, R.blur7(levels[l])), f"blur level {l}"
        xs, ys, sc = R.fast_nms(R.fast_score_map(levels[l]))
        h, w = levels[l].shape
        m = (xs >= 31) & (xs < w - 31) & (ys >= 31) & (ys < h - 31)
        ref = sorted(zip(ys[m].tolist(), xs[m].tolist(), sc[m].tolist()))
        mine = sorted(zip((cand[:, 0] >> 16).tolist(), (cand[:, 0] & 0xFFFF).tolist(), cand[:, 1].tolist()))
        assert ref == mine, f"FAST level {l}: {len(ref)} vs {len(mine)}"


@pytest.mark.parametrize("seed,n", [(0, 2000), (1, 3000), (2, 4000), (3, 500), (4, 2000)])
def test_detect_compute_vs_cv2(pkg, gpu_ctx, seed, n):
    L, Rimg, _ = pkg.synth.synth_pair(seed)
    for img in (L, Rimg):
        kp, desc = gpu_ctx.orb_detect_compute(img, n, 0)
        ref, ref_desc = _cv_canonical(img, n)
        _assert_kp_equal(kp, desc, ref, ref_desc)


def test_detect_compute_vs_oracle_other_geometry(pkg, gpu_ctx, pattern):
    rng = np.random.default_rng(5)
    img = pkg.synth.synth_canvas(9, 752, 360)
    kp, desc = gpu_ctx.orb_detect_compute(img, 1500, 0)
    ref, ref_desc = R.orb_detect_and_compute(img, 1500, pattern)
    _assert_kp_equal(kp, desc, ref, ref_desc)
    ref2, ref_desc2 = _cv_canonical(img, 1500)
    _assert_kp_equal(kp, desc, ref2, ref_desc2)


def test_batch_equals_single(pkg, gpu_ctx):
    imgs = np.stack([pkg.synth.synth_pair(s)[0] for s in range(6)])
    outs = gpu_ctx.orb_detect_compute(imgs, 2000, 0)
    for i in (0, 3, 5):
        kp, desc = gpu_ctx.orb_detect_compute(imgs[i], 2000, 0)
        assert np.array_equal(outs[i][0], kp) and np.array_equal(outs[i][1], desc)


def test_reference_defaults_anms(pkg, gpu_ctx, pattern):
    """the reference operating point: ORB(3000) -> ANMS(500) -> compute (visual_odometry.cpp:80-85)"""
    import cv2
    L, _, _ = pkg.synth.synth_pair(6)
    kp, desc = gpu_ctx.orb_detect_compute(L, 3000, 500, 1.11)
    ref, ref_desc = V.feature_detection(L, pattern, 3000, 500)
    _assert_kp_equal(kp, desc, ref, ref_desc)
    assert len(kp) >= 500
    # and against cv2 itself: detect -> (oracle ANMS on cv2 keypoints) -> cv2 compute
    orb = cv2.ORB_create(3000)
    kps = orb.detect(L)
    pt = np.array([k.pt for k in kps], np.float32); resp = np.array([k.response for k in kps], np.float32)
    keep = V.anms(pt, resp, 500)
    kps2, d2 = cv2.ORB_create().compute(L, [kps[i] for i in keep])
    got = sorted(zip(kp["x"].tolist(), kp["y"].tolist(), kp["octave"].tolist(), map(bytes, desc)))
    want = sorted(zip([k.pt[0] for k in kps2], [k.pt[1] for k in kps2], [k.octave for k in kps2], map(bytes, d2)))
    assert got == want


def test_null_image_and_capacity(pkg, gpu_ctx):
    with pytest.raises(pkg.VslamError) as e:
        gpu_ctx.orb_detect_compute(None)
    assert e.value.status == -1
    with pytest.raises(pkg.VslamError) as e:
        gpu_ctx.orb_detect_compute(np.zeros((400, 1300), np.uint8), 2000, 0)
    assert e.value.status == -2
    # featureless image: zero keypoints, no error
    kp, desc = gpu_ctx.orb_detect_compute(np.full((376, 1241), 128, np.uint8), 2000, 0)
    assert len(kp) == 0 and desc.shape == (0, 32)


@pytest.mark.parametrize("name,img_fn,n", [("orb_seed0_n500.npz", lambda s: s.synth_pair(0)[0], 500),
                                           ("orb_canvas5_400x240_n300.npz", lambda s: s.synth_canvas(5, 400, 240), 300)])
def test_golden_vectors(pkg, gpu_ctx, name, img_fn, n):
    """CUDA path vs the committed cv2-generated fixtures (tests/golden/make_golden.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    kp, desc = gpu_ctx.orb_detect_compute(img_fn(pkg.synth), n, 0)
    _assert_kp_equal(kp, desc, g, g["desc"])
