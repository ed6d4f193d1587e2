"""CPU: oracle/pipeline_restate.py (VO::pipeline + Map over live cv2 stages, visual_odometry.cpp:491-706,
map.cpp:13-152) -- self-consistency of the two PnP propagation modes, ground truth, window bookkeeping."""
import numpy as np
import pytest

from oracle import pipeline_restate as PR

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def seq(pkg):
    n = 10
    lefts, rights, t, _ = pkg.synth.synth_sequence(3, n)
    return lefts, rights, t, n


def _run(seq, **kw):
    lefts, rights, t, n = seq
    vo = PR.VO(lambda i: (lefts[i], rights[i]), **kw)
    for _ in range(n):
        vo.step()
    return vo


def test_cv2_pose_and_oracle_pose_propagation_agree(seq):
    """cv2.solvePnPRansac's pose vs the oracle's refit (cross-checked against cv2 on every frame's input):
    identical inlier counts / keyframe decisions / landmark counts, poses to 1e-8"""
    a = _run(seq, pnp="cv2", nfeatures=1000, anms_keep=110)
    b = _run(seq, pnp="oracle", cross_check=True, nfeatures=1000, anms_keep=110)
    for x, y in zip(a.log, b.log):
        assert (x["frame_id"], x["num_inliers"], x["is_keyframe"], x["n_keyframes"], x["n_landmarks"]) == \
               (y["frame_id"], y["num_inliers"], y["is_keyframe"], y["n_keyframes"], y["n_landmarks"])
        assert np.abs(x["T_w_c"] - y["T_w_c"]).max() < 1e-8
    assert sum(r["is_keyframe"] for r in a.log) >= 8          # few features: nearly every frame is a keyframe
    t = seq[2]
    assert max(np.abs(r["T_w_c"][:, 3] - t[r["frame_id"]]).max() for r in a.log) < 0.05


def test_window_eviction_and_landmark_cleanup(seq):
    """Map::insert_keyframe / remove_keyframe / clean_map with a 3-keyframe window: never more than 3 keyframes, an
    evicted keyframe's observations leave its landmarks, landmarks without observations disappear"""
    vo = _run(seq, pnp="oracle", cross_check=False, nfeatures=1000, anms_keep=110, num_keyframes=3)
    assert max(r["n_keyframes"] for r in vo.log) == 3 and len(vo.map.written) >= 5
    alive = set(vo.map.keyframes)
    for lm in vo.map.landmarks.values():
        assert lm.observed_times > 0
        assert all(k in alive for k, _ in lm.observations)
    ids = [fid for fid, _ in vo.map.written]
    assert len(set(ids)) == len(ids)


def test_reference_defaults_track_without_keyframes(seq):
    """ORB(3000) -> ANMS(500): >= 80 inliers and no yaw on this sequence, so no keyframe after the first"""
    vo = _run(seq, pnp="cv2")
    assert all(r["num_inliers"] >= 80 for r in vo.log[1:]) and not any(r["is_keyframe"] for r in vo.log)
    assert vo.log[-1]["n_keyframes"] == 1
