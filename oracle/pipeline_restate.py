"""ORACLE (test infrastructure only -- never imported by the product path).

Python restatement of the reference's per-frame state machine and sliding-window bookkeeping, driving LIVE cv2 4.13
for every OpenCV call the reference makes:

  VO::pipeline / initialization / tracking      /root/reference/src/stereo_visual_slam_main/visual_odometry.cpp:491-706
  VO::feature_detection (ORB 3000 -> ANMS 500)   visual_odometry.cpp:70-94, 96-157      cv2.ORB_create().detect/.compute
  VO::feature_matching                          visual_odometry.cpp:219-251            cv2.BFMatcher(NORM_HAMMING, True)
  VO::disparity_map / Frame::find_3d /          visual_odometry.cpp:159-217,           cv2.StereoSGBM_create(0, 96, 9, ...)
      set_ref_3d_position                       types_def.cpp:9-18
  VO::motion_estimation                         visual_odometry.cpp:253-314            cv2.solvePnPRansac / cv2.Rodrigues
  VO::check_motion_estimation, insert_key_frame visual_odometry.cpp:316-432
  Map::insert_keyframe / remove_keyframe /      map.cpp:13-152
      clean_map / insert_landmark
  main loop incl. the four BA calls             /root/reference/src/run_vslam.cpp:39-84, optimization.cpp:103-436
                                                (graph construction here, LM through oracle/ba_oracle.c)

Decisions where the reference is nondeterministic: keypoints are put in canonical order (octave asc,
response desc, y, x) before cv::ORB::compute -- the reference's std::sort inside ANMS leaves response ties in
unspecified order; unordered_map iteration order only affects summation order inside BA.

`pnp="cv2"` propagates cv2.solvePnPRansac's own pose (Levenberg-Marquardt stopped at FLT_EPSILON, ~1e-9 from the
optimum); `pnp="oracle"` propagates oracle/pnp_oracle's (same inlier list, Gauss-Newton to convergence) and checks
cv2's answer on the same input at every frame (`cross_check=True`).  Poses that differ by 1e-9 round a few landmark
coordinates per keyframe to neighbouring float32 values, so for the index-exact comparison with the GPU path the
oracle-refit mode is the one whose arithmetic the product shares.
"""
from __future__ import annotations

import numpy as np

from . import vo_restate as V

FX = FY = 718.856
CX, CY = 607.1928, 185.2157
BASELINE = 0.573


# ------------------------------------------------------------------ Sophus::SE3d ----------------
class SE3:
    __slots__ = ("R", "t")

    def __init__(self, R=None, t=None):
        self.R = np.eye(3) if R is None else np.array(R, dtype=np.float64)
        self.t = np.zeros(3) if t is None else np.array(t, dtype=np.float64)

    def inverse(self):
        Rt = self.R.T.copy()
        return SE3(Rt, -_mv(Rt, self.t))

    def __mul__(self, o):
        if isinstance(o, SE3):
            return SE3(_mm(self.R, o.R), _mv(self.R, o.t) + self.t)
        return _mv(self.R, o) + self.t

    def angleY(self):
        return float(np.arctan2(-self.R[2, 0], np.hypot(self.R[2, 1], self.R[2, 2])))

    def log(self):
        R = self.R
        c = min(1.0, max(-1.0, 0.5 * (R[0, 0] + R[1, 1] + R[2, 2] - 1.0)))
        th = np.arccos(c)
        w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
        f = 0.5 + th * th / 12.0 if th < 1e-10 else th / (2.0 * np.sin(th))
        if np.pi - th < 1e-6:
            ax = np.sqrt(np.maximum(0.0, (np.diag(R) - c) / (1 - c)))
            w = np.where(w < 0, -ax, ax) * th
            f = 1.0
        w = w * f
        th2 = float(w @ w)
        t = np.sqrt(th2)
        W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        k = 1.0 / 12.0 if t < 1e-10 else (1.0 - t * np.cos(0.5 * t) / (2.0 * np.sin(0.5 * t))) / th2
        Vi = np.eye(3) - 0.5 * W + k * (W @ W)
        return np.concatenate([Vi @ self.t, w])

    def matrix34(self):
        return np.hstack([self.R, self.t[:, None]])


def _mv(M, p):   # row-by-row, left to right, like the host layer's 3x3 helpers
    return np.array([M[0, 0] * p[0] + M[0, 1] * p[1] + M[0, 2] * p[2],
                     M[1, 0] * p[0] + M[1, 1] * p[1] + M[1, 2] * p[2],
                     M[2, 0] * p[0] + M[2, 1] * p[1] + M[2, 2] * p[2]])


def _mm(A, B):
    out = np.empty((3, 3))
    for i in range(3):
        for j in range(3):
            out[i, j] = A[i, 0] * B[0, j] + A[i, 1] * B[1, j] + A[i, 2] * B[2, j]
    return out


# ------------------------------------------------------------------ containers (types_def.hpp:17-121) -------------
class Feature:
    __slots__ = ("feature_id", "frame_id", "landmark_id", "pt", "desc", "is_inlier")

    def __init__(self, feature_id, frame_id, pt, desc):
        self.feature_id, self.frame_id, self.landmark_id = feature_id, frame_id, -1
        self.pt, self.desc, self.is_inlier = pt, desc, False      # pt = (float32 x, float32 y)


class Frame:
    def __init__(self):
        self.frame_id = 0
        self.left = self.right = self.disparity = None
        self.T_c_w = SE3()
        self.is_keyframe = False
        self.keyframe_id = 0
        self.features = []

    def copy(self):   # frames are copied by value (move_frame, Map::insert_keyframe)
        f = Frame()
        f.frame_id, f.left, f.right, f.disparity = self.frame_id, self.left, self.right, self.disparity
        f.T_c_w = SE3(self.T_c_w.R, self.T_c_w.t)
        f.is_keyframe, f.keyframe_id = self.is_keyframe, self.keyframe_id
        f.features = []
        for s in self.features:
            d = Feature(s.feature_id, s.frame_id, s.pt, s.desc)
            d.landmark_id, d.is_inlier = s.landmark_id, s.is_inlier
            f.features.append(d)
        return f

    def find_3d(self, pt):
        """types_def.cpp:9-18: disparity looked up at the truncated pixel; returns (world, relative) float64"""
        d = self.disparity[int(pt[1]), int(pt[0])]                       # cv::Mat::at<float>(float, float)
        depth = FX * BASELINE / float(d) if d != 0 else np.inf
        rel = np.array([(float(pt[0]) - CX) / FX * depth, (float(pt[1]) - CY) / FY * depth, depth])
        return self.T_c_w.inverse() * rel, rel


class Landmark:
    def __init__(self, landmark_id, pt_3d, desc, reliable_depth, obs):
        self.landmark_id, self.pt_3d, self.desc = landmark_id, pt_3d, desc     # pt_3d float32[3]
        self.observed_times = 1
        self.observations = [obs]          # (keyframe_id, feature_id)
        self.is_inlier = True
        self.reliable_depth = reliable_depth


class Map:
    """map.cpp:13-152"""

    def __init__(self, num_keyframes=10):
        self.keyframes, self.landmarks = {}, {}
        self.num_keyframes = num_keyframes
        self.current_keyframe_id = 0
        self.written = []                  # (frame id, T_w_c 3x4) in the order Map::write_pose is called

    def insert_keyframe(self, frame):
        self.current_keyframe_id = frame.keyframe_id
        self.keyframes[frame.keyframe_id] = frame.copy()
        if len(self.keyframes) > self.num_keyframes:
            self.remove_keyframe()

    def insert_landmark(self, lm):
        self.landmarks[lm.landmark_id] = lm

    def remove_keyframe(self):
        T_w_c = self.keyframes[self.current_keyframe_id].T_c_w.inverse()
        max_d, min_d, max_id, min_id = 0.0, 1000000.0, 0, 0
        for kid, kf in self.keyframes.items():
            if kid == self.current_keyframe_id:
                continue
            d = float(np.linalg.norm((kf.T_c_w * T_w_c).log()))
            if d > max_d:
                max_d, max_id = d, kid
            if d < min_d:
                min_d, min_id = d, kid
        victim = min_id if min_d < 0.2 else max_id
        for feat in self.keyframes[victim].features:
            lm = self.landmarks[feat.landmark_id]
            lm.observations = [o for o in lm.observations if not (o[0] == victim and o[1] == feat.feature_id)]
            lm.observed_times -= 1
        self.written.append((self.keyframes[victim].frame_id, self.keyframes[victim].T_c_w.inverse().matrix34()))
        del self.keyframes[victim]
        self.landmarks = {k: v for k, v in self.landmarks.items() if v.observed_times != 0}   # clean_map


# ------------------------------------------------------------------ the front end ---------------------------------
class VO:
    def __init__(self, read_img, nfeatures=3000, anms_keep=500, num_keyframes=10, pnp="cv2", cross_check=True,
                 do_ba=False):
        import cv2
        self.cv2 = cv2
        self.read_img = read_img
        self.detector = cv2.ORB_create(nfeatures)
        self.descriptor = cv2.ORB_create()
        self.matcher = cv2.BFMatcher(cv2.NORM_HAMMING, True)
        self.sgbm = cv2.StereoSGBM_create(0, 96, 9, 8 * 9 * 9, 32 * 9 * 9, 1, 63, 10, 100, 32)
        self.anms_keep = anms_keep
        self.map = Map(num_keyframes)
        self.K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1.0]])
        self.pnp_mode, self.cross_check, self.do_ba = pnp, cross_check, do_ba
        self.state = "Init"
        self.frame_last, self.frame_current = Frame(), Frame()
        self.T_c_w, self.T_c_l = SE3(), SE3()
        self.num_inliers = 0
        self.seq = 1
        self.num_lost = 0
        self.curr_keyframe_id = self.curr_landmark_id = 0
        self.log = []                      # per-frame records (see step())
        self.stage_log = []                # per-frame stage outputs for stage-level comparisons

    # -- stages ------------------------------------------------------------------------------------------------
    def feature_detection(self, img):
        cv2 = self.cv2
        kps = self.detector.detect(img)
        pt = np.array([k.pt for k in kps], dtype=np.float32).reshape(-1, 2)
        resp = np.array([k.response for k in kps], dtype=np.float32)
        keep = V.anms(pt, resp, self.anms_keep) if self.anms_keep > 0 else np.arange(len(kps))
        kps = [kps[i] for i in keep]
        kps.sort(key=lambda k: (k.octave, -k.response, k.pt[1], k.pt[0]))        # canonical order
        kps, desc = self.descriptor.compute(img, kps)
        pts = [(np.float32(k.pt[0]), np.float32(k.pt[1])) for k in kps]
        return pts, (desc if desc is not None else np.zeros((0, 32), np.uint8))

    def feature_matching(self, d1, d2):
        if len(d1) == 0 or len(d2) == 0:
            return []
        m = self.matcher.match(d1, d2)
        if not m:
            return []
        gap = self.frame_current.frame_id - self.frame_last.frame_id
        thr = max(2.0 * min(x.distance for x in m), 30.0 * gap)
        return [(x.queryIdx, x.trainIdx) for x in m if x.distance <= thr]

    def disparity_map(self, frame):
        d16 = self.sgbm.compute(frame.left, frame.right)
        frame.disparity = d16.astype(np.float32) * np.float32(1.0 / 16.0)

    def set_ref_3d_position(self, pts, desc, frame):
        pts_3d, kept, reliable = [], [], []
        for i, p in enumerate(pts):
            world, rel = frame.find_3d(p)
            if rel[2] > 10 and rel[2] < 400:
                pts_3d.append(world.astype(np.float32))
                kept.append(i)
                reliable.append(bool(rel[2] < 40))
        return pts_3d, [pts[i] for i in kept], desc[kept], reliable

    def motion_estimation(self, frame):
        cv2 = self.cv2
        n = len(frame.features)
        xyz = np.array([self.map.landmarks[f.landmark_id].pt_3d for f in frame.features], dtype=np.float32).reshape(-1, 3)
        uv = np.array([f.pt for f in frame.features], dtype=np.float32).reshape(-1, 2)
        inl, T = np.zeros(0, np.int32), SE3()
        if n >= 5:
            ok, rvec, tvec, cinl = cv2.solvePnPRansac(xyz, uv, self.K, None, iterationsCount=100, reprojectionError=4.0,
                                                      confidence=0.99)
            cinl = cinl.ravel().astype(np.int32) if ok else np.zeros(0, np.int32)
            if self.pnp_mode == "cv2":
                if ok:
                    inl, T = cinl, SE3(cv2.Rodrigues(rvec)[0], tvec.ravel())
            else:
                from . import pnp_oracle as P
                o = P.solve_pnp_ransac(xyz, uv, self.K)
                if self.cross_check:
                    assert o["ok"] == ok and np.array_equal(o["inliers"], cinl), "pnp oracle != cv2 on a pipeline frame"
                    if ok:
                        assert np.abs(o["rvec"] - rvec.ravel()).max() < 1e-6 and np.abs(o["tvec"] - tvec.ravel()).max() < 1e-6
                if o["ok"]:
                    inl, T = o["inliers"], SE3(o["T_c_w"][:, :3], o["T_c_w"][:, 3])
        self.stage_log[-1].update(pnp_xyz=xyz, pnp_uv=uv, pnp_inliers=inl)
        self.num_inliers = len(inl)
        self.T_c_w = T
        for i in inl:
            frame.features[i].is_inlier = True
        frame.features = [f for f in frame.features if f.is_inlier]

    # -- state machine -----------------------------------------------------------------------------------------
    def check_motion_estimation(self):
        if self.num_inliers < 10:
            return False
        gap = self.frame_current.frame_id - self.frame_last.frame_id
        return not (np.linalg.norm(self.T_c_l.log()) > 5.0 * gap)

    def insert_key_frame(self, check, pts, desc):
        if (self.num_inliers >= 80 and self.T_c_l.angleY() < 0.03) or not check:
            return False
        fc, mp = self.frame_current, self.map
        fc.is_keyframe, fc.keyframe_id = True, self.curr_keyframe_id
        for f in fc.features:
            lm = mp.landmarks[f.landmark_id]
            lm.observed_times += 1
            lm.observations.append((fc.keyframe_id, f.feature_id))
        self.disparity_map(fc)
        pts_3d, pts, desc, reliable = self.set_ref_3d_position(pts, desc, fc)
        tracked = {}
        for j, f in enumerate(fc.features):
            tracked.setdefault((float(f.pt[0]), float(f.pt[1])), []).append(j)
        feature_id = len(fc.features)
        for i, p in enumerate(pts):
            key = (float(p[0]), float(p[1]))
            hits = tracked.get(key, [])
            for j in hits:
                lm = mp.landmarks[fc.features[j].landmark_id]
                if not lm.reliable_depth and reliable[i]:
                    lm.pt_3d, lm.reliable_depth = pts_3d[i], True
            if not hits:
                f = Feature(feature_id, fc.frame_id, p, desc[i])
                f.landmark_id = self.curr_landmark_id
                fc.features.append(f)
                tracked.setdefault(key, []).append(len(fc.features) - 1)
                mp.insert_landmark(Landmark(self.curr_landmark_id, pts_3d[i], desc[i], reliable[i], (fc.keyframe_id, feature_id)))
                self.curr_landmark_id += 1
                feature_id += 1
        self.curr_keyframe_id += 1
        mp.insert_keyframe(fc)
        return True

    def initialization(self):
        fl = self.frame_last = Frame()
        fl.left, fl.right = self.read_img(0)
        fl.frame_id = 0
        pts, desc = self.feature_detection(fl.left)
        self.stage_log.append(dict(frame=0, n_detected=len(pts)))
        self.disparity_map(fl)
        pts_3d, pts, desc, reliable = self.set_ref_3d_position(pts, desc, fl)
        for i, p in enumerate(pts):
            f = Feature(i, 0, p, desc[i])
            f.landmark_id = self.curr_landmark_id
            fl.features.append(f)
            self.map.insert_landmark(Landmark(self.curr_landmark_id, pts_3d[i], desc[i], reliable[i], (0, i)))
            self.curr_landmark_id += 1
        fl.T_c_w, fl.is_keyframe, fl.keyframe_id = SE3(), True, self.curr_keyframe_id
        self.curr_keyframe_id += 1
        self.map.insert_keyframe(fl)
        return True

    def tracking(self):
        fc = self.frame_current = Frame()
        if self.frame_last.is_keyframe:
            self.frame_last = self.map.keyframes[self.frame_last.keyframe_id].copy()
        fl = self.frame_last
        fc.left, fc.right = self.read_img(self.seq)
        fc.frame_id = self.seq
        pts, desc = self.feature_detection(fc.left)
        d_last = np.array([f.desc for f in fl.features], dtype=np.uint8).reshape(-1, 32)
        matches = self.feature_matching(d_last, desc)
        self.stage_log.append(dict(frame=self.seq, n_detected=len(pts), matches=list(matches)))
        for i, (q, t) in enumerate(matches):
            f = Feature(i, self.seq, pts[t], desc[t])
            f.landmark_id = fl.features[q].landmark_id
            fc.features.append(f)
        self.motion_estimation(fc)
        fc.T_c_w = SE3(self.T_c_w.R, self.T_c_w.t)
        self.T_c_l = fc.T_c_w * fl.T_c_w.inverse()
        check = self.check_motion_estimation()
        kf = self.insert_key_frame(check, pts, desc)
        if check:
            self.frame_last = fc.copy()
        self.seq += 1
        return check, kf

    def pipeline(self):
        """VO::pipeline: returns (not_lost, if_insert_keyframe)"""
        kf = False
        if self.state == "Init":
            if self.initialization():
                self.state = "Track"
            else:
                self.num_lost += 1
                if self.num_lost > 10:
                    self.state = "Lost"
        elif self.state == "Track":
            ok, kf = self.tracking()
            if ok:
                self.num_lost = 0
            else:
                self.num_lost += 1
                if self.num_lost > 10:
                    self.state = "Lost"
        else:
            return False, False
        return True, kf

    # -- run_vslam.cpp:39-84 -----------------------------------------------------------------------------------
    def step(self):
        first = self.state == "Init"
        not_lost, kf = self.pipeline()
        if kf and self.do_ba and len(self.map.keyframes) >= self.map.num_keyframes:
            optimize(self.map, self.K, False, False, False, 5)
            optimize(self.map, self.K, False, False, False, 5)
            optimize(self.map, self.K, False, True, False, 10)
            optimize(self.map, self.K, True, True, False, 10)
        f = self.frame_last if first else self.frame_current
        rec = dict(frame_id=f.frame_id, T_w_c=f.T_c_w.inverse().matrix34(), num_inliers=self.num_inliers,
                   is_keyframe=bool(kf), n_keyframes=len(self.map.keyframes), n_landmarks=len(self.map.landmarks),
                   not_lost=not_lost)
        self.log.append(rec)
        return rec


def write_remaining_pose(mp: Map):
    """Map::write_remaining_pose (map.cpp:198-204): the keyframes still in the window, container order"""
    for kf in mp.keyframes.values():
        mp.written.append((kf.frame_id, kf.T_c_w.inverse().matrix34()))


def optimize(mp: Map, K, pose_only, if_update_map, if_update_landmark, num_ite):
    """optimize_map (optimization.cpp:103-288) / optimize_pose_only (:290-436): the reference's graph construction
    (inlier -- and for the full BA reliable-depth -- landmarks, every stored observation an edge, edges of one landmark
    contiguous), LM + relabel through oracle/ba_oracle.c, the landmark's last edge decides its inlier flag."""
    from . import ba_oracle
    kf_ids = list(mp.keyframes.keys())
    if not kf_ids:
        return
    pidx = {k: i for i, k in enumerate(kf_ids)}
    poses = np.array([mp.keyframes[k].T_c_w.matrix34().reshape(12) for k in kf_ids])
    lm_ids, pts, op, ol, uv = [], [], [], [], []
    for lid, lm in mp.landmarks.items():
        if not lm.is_inlier or (not pose_only and not lm.reliable_depth):
            continue
        for (kid, fid) in lm.observations:
            feat = mp.keyframes[kid].features[fid]
            if not lm_ids or lm_ids[-1] != lid:
                lm_ids.append(lid)
                pts.append(lm.pt_3d.astype(np.float64))
            op.append(pidx[kid]); ol.append(len(lm_ids) - 1); uv.append((float(feat.pt[0]), float(feat.pt[1])))
    r = ba_oracle.optimize(poses, np.array(pts).reshape(-1, 3), np.array(op, np.int32), np.array(ol, np.int32),
                           np.array(uv).reshape(-1, 2), K, num_iterations=num_ite, pose_only=pose_only)
    for i, lid in enumerate(lm_ids):
        mp.landmarks[lid].is_inlier = bool(r["point_inlier"][i])
    if if_update_map:
        for i, k in enumerate(kf_ids):
            T = r["poses"][i].reshape(3, 4)
            mp.keyframes[k].T_c_w = SE3(T[:, :3], T[:, 3])
        if if_update_landmark and not pose_only:
            for i, lid in enumerate(lm_ids):
                mp.landmarks[lid].pt_3d = r["points"][i].astype(np.float32)
