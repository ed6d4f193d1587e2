"""ctypes binding of include/vslam_b200.h -- the only way Python reaches the CUDA path.

There is no fallback: if libvslam_b200.so is missing or no B200 is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvslam_b200.so")

VSLAM_OK = 0
STATUS = {0: "VSLAM_OK", -1: "VSLAM_E_INVALID", -2: "VSLAM_E_CAPACITY", -3: "VSLAM_E_CUDA",
          -4: "VSLAM_E_NODEVICE", -5: "VSLAM_E_NUMERIC", -6: "VSLAM_E_OVERFLOW"}


class VslamError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__(f"{where}: {STATUS.get(status, status)} {detail}".strip())


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_images", C.c_int32), ("max_width", C.c_int32),
                ("max_height", C.c_int32), ("max_keypoints", C.c_int32), ("max_ba_poses", C.c_int32),
                ("max_ba_points", C.c_int32), ("max_ba_obs", C.c_int32)]


class BaOptions(C.Structure):
    _fields_ = [("huber_delta", C.c_double), ("chi2_threshold", C.c_double), ("num_iterations", C.c_int32),
                ("pose_only", C.c_int32), ("max_trials", C.c_int32), ("reserved", C.c_int32), ("tau", C.c_double)]


class BaResult(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("trials", C.c_int32), ("accepted", C.c_int32), ("reserved", C.c_int32),
                ("chi2_initial", C.c_double), ("chi2_final", C.c_double), ("lambda_final", C.c_double),
                ("chi2_threshold", C.c_double), ("n_inlier_obs", C.c_int32), ("n_outlier_obs", C.c_int32)]


class SgbmParams(C.Structure):
    """vslam_sgbm_params; defaults = cv::StereoSGBM::create(0, 96, 9, 8*9*9, 32*9*9, 1, 63, 10, 100, 32)
    (visual_odometry.cpp:163-164)."""
    _fields_ = [(n, C.c_int32) for n in ("min_disparity", "num_disparities", "block_size", "P1", "P2",
                                         "disp12_max_diff", "pre_filter_cap", "uniqueness_ratio",
                                         "speckle_window_size", "speckle_range")]


KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
assert KEYPOINT_DTYPE.itemsize == 28 and DMATCH_DTYPE.itemsize == 16

_lib = None

_vp, _i, _d, _f, _i64 = C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_int64
_pi = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol include/vslam_b200.h declares
SIGNATURES = {
    "vslam_abi_version": (_i, []),
    "vslam_status_string": (C.c_char_p, [_i]),
    "vslam_last_error": (C.c_char_p, [_vp]),
    "vslam_ctx_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "vslam_ctx_destroy": (None, [_vp]),
    "vslam_ctx_set_stream": (_i, [_vp, _vp]),
    "vslam_host_alloc": (_vp, [C.c_size_t]),
    "vslam_host_free": (None, [_vp]),
    "vslam_ctx_synchronize": (_i, [_vp]),
    "vslam_ctx_set_concurrency": (_i, [_vp, _i]),
    "vslam_ctx_launch_count": (_i64, [_vp]),
    "vslam_kernel_count": (_i, []),
    "vslam_kernel_name": (C.c_char_p, [_i]),
    "vslam_ctx_timing_enable": (_i, [_vp, _i]),
    "vslam_ctx_timing_read": (_i, [_vp, _i, C.POINTER(_d), C.POINTER(_i64)]),
    "vslam_match_hamming": (_i, [_vp, _vp, _i, _vp, _i, _i, _d, _d, _vp, _pi]),
    "vslam_orb_keypoint_capacity": (_i, [_vp]),
    "vslam_orb_detect_compute": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "vslam_orb_detect_compute_batch": (_i, [_vp, _vp, _i, _i, _i, _i, C.c_longlong, _i, _i, _f, _vp, _vp, _vp]),
    "vslam_orb_detect_compute_batch_dev": (_i, [_vp, _vp, _i, _i, _i, _i, C.c_longlong, _i, _i, _f, _vp, _vp, _vp]),
    "vslam_orb_last_flags": (_i, [_vp, _i]),
    "vslam_orb_debug_read": (_i, [_vp, _i, _i, _vp, _vp, _pi, _pi, _vp, _i, _pi]),
    "vslam_triangulate": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "vslam_triangulate_matches_batch_dev": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "vslam_stereo_frontend_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_longlong, _i, _i, _f, _d, _d, _vp, _vp, _vp,
                                         _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vslam_stereo_frontend_batch_begin": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_longlong, _i, _i, _f, _d, _d, _vp, _vp, _vp,
                                               _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vslam_stereo_frontend_batch_end": (_i, [_vp]),
    "vslam_stereo_frontend_batch_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_longlong, _i, _i, _f, _d, _d, _vp, _vp,
                                             _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vslam_ba_optimize": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, C.POINTER(BaOptions),
                               C.POINTER(BaResult), _vp, _vp]),
    "vslam_ba_optimize_multi": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, C.POINTER(BaOptions),
                                     C.POINTER(BaResult), _vp, _vp]),
    "vslam_pnp_ransac": (_i, [_vp, _vp, _vp, _i, _vp, _i, _f, _d, _vp, _vp, _vp, _vp, _pi]),
    "vslam_pnp_debug_read": (_i, [_vp, _i, _vp, _pi, _pi]),
    "vslam_anms": (_i, [_vp, _vp, _i, _i, _f, _vp, _pi]),
    "vslam_ba_last_phase_ns": (_i, [_vp, _vp]),
    "vslam_ba_multi_last_profile_ns": (_i, [_vp, _vp]),
    "vslam_ba_reduce_sizes": (_i, [_i, _pi, _pi, _pi]),
    "vslam_ba_session_begin": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, C.POINTER(BaOptions), _i, _i, _vp,
                                    _vp, _vp]),
    "vslam_ba_session_phase": (_i, [_vp, _i, _d]),
    "vslam_ba_session_trial_done": (_i, [_vp, _i]),
    "vslam_ba_session_end": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "vslam_ba_session_schur_dense": (_i, [_vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "vslam_sgbm_default_params": (None, [C.POINTER(SgbmParams)]),
    "vslam_sgbm_compute": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_longlong, C.POINTER(SgbmParams), _vp, _vp]),
    "vslam_sgbm_compute_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, C.c_longlong, C.POINTER(SgbmParams), _vp, _vp]),
    "vslam_sgbm_debug_read": (_i, [_vp, _i, _i, _vp, C.c_size_t]),
    "vslam_sgbm_debug_stop_after": (_i, [_vp, _i]),
    "vslam_match_hamming_batch_dev": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _d, _d, _vp, _i, _vp]),
}


def load_library() -> C.CDLL:
    """dlopen the in-tree library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    """host numpy array, torch CUDA tensor, or int -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())  # torch tensor


class Context:
    """Owns a vslam_ctx (device memory, stream).  Mirrors the role of the reference's VO object's
    cv::Ptr<ORB>/cv::Ptr<BFMatcher> members (visual_odometry.cpp:20-35): created once, reused per frame."""

    def __init__(self, device: int = 0, max_images: int = 2, max_width: int = 1241, max_height: int = 376,
                 max_keypoints: int = 4096, max_ba_poses: int = 64, max_ba_points: int = 32768,
                 max_ba_obs: int = 262144):
        self.lib = load_library()
        self.cfg = Config(device, max_images, max_width, max_height, max_keypoints, max_ba_poses,
                          max_ba_points, max_ba_obs)
        h = C.c_void_p()
        st = self.lib.vslam_ctx_create(C.byref(self.cfg), C.byref(h))
        if st != VSLAM_OK:
            raise VslamError(st, "vslam_ctx_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.vslam_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    def check(self, st: int, where: str):
        if st != VSLAM_OK:
            raise VslamError(st, where, self.lib.vslam_last_error(self.h).decode())

    def set_stream(self, cuda_stream_handle: int | None):
        self.check(self.lib.vslam_ctx_set_stream(self.h, C.c_void_p(cuda_stream_handle or 0)), "set_stream")

    def set_concurrency(self, on: bool):
        """on=False keeps every launch on the context stream (kernels timed in isolation)"""
        self.check(self.lib.vslam_ctx_set_concurrency(self.h, int(on)), "vslam_ctx_set_concurrency")

    def synchronize(self):
        self.check(self.lib.vslam_ctx_synchronize(self.h), "synchronize")

    @property
    def launch_count(self) -> int:
        return int(self.lib.vslam_ctx_launch_count(self.h))

    def timing_enable(self, on: bool = True):
        self.check(self.lib.vslam_ctx_timing_enable(self.h, int(on)), "timing_enable")

    def timing_read(self) -> dict:
        """{kernel name: (total device ms, launches)} since timing_enable (CUDA events on the launching stream)."""
        out = {}
        for i in range(self.lib.vslam_kernel_count()):
            ms, n = C.c_double(0), C.c_int64(0)
            self.check(self.lib.vslam_ctx_timing_read(self.h, i, C.byref(ms), C.byref(n)), "timing_read")
            if n.value:
                out[self.lib.vslam_kernel_name(i).decode()] = (ms.value, n.value)
        return out

    # ---- K10 --------------------------------------------------------------------------------
    def match_hamming(self, query: np.ndarray, train: np.ndarray, cross_check: bool = True,
                      gate_rel: float = -1.0, gate_abs: float = 0.0) -> np.ndarray:
        """cv::BFMatcher(NORM_HAMMING, crossCheck)::match + optional VO::feature_matching gate.
        Returns a DMATCH_DTYPE array ordered by queryIdx."""
        q = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, dtype=np.uint8).reshape(-1, 32)
        out = np.zeros(max(len(q), 1), dtype=DMATCH_DTYPE)
        n = C.c_int(0)
        st = self.lib.vslam_match_hamming(self.h, _ptr(q), len(q), _ptr(t), len(t), int(cross_check),
                                          float(gate_rel), float(gate_abs), _ptr(out), C.byref(n))
        self.check(st, "vslam_match_hamming")
        return out[:n.value].copy()

    def match_hamming_batch_dev(self, d_query, d_nq, q_stride_rows, d_train, d_nt, t_stride_rows, batch, max_rows,
                                cross_check, gate_rel, gate_abs, d_out, out_stride, d_n_out):
        st = self.lib.vslam_match_hamming_batch_dev(self.h, _ptr(d_query), _ptr(d_nq), q_stride_rows, _ptr(d_train),
                                                    _ptr(d_nt), t_stride_rows, batch, max_rows, int(cross_check),
                                                    float(gate_rel), float(gate_abs), _ptr(d_out), out_stride,
                                                    _ptr(d_n_out))
        self.check(st, "vslam_match_hamming_batch_dev")

    # ---- K1-K9 ------------------------------------------------------------------------------
    @property
    def kp_cap(self) -> int:
        return int(self.lib.vslam_orb_keypoint_capacity(self.h))

    def orb_detect_compute(self, images, nfeatures: int = 3000, anms_keep: int = 500, anms_c: float = 1.11):
        """VO::feature_detection on one image (H x W u8) or a batch (B x H x W u8), host buffers.
        Returns a list of (keypoints KEYPOINT_DTYPE[n], descriptors u8[n,32]) -- one per image."""
        if images is None:
            raise VslamError(-1, "vslam_orb_detect_compute", "Could not open or find the image")
        imgs = np.ascontiguousarray(images, dtype=np.uint8)
        single = imgs.ndim == 2
        if single:
            imgs = imgs[None]
        b, h, w = imgs.shape
        cap = self.kp_cap
        kp = np.zeros((b, cap), dtype=KEYPOINT_DTYPE)
        desc = np.zeros((b, cap, 32), dtype=np.uint8)
        n = np.zeros(b, dtype=np.int32)
        st = self.lib.vslam_orb_detect_compute_batch(self.h, _ptr(imgs), b, w, h, w, h * w, int(nfeatures),
                                                     int(anms_keep), float(anms_c), _ptr(kp), _ptr(desc), _ptr(n))
        self.check(st, "vslam_orb_detect_compute_batch")
        out = [(kp[i, :n[i]].copy(), desc[i, :n[i]].copy()) for i in range(b)]
        return out[0] if single else out

    def orb_detect_compute_batch_dev(self, d_images, n_images, w, h, pitch, img_stride, nfeatures, anms_keep, anms_c,
                                     d_kp, d_desc, d_n):
        st = self.lib.vslam_orb_detect_compute_batch_dev(self.h, _ptr(d_images), n_images, w, h, pitch, img_stride,
                                                         int(nfeatures), int(anms_keep), float(anms_c), _ptr(d_kp),
                                                         _ptr(d_desc), _ptr(d_n))
        self.check(st, "vslam_orb_detect_compute_batch_dev")

    def orb_last_flags(self, n_images: int):
        self.check(self.lib.vslam_orb_last_flags(self.h, n_images), "vslam_orb_last_flags")

    def orb_debug_level(self, img: int, level: int, want_cand: bool = True):
        """(level pixels or None for level 0, blurred pixels, candidates[n,2] u32) of the last ORB call."""
        w, hh, nc = C.c_int(0), C.c_int(0), C.c_int(0)
        self.check(self.lib.vslam_orb_debug_read(self.h, img, level, None, None, C.byref(w), C.byref(hh), None, 0,
                                                 C.byref(nc)), "vslam_orb_debug_read")
        lv = np.zeros((hh.value, w.value), dtype=np.uint8)
        bl = np.zeros((hh.value, w.value), dtype=np.uint8)
        cand = np.zeros((max(nc.value, 1), 2), dtype=np.uint32)
        self.check(self.lib.vslam_orb_debug_read(self.h, img, level, _ptr(lv), _ptr(bl), C.byref(w), C.byref(hh),
                                                 _ptr(cand), nc.value, C.byref(nc)), "vslam_orb_debug_read")
        return (lv if level > 0 else None), bl, cand[:nc.value]

    # ---- K11 + frontend ---------------------------------------------------------------------
    def triangulate(self, xl, xr, P1, P2, T_c_w=None):
        """(xyz_world float32 [n,3], flags u8 [n]) -- see vslam_triangulate."""
        xl = np.ascontiguousarray(xl, dtype=np.float32).reshape(-1, 2)
        xr = np.ascontiguousarray(xr, dtype=np.float32).reshape(-1, 2)
        P1 = np.ascontiguousarray(P1, dtype=np.float64).reshape(12)
        P2 = np.ascontiguousarray(P2, dtype=np.float64).reshape(12)
        T = None if T_c_w is None else np.ascontiguousarray(T_c_w, dtype=np.float64).reshape(12)
        n = len(xl)
        xyz = np.zeros((max(n, 1), 3), dtype=np.float32)
        fl = np.zeros(max(n, 1), dtype=np.uint8)
        st = self.lib.vslam_triangulate(self.h, _ptr(xl), _ptr(xr), n, _ptr(P1), _ptr(P2), _ptr(T), _ptr(xyz), _ptr(fl))
        self.check(st, "vslam_triangulate")
        return xyz[:n], fl[:n]

    def stereo_frontend(self, left, right, P1, P2, T_c_w=None, nfeatures=2000, anms_keep=0, anms_c=1.11,
                        gate_rel=2.0, gate_abs=30.0, out=None, begin_only=False):
        """Host-buffer batched frontend.  left/right: [B,H,W] u8 numpy arrays or pinned torch tensors.
        Returns a dict of full-stride numpy arrays (see include/vslam_b200.h) plus per-pair views."""
        is_np = isinstance(left, np.ndarray)
        if is_np:
            left = np.ascontiguousarray(left, dtype=np.uint8)
            right = np.ascontiguousarray(right, dtype=np.uint8)
        b, h, w = left.shape
        cap = self.kp_cap
        if out is None:
            out = self.alloc_frontend_outputs(b)
        P1 = np.ascontiguousarray(P1, dtype=np.float64).reshape(12)
        P2 = np.ascontiguousarray(P2, dtype=np.float64).reshape(12)
        T = None if T_c_w is None else np.ascontiguousarray(T_c_w, dtype=np.float64).reshape(b, 12)
        fn = self.lib.vslam_stereo_frontend_batch_begin if begin_only else self.lib.vslam_stereo_frontend_batch
        st = fn(
            self.h, _ptr(left), _ptr(right), b, w, h, w, h * w, int(nfeatures), int(anms_keep), float(anms_c),
            float(gate_rel), float(gate_abs), _ptr(P1), _ptr(P2), _ptr(T), _ptr(out["kp"]), _ptr(out["desc"]),
            _ptr(out["n_kp"]), _ptr(out["matches"]), _ptr(out["n_matches"]), _ptr(out["xyz"]), _ptr(out["flags"]))
        self.check(st, "vslam_stereo_frontend_batch")
        if begin_only:  # the arrays the enqueued copies read must outlive the call
            self._front_keep = (left, right, P1, P2, T, out)
        return out

    def stereo_frontend_begin(self, left, right, P1, P2, out, **kw):
        """Enqueue a host-buffer batch and return (vslam_stereo_frontend_batch_begin); `out` (alloc_frontend_outputs,
        ideally pinned) is filled when stereo_frontend_end() returns.  left/right should be pinned."""
        return self.stereo_frontend(left, right, P1, P2, out=out, begin_only=True, **kw)

    def stereo_frontend_end(self):
        self.check(self.lib.vslam_stereo_frontend_batch_end(self.h), "vslam_stereo_frontend_batch_end")
        self._front_keep = None

    def alloc_frontend_outputs(self, n_pairs: int):
        cap = self.kp_cap
        return dict(kp=np.zeros((2 * n_pairs, cap), dtype=KEYPOINT_DTYPE),
                    desc=np.zeros((2 * n_pairs, cap, 32), dtype=np.uint8),
                    n_kp=np.zeros(2 * n_pairs, dtype=np.int32),
                    matches=np.zeros((n_pairs, cap), dtype=DMATCH_DTYPE),
                    n_matches=np.zeros(n_pairs, dtype=np.int32),
                    xyz=np.zeros((n_pairs, cap, 3), dtype=np.float32),
                    flags=np.zeros((n_pairs, cap), dtype=np.uint8))

    def stereo_frontend_dev(self, d_left, d_right, n_pairs, w, h, pitch, img_stride, P1, P2, d_T, d_kp, d_desc, d_n_kp,
                            d_matches, d_n_matches, d_xyz, d_flags, nfeatures=2000, anms_keep=0, anms_c=1.11,
                            gate_rel=2.0, gate_abs=30.0):
        P1 = np.ascontiguousarray(P1, dtype=np.float64).reshape(12)
        P2 = np.ascontiguousarray(P2, dtype=np.float64).reshape(12)
        st = self.lib.vslam_stereo_frontend_batch_dev(
            self.h, _ptr(d_left), _ptr(d_right), n_pairs, w, h, pitch, img_stride, int(nfeatures), int(anms_keep),
            float(anms_c), float(gate_rel), float(gate_abs), _ptr(P1), _ptr(P2), _ptr(d_T), _ptr(d_kp), _ptr(d_desc),
            _ptr(d_n_kp), _ptr(d_matches), _ptr(d_n_matches), _ptr(d_xyz), _ptr(d_flags))
        self.check(st, "vslam_stereo_frontend_batch_dev")

    # ---- K13-K16 ----------------------------------------------------------------------------
    def ba_optimize(self, poses, points, obs_pose, obs_point, obs_uv, K, num_iterations=10, pose_only=False,
                    huber_delta=5.991, chi2_th=5.991, max_trials=10, tau=1e-5, point_inlier=None):
        """optimize_map / optimize_pose_only arithmetic (inputs are not modified).  Returns a dict."""
        poses = np.array(poses, dtype=np.float64, order="C").reshape(-1, 12).copy()
        points = np.array(points, dtype=np.float64, order="C").reshape(-1, 3).copy()
        op = np.ascontiguousarray(obs_pose, dtype=np.int32)
        ol = np.ascontiguousarray(obs_point, dtype=np.int32)
        uv = np.ascontiguousarray(obs_uv, dtype=np.float64).reshape(-1, 2)
        Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        opt = BaOptions(huber_delta, chi2_th, int(num_iterations), int(pose_only), int(max_trials), 0, tau)
        res = BaResult()
        chi2 = np.zeros(max(len(op), 1), dtype=np.float64)
        inl = (np.ones(max(len(points), 1), dtype=np.uint8) if point_inlier is None
               else np.ascontiguousarray(point_inlier, dtype=np.uint8).copy())
        st = self.lib.vslam_ba_optimize(self.h, len(poses), _ptr(poses), len(points), _ptr(points), len(op), _ptr(op),
                                        _ptr(ol), _ptr(uv), _ptr(Kc), C.byref(opt), C.byref(res), _ptr(chi2), _ptr(inl))
        self.check(st, "vslam_ba_optimize")
        return dict(poses=poses, points=points, chi2_per_obs=chi2[:len(op)], point_inlier=inl[:len(points)].astype(bool),
                    iterations=res.iterations, trials=res.trials, accepted=res.accepted,
                    chi2_initial=res.chi2_initial, chi2_final=res.chi2_final, lambda_final=res.lambda_final,
                    chi2_threshold=res.chi2_threshold, n_inlier_obs=res.n_inlier_obs, n_outlier_obs=res.n_outlier_obs)

    @staticmethod
    def ba_optimize_multi(ctxs, poses, points, obs_pose, obs_point, obs_uv, K, num_iterations=10, pose_only=False,
                          huber_delta=5.991, chi2_th=5.991, max_trials=10, tau=1e-5, point_inlier=None):
        """ONE window on len(ctxs) GPUs of this process (vslam_ba_optimize_multi): device-side LM loop, the ranks
        exchange the reduced camera system through peer memory.  Same dict as ba_optimize plus `exchanges`."""
        poses = np.array(poses, dtype=np.float64, order="C").reshape(-1, 12).copy()
        points = np.array(points, dtype=np.float64, order="C").reshape(-1, 3).copy()
        op = np.ascontiguousarray(obs_pose, dtype=np.int32)
        ol = np.ascontiguousarray(obs_point, dtype=np.int32)
        uv = np.ascontiguousarray(obs_uv, dtype=np.float64).reshape(-1, 2)
        Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        opt = BaOptions(huber_delta, chi2_th, int(num_iterations), int(pose_only), int(max_trials), 0, tau)
        res = BaResult()
        chi2 = np.zeros(max(len(op), 1), dtype=np.float64)
        inl = (np.ones(max(len(points), 1), dtype=np.uint8) if point_inlier is None
               else np.ascontiguousarray(point_inlier, dtype=np.uint8).copy())
        handles = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        lib = ctxs[0].lib
        st = lib.vslam_ba_optimize_multi(handles, len(ctxs), len(poses), _ptr(poses), len(points), _ptr(points), len(op),
                                         _ptr(op), _ptr(ol), _ptr(uv), _ptr(Kc), C.byref(opt), C.byref(res), _ptr(chi2),
                                         _ptr(inl))
        ctxs[0].check(st, "vslam_ba_optimize_multi")
        return dict(poses=poses, points=points, chi2_per_obs=chi2[:len(op)], point_inlier=inl[:len(points)].astype(bool),
                    iterations=res.iterations, trials=res.trials, accepted=res.accepted, exchanges=res.reserved,
                    chi2_initial=res.chi2_initial, chi2_final=res.chi2_final, lambda_final=res.lambda_final,
                    chi2_threshold=res.chi2_threshold, n_inlier_obs=res.n_inlier_obs, n_outlier_obs=res.n_outlier_obs)

    # ---- K12 / K7 ---------------------------------------------------------------------------
    def pnp_ransac(self, xyz, uv, K, iters=100, reproj_err=4.0, confidence=0.99):
        """cv::solvePnPRansac arithmetic.  Returns dict(rvec, tvec, T_c_w (3x4), inliers int32[])."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        Kc = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        rvec, tvec, T = np.zeros(3), np.zeros(3), np.zeros(12)
        inl = np.zeros(max(len(xyz), 1), dtype=np.int32)
        n = C.c_int(0)
        st = self.lib.vslam_pnp_ransac(self.h, _ptr(xyz), _ptr(uv), len(xyz), _ptr(Kc), int(iters), float(reproj_err),
                                       float(confidence), _ptr(rvec), _ptr(tvec), _ptr(T), _ptr(inl), C.byref(n))
        self.check(st, "vslam_pnp_ransac")
        return dict(rvec=rvec, tvec=tvec, T_c_w=T.reshape(3, 4), inliers=inl[:n.value].copy())

    def pnp_debug(self, n_samples):
        """Test tap: per-sample models [n_samples, 12] (R row-major, t), inlier counts and the executed iterations of
        the last pnp_ransac call."""
        models = np.zeros((n_samples, 12))
        counts = np.zeros(n_samples, dtype=np.int32)
        ex = C.c_int(0)
        for i in range(n_samples):
            c = C.c_int(0)
            self.check(self.lib.vslam_pnp_debug_read(self.h, i, _ptr(models[i]), C.byref(c), C.byref(ex)), "vslam_pnp_debug_read")
            counts[i] = c.value
        return models, counts, ex.value

    def anms(self, keypoints, num=500, c_robust=1.11):
        kp = np.ascontiguousarray(keypoints, dtype=KEYPOINT_DTYPE)
        keep = np.zeros(max(len(kp), 1), dtype=np.int32)
        n = C.c_int(0)
        st = self.lib.vslam_anms(self.h, _ptr(kp), len(kp), int(num), float(c_robust), _ptr(keep), C.byref(n))
        self.check(st, "vslam_anms")
        return keep[:n.value].copy()

    # ---- dense stereo (VO::disparity_map, visual_odometry.cpp:159-174) -----------------------------------
    def sgbm_params(self, **kw) -> SgbmParams:
        p = SgbmParams()
        self.lib.vslam_sgbm_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, int(v))
        return p

    def sgbm_compute(self, left, right, params: SgbmParams | None = None, want_float: bool = False):
        """left/right: (H, W) or (n, H, W) uint8 host arrays -> int16 disparity*16 (and float32 disparity)."""
        left = np.ascontiguousarray(left, dtype=np.uint8)
        right = np.ascontiguousarray(right, dtype=np.uint8)
        single = left.ndim == 2
        if single:
            left, right = left[None], right[None]
        n, h, w = left.shape
        d16 = np.empty((n, h, w), np.int16)
        df = np.empty((n, h, w), np.float32) if want_float else None
        st = self.lib.vslam_sgbm_compute(self.h, _ptr(left), _ptr(right), n, w, h, w, w * h,
                                         C.byref(params) if params is not None else None, _ptr(d16), _ptr(df))
        self.check(st, "vslam_sgbm_compute")
        if single:
            d16, df = d16[0], (df[0] if df is not None else None)
        return (d16, df) if want_float else d16

    def sgbm_compute_dev(self, d_left, d_right, n_pairs, w, h, pitch, img_stride, d_disp16, d_disp_f32=None, params=None):
        st = self.lib.vslam_sgbm_compute_dev(self.h, _ptr(d_left), _ptr(d_right), n_pairs, w, h, pitch, img_stride,
                                             C.byref(params) if params is not None else None, _ptr(d_disp16),
                                             _ptr(d_disp_f32))
        self.check(st, "vslam_sgbm_compute_dev")

    def sgbm_debug_volume(self, pair: int, stage: int, h: int, w: int):
        """stage 0 = C, 1..3 = path volumes -> (h, w-96, 96) uint16; 4 / 5 = raw / median disparity (h, w) int16."""
        out = np.empty((h, w - 96, 96), np.uint16) if stage <= 3 else np.empty((h, w), np.int16)
        self.check(self.lib.vslam_sgbm_debug_read(self.h, pair, stage, _ptr(out), out.nbytes), "vslam_sgbm_debug_read")
        return out

    def sgbm_debug_stop_after(self, stage: int):
        self.check(self.lib.vslam_sgbm_debug_stop_after(self.h, stage), "vslam_sgbm_debug_stop_after")

    def ba_last_phase_us(self) -> dict:
        ns = np.zeros(8, dtype=np.uint64)
        self.check(self.lib.vslam_ba_last_phase_ns(self.h, _ptr(ns)), "vslam_ba_last_phase_ns")
        names = ["zero", "build", "schur_init", "schur", "schur_reduce", "solve", "update", "trial_err"]
        return {k: float(v) / 1e3 for k, v in zip(names, ns)}

    def ba_multi_last_profile_us(self) -> dict:
        ns = np.zeros(4, dtype=np.uint64)
        self.check(self.lib.vslam_ba_multi_last_profile_ns(self.h, _ptr(ns)), "vslam_ba_multi_last_profile_ns")
        out = {k: float(v) / 1e3 for k, v in zip(["start_to_first_exchange", "first_exchange_to_end", "system_exchanges",
                                                    "publish_and_wait"], ns)}
        ns8 = np.zeros(8, dtype=np.uint64)
        self.check(self.lib.vslam_ba_last_phase_ns(self.h, _ptr(ns8)), "vslam_ba_last_phase_ns")
        out.update(peer_store_loops=float(ns8[0]) / 1e3, barrier_after_stores=float(ns8[1]) / 1e3,
                   fence_and_flags=float(ns8[2]) / 1e3, sum_and_barrier=float(ns8[4]) / 1e3)
        return out

    # ---- K17: landmark-sharded BA session (driver: sharding.ba_optimize_sharded) ------------
    def ba_session(self, problem, shard, r1, r2, r3, **opt):
        return GpuBaSession(self, problem, shard, r1, r2, r3, **opt)


def ba_reduce_sizes(n_poses: int):
    """(r1, r2, r3) sizes in doubles of the three reduce buffers for a window of n_poses keyframes."""
    return 42 * n_poses + 2, 36 * n_poses * n_poses + 6 * n_poses, 10


class GpuBaSession:
    """One rank's side of the landmark-sharded BA: thin wrapper over vslam_ba_session_* (see include/vslam_b200.h)."""
    BUILD, IMPORT_BUILD, SCHUR, SOLVE_UPDATE, RELABEL_COUNT, RELABEL_APPLY = 1, 2, 3, 4, 5, 6

    def __init__(self, ctx, problem, shard, r1, r2, r3, num_iterations=10, pose_only=False, huber_delta=5.991,
                 chi2_th=5.991, max_trials=10, tau=1e-5):
        self.ctx = ctx
        # Stream contract (vslam_b200.h K17): the phases run on the context stream, the caller's collectives and
        # host reads on the stream that owns r1/r2/r3 -- they must be the SAME stream.  With torch reduce buffers the
        # session therefore moves the context onto torch's current stream of that device.
        if hasattr(r1, "is_cuda") and r1.is_cuda:
            import torch
            ctx.set_stream(torch.cuda.current_stream(r1.device).cuda_stream)
        self.poses = np.ascontiguousarray(problem["poses"], dtype=np.float64).reshape(-1, 12)
        self.points = np.ascontiguousarray(problem["points"], dtype=np.float64).reshape(-1, 3)
        self.op = np.ascontiguousarray(problem["obs_pose"], dtype=np.int32)
        self.ol = np.ascontiguousarray(problem["obs_point"], dtype=np.int32)
        self.uv = np.ascontiguousarray(problem["obs_uv"], dtype=np.float64).reshape(-1, 2)
        Kc = np.ascontiguousarray(problem["K"], dtype=np.float64).reshape(9)
        opt = BaOptions(huber_delta, chi2_th, int(num_iterations), int(pose_only), int(max_trials), 0, tau)
        st = ctx.lib.vslam_ba_session_begin(ctx.h, len(self.poses), _ptr(self.poses), len(self.points), _ptr(self.points),
                                            len(self.op), _ptr(self.op), _ptr(self.ol), _ptr(self.uv), _ptr(Kc),
                                            C.byref(opt), int(shard[0]), int(shard[1]), _ptr(r1), _ptr(r2), _ptr(r3))
        ctx.check(st, "vslam_ba_session_begin")

    def phase(self, phase: int, value: float = 0.0):
        self.ctx.check(self.ctx.lib.vslam_ba_session_phase(self.ctx.h, int(phase), float(value)), "vslam_ba_session_phase")

    def trial_done(self, accept: bool):
        self.ctx.check(self.ctx.lib.vslam_ba_session_trial_done(self.ctx.h, int(accept)), "vslam_ba_session_trial_done")

    def schur_dense(self, d_S):
        """N1 probe: dense DMMA SYRK form of the Schur product into the device tensor d_S (n x n); returns (ms_fill, ms_syrk)"""
        a, b = C.c_float(0), C.c_float(0)
        self.ctx.check(self.ctx.lib.vslam_ba_session_schur_dense(self.ctx.h, _ptr(d_S), C.byref(a), C.byref(b)),
                       "vslam_ba_session_schur_dense")
        return a.value, b.value

    def end(self):
        poses = np.zeros_like(self.poses); points = np.zeros_like(self.points)
        chi2 = np.zeros(len(self.op)); inl = np.zeros(len(self.points), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.vslam_ba_session_end(self.ctx.h, _ptr(poses), _ptr(points), _ptr(chi2), _ptr(inl)),
                       "vslam_ba_session_end")
        return poses, points, chi2, inl
