"""Small BA / PnP / ANMS / dense-Schur run for ncu:  python tools/profile_ba_pnp.py
ba_lm_kernel on configs[2] (K=10, L=5000) and configs[4] (K=50, L=20000, 100k obs), the PnP-RANSAC kernels on 400
correspondences, the stand-alone ANMS on 3000 keypoints and the dense DMMA Schur probe on configs[4]."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
import vslam_b200_loader  # noqa: E402

pkg = vslam_b200_loader.pkg
from pnp_scenes import scene  # noqa: E402

ctx = pkg.Context(device=0, max_images=2, max_keypoints=4096, max_ba_poses=64, max_ba_points=32768, max_ba_obs=262144)
for seed, nk, nl, nobs in ((42, 10, 5000, None), (43, 50, 20000, 100000)):
    p = pkg.synth.synth_ba_problem(seed, nk, nl, n_obs_exact=nobs)
    a = (p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"])
    for _ in range(2):
        r = ctx.ba_optimize(*a, num_iterations=10)
    print("ba", nk, nl, r["trials"], r["chi2_final"])
n = 6 * 50
n1, n2, n3 = pkg.ffi.ba_reduce_sizes(50)
q1, q2, q3 = (torch.zeros(k, dtype=torch.float64, device="cuda:0") for k in (n1, n2, n3))
Sd = torch.zeros(n * n, dtype=torch.float64, device="cuda:0")
torch.cuda.synchronize()
sess = ctx.ba_session(p, (0, 20000), q1, q2, q3, num_iterations=1)
sess.phase(sess.BUILD)
sess.phase(sess.SCHUR, 1e-3)
print("dense", sess.schur_dense(Sd))
sess.end()
pw, uv, K, *_ = scene(pkg, 0, 400, 0.2)
for _ in range(2):
    g = ctx.pnp_ransac(pw, uv, K)
print("pnp inliers", len(g["inliers"]))
rng = np.random.default_rng(4)
kp = np.zeros(3000, dtype=pkg.KEYPOINT_DTYPE)
kp["x"] = rng.uniform(0, 1241, 3000).astype(np.float32)
kp["y"] = rng.uniform(0, 376, 3000).astype(np.float32)
kp["response"] = rng.uniform(1e-6, 1e-2, 3000).astype(np.float32)
print("anms", len(ctx.anms(kp, 500)))
ctx.close()
