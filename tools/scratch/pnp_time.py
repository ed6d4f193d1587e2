import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, vslam_b200_loader
pkg = vslam_b200_loader.pkg
ctx = pkg.Context(max_images=2, max_keypoints=2048)
rng = np.random.default_rng(0); n = 400
K = pkg.synth.kitti_K()
R, t = pkg.synth.se3_exp(np.array([0.4, -0.05, 0.9, 0.01, 0.03, -0.005]))
pw = np.stack([rng.uniform(-15, 15, n), rng.uniform(-3, 3, n), rng.uniform(6, 45, n)], 1).astype(np.float32)
pc = pw.astype(np.float64) @ R.T + t
uv = ((pc[:, :2] / pc[:, 2:3]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]] + rng.normal(0, 0.3, (n, 2))).astype(np.float32)
for _ in range(3): ctx.pnp_ransac(pw, uv, K, 100, 4.0, 0.99)
ctx.timing_enable(True)
t0 = time.perf_counter()
for _ in range(20): g = ctx.pnp_ransac(pw, uv, K, 100, 4.0, 0.99)
e2e = (time.perf_counter() - t0) / 20 * 1e3
kt = ctx.timing_read()
print("e2e ms", e2e, "kernels", {k: (v[0] / v[1], v[1]) for k, v in kt.items()})
