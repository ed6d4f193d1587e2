"""Seeded 3D-2D correspondence sets for the PnP tests (CPU and GPU share them)."""
import numpy as np

# (seed, n, outlier fraction, pixel noise sigma)
CASES = [(0, 400, 0.2, 0.3), (1, 150, 0.3, 0.3), (2, 1500, 0.1, 0.3), (3, 60, 0.0, 0.3), (4, 500, 0.5, 1.0),
         (5, 30, 0.4, 0.5), (6, 8, 0.0, 0.2), (7, 6, 0.0, 0.1), (8, 500, 0.25, 0.0), (9, 2000, 0.6, 0.8),
         (10, 12, 0.5, 0.3), (11, 300, 0.05, 2.0)]


def scene(pkg, seed, n, outlier_frac, noise=0.3):
    rng = np.random.default_rng(seed)
    K = pkg.synth.kitti_K()
    R, t = pkg.synth.se3_exp(np.array([0.4, -0.05, 0.9, 0.01, 0.03, -0.005]))
    pw = np.stack([rng.uniform(-15, 15, n), rng.uniform(-3, 3, n), rng.uniform(6, 45, n)], 1).astype(np.float32)
    pc = pw.astype(np.float64) @ R.T + t
    uv = (pc[:, :2] / pc[:, 2:3]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]]
    uv += rng.normal(0, noise, uv.shape)
    no = int(n * outlier_frac)
    bad = rng.permutation(n)[:no]
    uv[bad] += rng.uniform(30, 200, (no, 2)) * rng.choice([-1, 1], (no, 2))   # gross outliers, far beyond 4 px
    return pw, uv.astype(np.float32), K, R, t, np.setdiff1d(np.arange(n), bad)


def garbage(seed, n=50):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-5, 5, (n, 3)).astype(np.float32) + np.float32([0, 0, 20]),
            rng.uniform(0, 1200, (n, 2)).astype(np.float32))
