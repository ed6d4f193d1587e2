"""BA micro-benchmark with the device-side phase profile.  python tools/bench_ba.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vslam_b200_loader as L
pkg = L.pkg
ctx = pkg.Context(max_images=0, max_width=0, max_height=0, max_keypoints=1)
for name, (seed, nk, nl, nobs) in {"vo_window": (7, 10, 600, None), "cfg3": (42, 10, 5000, None), "cfg5": (43, 50, 20000, 100000)}.items():
    p = pkg.synth.synth_ba_problem(seed, nk, nl, n_obs_exact=nobs)
    a = (p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"])
    for _ in range(2):
        ctx.ba_optimize(*a, num_iterations=10)
    ctx.timing_enable(True)
    t0 = time.perf_counter()
    for _ in range(5):
        r = ctx.ba_optimize(*a, num_iterations=10)
    dt = (time.perf_counter() - t0) / 5
    k = ctx.timing_read()["ba_lm_kernel"]; ctx.timing_enable(False)
    ph = ctx.ba_last_phase_us()
    print(name, "kernel ms %.3f" % (k[0] / k[1]), "e2e ms %.3f" % (dt * 1e3), "iters", r["iterations"], "trials", r["trials"],
          "| phase us per trial:", {k2: round(v / r["trials"], 1) for k2, v in ph.items()})
