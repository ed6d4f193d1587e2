import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import vslam_b200_loader as L
pkg = L.pkg
ctx = pkg.Context(max_images=0, max_width=0, max_height=0, max_keypoints=1)
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 6
p = pkg.synth.synth_ba_problem(1, nk, 300)
r = ctx.ba_optimize(p["poses"], p["points"], p["obs_pose"], p["obs_point"], p["obs_uv"], p["K"], num_iterations=3)
print("ok", r["chi2_initial"], r["chi2_final"], r["trials"])
