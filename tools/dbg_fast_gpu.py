import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vslam_b200_loader
pkg = vslam_b200_loader.pkg
from oracle import orb_restate as R
ctx = pkg.Context(max_images=2, max_keypoints=4096)
L, _, _ = pkg.synth.synth_pair(0)
ctx.orb_detect_compute(L, 2000, 0)
levels = R.build_pyramid(L)
for l in (0, 3):
    lv, bl, cand = ctx.orb_debug_level(0, l)
    xs, ys, sc = R.fast_nms(R.fast_score_map(levels[l]))
    h, w = levels[l].shape
    m = (xs >= 31) & (xs < w - 31) & (ys >= 31) & (ys < h - 31)
    ref = set(zip(ys[m].tolist(), xs[m].tolist(), sc[m].tolist()))
    mine = set(zip((cand[:, 0] >> 16).tolist(), (cand[:, 0] & 0xFFFF).tolist(), cand[:, 1].tolist()))
    miss = sorted(ref - mine); extra = sorted(mine - ref)
    print("level", l, "ref", len(ref), "mine", len(mine), "n_unique_mine", len(mine), "raw", len(cand))
    print(" missing", len(miss), miss[:12])
    print(" extra", len(extra), extra[:12])
    if miss:
        a = np.array(miss); print(" missing x%64 hist", np.bincount(a[:,1] % 64, minlength=64).tolist()); print(" missing y%32 hist", np.bincount(a[:,0] % 32, minlength=32).tolist())
    if extra:
        a = np.array(extra); print(" extra x%64 hist", np.bincount(a[:,1] % 64, minlength=64).tolist()); print(" extra y%32 hist", np.bincount(a[:,0] % 32, minlength=32).tolist())
