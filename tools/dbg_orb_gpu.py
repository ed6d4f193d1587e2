import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vslam_b200_loader
pkg = vslam_b200_loader.pkg
from oracle import orb_restate as R
from oracle.orb_pattern import load_pattern
ctx = pkg.Context(max_images=2, max_keypoints=8192)
L, _, _ = pkg.synth.synth_pair(0)
kp, desc = ctx.orb_detect_compute(L, 2000, 0)
ref, rdesc = R.orb_detect_and_compute(L, 2000, load_pattern())
print("counts gpu", np.bincount(kp["octave"], minlength=8).tolist(), "ref", np.bincount(ref["octave"], minlength=8).tolist(), "quota", R.level_quotas(2000))
for l in range(8):
    a = kp[kp["octave"] == l]; m = ref["octave"] == l
    ra = ref["response"][m]; rp = ref["pt"][m]
    n = min(len(a), len(ra))
    same = (a["response"][:n] == ra[:n]) & (a["x"][:n] == rp[:n, 0]) & (a["y"][:n] == rp[:n, 1])
    bad = np.nonzero(~same)[0]
    print("level", l, len(a), len(ra), "first bad", bad[:5].tolist())
    if len(bad):
        i = bad[0]
        print("  gpu", a[i], " ref", ra[i], rp[i])
        # is the gpu set a subset of the oracle's harris candidates?
    sa = set(zip(a["x"].tolist(), a["y"].tolist())); sr = set(zip(rp[:,0].tolist(), rp[:,1].tolist()))
    print("  only gpu", len(sa - sr), "only ref", len(sr - sa))
    ang_bad = 0
    if len(a) == len(ra) and not len(bad):
        ang_bad = (a["angle"] != ref["angle"][m]).sum()
        dbad = (desc[kp["octave"] == l] != rdesc[m]).any(axis=1).sum()
        print("  angle mismatches", ang_bad, "desc mismatches", dbad)
