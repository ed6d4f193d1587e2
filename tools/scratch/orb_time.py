import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, vslam_b200_loader
pkg = vslam_b200_loader.pkg
ctx = pkg.Context(max_images=2, max_keypoints=8192, max_width=1920, max_height=1200)
L, R, _ = pkg.synth.synth_pair(3)
for _ in range(3): ctx.orb_detect_compute(L[None], nfeatures=3000, anms_keep=500)
ctx.timing_enable(True)
t0 = time.perf_counter()
for _ in range(20): out = ctx.orb_detect_compute(L[None], nfeatures=3000, anms_keep=500); n = len(out[0][0])
e2e = (time.perf_counter() - t0) / 20 * 1e3
kt = ctx.timing_read()
print("e2e ms %.3f n=%s" % (e2e, n), {k: round(v[0] / 20, 4) for k, v in kt.items()}, "sum", round(sum(v[0] for v in kt.values()) / 20, 4))
