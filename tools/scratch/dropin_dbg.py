import os, subprocess, sys, tempfile
import numpy as np
sys.path.insert(0, "/root/repo")
import vslam_b200_loader
pkg = vslam_b200_loader.pkg
d = tempfile.mkdtemp()
n = 16
lefts, rights, t, disp = pkg.synth.synth_sequence(3, n)
os.makedirs(d + "/image_0"); os.makedirs(d + "/image_1")
for i in range(n):
    pkg.synth.write_pgm(f"{d}/image_0/{i:06d}.pgm", lefts[i]); pkg.synth.write_pgm(f"{d}/image_1/{i:06d}.pgm", rights[i])
BIN = "/root/repo/stereo-visual-slam_b200/run_vslam"
for extra in (["--window", "12"], ["--window", "12", "--update-landmarks"], ["--update-landmarks"], []):
    w = tempfile.mkdtemp()
    r = subprocess.run([BIN, d + "/", str(n), "--nfeatures", "1000", "--anms", "110", *extra], cwd=w, capture_output=True, text=True)
    rows = [l.split() for l in r.stdout.splitlines() if l.startswith("frame ")]
    T = np.array([[float(v) for v in x[2:14]] for x in rows]).reshape(-1, 3, 4)
    meta = np.array([[int(v) for v in x[14:18]] for x in rows])
    err = np.abs(T[:, :, 3] - t[:len(T)]).max(axis=1)
    print(extra, "frames", len(rows), "max kf", meta[:, 2].max(), "err max %.3f" % err.max(), "per-frame", np.round(err, 3).tolist(), "LOST" if "LOST" in r.stdout else "")
