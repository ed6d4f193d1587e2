"""Per-stage wall clock of the sequential VO loop (VSLAM_VO_TRACE=1):  python tools/vo_trace.py [run_vslam args...]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vslam_b200_loader
pkg = vslam_b200_loader.pkg

n = 24
lefts, rights, t, _ = pkg.synth.synth_sequence(3, n)
exe = os.path.join(ROOT, "stereo-visual-slam_b200", "run_vslam")
with tempfile.TemporaryDirectory() as d:
    os.makedirs(d + "/image_0")
    os.makedirs(d + "/image_1")
    for i in range(n):
        pkg.synth.write_pgm(f"{d}/image_0/{i:06d}.pgm", lefts[i])
        pkg.synth.write_pgm(f"{d}/image_1/{i:06d}.pgm", rights[i])
    extra = sys.argv[1:] or ["--nfeatures", "1000", "--anms", "110"]
    with tempfile.TemporaryDirectory() as w:
        r = subprocess.run([exe, d + "/", str(n), *extra], cwd=w, capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, VSLAM_VO_TRACE="1"))
    print(r.stdout[-3000:])
    print(r.stderr[-6000:])
