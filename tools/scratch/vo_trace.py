import os, subprocess, sys, tempfile
sys.path.insert(0, "/root/repo")
import vslam_b200_loader
pkg = vslam_b200_loader.pkg
d = tempfile.mkdtemp(); n = 12
lefts, rights, t, disp = pkg.synth.synth_sequence(3, n)
os.makedirs(d + "/image_0"); os.makedirs(d + "/image_1")
for i in range(n):
    pkg.synth.write_pgm(f"{d}/image_0/{i:06d}.pgm", lefts[i]); pkg.synth.write_pgm(f"{d}/image_1/{i:06d}.pgm", rights[i])
for extra in (["--dense"], ["--dense", "--nfeatures", "1000", "--anms", "110"]):
    r = subprocess.run(["/root/repo/stereo-visual-slam_b200/run_vslam", d + "/", str(n), *extra], cwd=tempfile.mkdtemp(), capture_output=True, text=True,
                       env=dict(os.environ, VSLAM_VO_TRACE="1"))
    print(extra); print("\n".join(r.stderr.splitlines()[-4:]))
    print("\n".join(" ".join(l.split()[14:]) for l in r.stdout.splitlines() if l.startswith("frame "))[-200:])
