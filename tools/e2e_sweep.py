"""Host-buffer stereo frontend (e2e) under different chunk sizes / staging modes:  python tools/e2e_sweep.py
Each setting runs in its own process (the library reads the environment once)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
import vslam_b200_loader
pkg = vslam_b200_loader.pkg
from bench import make_batch, W, H
B = 256
ctx = pkg.Context(device=0, max_images=2 * B, max_keypoints=2304, max_ba_poses=0, max_ba_points=0, max_ba_obs=0)
P1, P2 = pkg.synth.stereo_projection_matrices()
sets = []
for k in range(2):
    L, R = make_batch(pkg, B, 17 * k)
    sets.append((torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory()))
out = ctx.alloc_frontend_outputs(B)
keep = []
for k, v in list(out.items()):
    t = torch.from_numpy(v.view(np.uint8).reshape(-1)).pin_memory(); keep.append(t)
    out[k] = t.numpy().view(v.dtype).reshape(v.shape)
for i in range(3):
    ctx.stereo_frontend(sets[i & 1][0], sets[i & 1][1], P1, P2, nfeatures=2000, out=out)
t0 = time.perf_counter()
n = 8
for i in range(n):
    ctx.stereo_frontend(sets[i & 1][0], sets[i & 1][1], P1, P2, nfeatures=2000, out=out)
dt = (time.perf_counter() - t0) / n
print("RESULT", B / dt, dt * 1e3)
''' % ROOT

res = {}
for name, env in [("chunk16", {"VSLAM_FRONT_CHUNK": "16"}), ("chunk32", {"VSLAM_FRONT_CHUNK": "32"}),
                  ("chunk64", {"VSLAM_FRONT_CHUNK": "64"}), ("chunk32_aligned", {"VSLAM_FRONT_CHUNK": "32", "VSLAM_FRONT_ALIGN": "1"}),
                  ("chunk64_aligned", {"VSLAM_FRONT_CHUNK": "64", "VSLAM_FRONT_ALIGN": "1"})]:
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    res[name] = [float(x) for x in line[0].split()[1:]] if line else r.stderr[-300:]
    print(name, res[name])
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "e2e_sweep.json"), "w"), indent=1)
