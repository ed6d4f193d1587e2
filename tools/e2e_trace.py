"""Pipeline trace of one host-buffer stereo frontend call (VSLAM_FRONT_TRACE=1):  python tools/e2e_trace.py [pairs]"""
import os
import sys
import time

os.environ["VSLAM_FRONT_TRACE"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vslam_b200_loader
pkg = vslam_b200_loader.pkg
from bench import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = pkg.Context(device=0, max_images=2 * B, max_keypoints=2304, max_ba_poses=0, max_ba_points=0, max_ba_obs=0)
P1, P2 = pkg.synth.stereo_projection_matrices()
L, R = make_batch(pkg, B, 0)
hl, hr = torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory()
out = ctx.alloc_frontend_outputs(B)
keep = []
for k, v in list(out.items()):
    t = torch.from_numpy(v.view(np.uint8).reshape(-1)).pin_memory(); keep.append(t)
    out[k] = t.numpy().view(v.dtype).reshape(v.shape)
for i in range(3):
    t0 = time.perf_counter()
    ctx.stereo_frontend(hl, hr, P1, P2, nfeatures=2000, out=out)
    print("call ms", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
