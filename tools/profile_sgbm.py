"""Small dense-stereo run for ncu: python tools/profile_sgbm.py [pairs] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vslam_b200_loader
pkg = vslam_b200_loader.pkg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
s = pkg.synth
pairs = [s.synth_pair(i)[:2] for i in range(B)]
L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
H, W = L.shape[1:]
ctx = pkg.Context(max_images=2, max_keypoints=2048)
dev = torch.device("cuda:0")
dl, dr = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
d16 = torch.empty((B, H, W), dtype=torch.int16, device=dev)
torch.cuda.synchronize()
for _ in range(iters):
    ctx.sgbm_compute_dev(dl, dr, B, W, H, W, W * H, d16)
ctx.synchronize()
print("valid fraction", float((d16 != -16).float().mean()))
