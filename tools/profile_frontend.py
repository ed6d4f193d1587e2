"""Small frontend run for ncu: python tools/profile_frontend.py [pairs] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vslam_b200_loader
pkg = vslam_b200_loader.pkg
from bench import make_batch, to_pitched, W, H, NFEAT, PITCH
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cap = 2304
ctx = pkg.Context(max_images=2 * B, max_keypoints=cap)
dev = torch.device("cuda:0")
s = pkg.synth
K = s.kitti_K()
P1 = np.hstack([K, np.zeros((3, 1))]); P2 = np.hstack([K, K @ np.array([[-s.BASELINE_M], [0], [0]])])
L, R = make_batch(pkg, B, 0)
dl, dr = to_pitched(torch, L, dev), to_pitched(torch, R, dev)  # 16-byte aligned rows, as bench.py
d_kp = torch.zeros((2 * B, cap, 7), dtype=torch.int32, device=dev); d_desc = torch.zeros((2 * B, cap, 32), dtype=torch.uint8, device=dev)
d_nkp = torch.zeros(2 * B, dtype=torch.int32, device=dev); d_m = torch.zeros((B, cap, 4), dtype=torch.int32, device=dev)
d_nm = torch.zeros(B, dtype=torch.int32, device=dev); d_xyz = torch.zeros((B, cap, 3), dtype=torch.float32, device=dev)
d_fl = torch.zeros((B, cap), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for _ in range(iters):
    ctx.stereo_frontend_dev(dl, dr, B, W, H, PITCH, PITCH * H, P1, P2, None, d_kp, d_desc, d_nkp, d_m, d_nm, d_xyz, d_fl, nfeatures=NFEAT)
ctx.synchronize()
print("kp", d_nkp[:4].tolist(), "matches", d_nm[:4].tolist())
