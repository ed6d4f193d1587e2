import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, vslam_b200_loader
pkg = vslam_b200_loader.pkg
from bench import make_batch, W, H, NFEAT
B = 64; cap = 2304
ctx = pkg.Context(max_images=2 * B, max_keypoints=cap)
dev = torch.device("cuda:0")
s = pkg.synth; K = s.kitti_K()
P1 = np.hstack([K, np.zeros((3, 1))]); P2 = np.hstack([K, K @ np.array([[-s.BASELINE_M], [0], [0]])])
L, R = make_batch(pkg, B, 0)
dl, dr = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
d_kp = torch.zeros((2 * B, cap, 7), dtype=torch.int32, device=dev); d_desc = torch.zeros((2 * B, cap, 32), dtype=torch.uint8, device=dev)
d_nkp = torch.zeros(2 * B, dtype=torch.int32, device=dev); d_m = torch.zeros((B, cap, 4), dtype=torch.int32, device=dev)
d_nm = torch.zeros(B, dtype=torch.int32, device=dev); d_xyz = torch.zeros((B, cap, 3), dtype=torch.float32, device=dev)
d_fl = torch.zeros((B, cap), dtype=torch.uint8, device=dev)
ctx.set_concurrency(False)
for _ in range(3):
    ctx.stereo_frontend_dev(dl, dr, B, W, H, W, W * H, P1, P2, None, d_kp, d_desc, d_nkp, d_m, d_nm, d_xyz, d_fl, nfeatures=NFEAT)
ctx.synchronize(); ctx.timing_enable(True)
for _ in range(10):
    ctx.stereo_frontend_dev(dl, dr, B, W, H, W, W * H, P1, P2, None, d_kp, d_desc, d_nkp, d_m, d_nm, d_xyz, d_fl, nfeatures=NFEAT)
kt = ctx.timing_read()
print({k: round(v[0] / 10, 4) for k, v in kt.items()}, "sum", round(sum(v[0] for v in kt.values()) / 10, 3))
