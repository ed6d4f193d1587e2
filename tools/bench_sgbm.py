"""Dense-stereo timing: python tools/bench_sgbm.py [pairs] [iters]  -> per-kernel CUDA-event times + cv2 beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vslam_b200_loader
pkg = vslam_b200_loader.pkg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
s = pkg.synth
pairs = [s.synth_pair(i)[:2] for i in range(B)]
L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
H, W = L.shape[1:]
ctx = pkg.Context(max_images=2, max_keypoints=2048)
dev = torch.device("cuda:0")
dl, dr = torch.from_numpy(L).to(dev), torch.from_numpy(R).to(dev)
d16 = torch.empty((B, H, W), dtype=torch.int16, device=dev)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
torch.cuda.set_stream(stream)
for _ in range(2):
    ctx.sgbm_compute_dev(dl, dr, B, W, H, W, W * H, d16)
torch.cuda.synchronize()
ctx.set_concurrency(False)  # kernels alone on the stream for the per-kernel numbers
ctx.timing_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ctx.sgbm_compute_dev(dl, dr, B, W, H, W, W * H, d16)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
kt = {k: v for k, v in ctx.timing_read().items() if k.startswith("sgbm")}
ctx.timing_enable(False)
ctx.set_concurrency(True)
# same without the per-launch events
e0.record()
for _ in range(iters):
    ctx.sgbm_compute_dev(dl, dr, B, W, H, W, W * H, d16)
e1.record(); torch.cuda.synchronize()
ms2 = e0.elapsed_time(e1) / iters
W1 = W - 96
vol = H * W1 * 96 * 2
alg = {"sgbm_cost_kernel": vol, "sgbm_vertical_kernel": vol * 4, "sgbm_row_forward_kernel": vol * 5, "sgbm_row_backward_kernel": vol * 2}
out = {"pairs": B, "ms_per_batch": ms2, "ms_per_pair": ms2 / B, "fps": B / ms2 * 1e3, "ms_per_batch_with_events": ms,
       "kernels": {k: {"ms_per_launch": v[0] / max(v[1], 1), "launches": v[1],
                       "alg_GBps": (alg[k] * B / (v[0] / max(v[1], 1) * 1e-3) / 1e9 if k in alg else None)} for k, v in kt.items()}}
import cv2
sg = cv2.StereoSGBM_create(0, 96, 9, 8 * 81, 32 * 81, 1, 63, 10, 100, 32)
t = time.perf_counter(); n = 0
while time.perf_counter() - t < 3 and n < B:
    ref = sg.compute(L[n], R[n]); n += 1
out["cv2_ms_per_pair"] = (time.perf_counter() - t) / n * 1e3
out["cv2_threads"] = cv2.getNumThreads()
out["parity_pair0"] = bool(np.array_equal(d16[0].cpu().numpy(), sg.compute(L[0], R[0])))
print(json.dumps(out))
