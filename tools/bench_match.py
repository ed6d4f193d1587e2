"""Micro-benchmark of K10 (device-resident, batched).  python tools/bench_match.py [N] [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vslam_b200_loader
pkg = vslam_b200_loader.pkg

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = pkg.Context(max_images=B, max_keypoints=max(N, 64))
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
dev = torch.device("cuda:0")
q = torch.from_numpy(np.stack([pkg.synth.synth_descriptors(s, N) for s in range(B)])).to(dev)
t = torch.from_numpy(np.stack([pkg.synth.synth_descriptors(1000 + s, N) for s in range(B)])).to(dev)
nq = torch.full((B,), N, dtype=torch.int32, device=dev)
out = torch.zeros((B, N, 4), dtype=torch.int32, device=dev)
n_out = torch.zeros((B,), dtype=torch.int32, device=dev)
torch.cuda.synchronize()
for b in (1, B):
    with torch.cuda.stream(stream):
        for _ in range(5):
            ctx.match_hamming_batch_dev(q, nq, N, t, nq, N, b, N, True, 2.0, 30.0, out, N, n_out)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        it = 50
        for _ in range(it):
            ctx.match_hamming_batch_dev(q, nq, N, t, nq, N, b, N, True, 2.0, 30.0, out, N, n_out)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    pairs = b * N * N
    print(f"N={N} batch={b}: {ms*1e3:.1f} us/call, {ms*1e3/b:.2f} us/pair-set, {pairs/ms/1e6:.1f} G dist/s, "
          f"popc {pairs*8/ms/1e9:.2f} Tpopc/s, algorithmic HBM {(b*(2*N*32+N*16))/ms/1e6:.1f} GB/s, matches[0]={int(n_out[0])}")
