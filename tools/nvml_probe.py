"""ClockSampler self-test under load: python tools/nvml_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
x = torch.randn(8192, 8192, device="cuda")
torch.cuda.synchronize()
c = bench.ClockSampler(0)
c.start()
t0 = time.time()
for _ in range(40):
    y = x @ x
t1 = time.time()
torch.cuda.synchronize()
t2 = time.time()
print("enqueue %.1f ms, wait %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
print(c.stop())
