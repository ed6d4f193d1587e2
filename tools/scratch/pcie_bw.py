import torch, time
n = 119453696
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(43943424, dtype=torch.uint8).pin_memory()
d2 = torch.empty(43943424, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D GB/s", n * 10 / e0.elapsed_time(e1) / 1e6)
e0.record()
for _ in range(10): h2.copy_(d2, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("D2H GB/s", 43943424 * 10 / e0.elapsed_time(e1) / 1e6)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("bidirectional: H2D GB/s", n * 10 / dt / 1e9, "D2H GB/s", 43943424 * 10 / dt / 1e9)
