import time, torch
n = 1 << 20
d = torch.empty(n * 121, dtype=torch.uint8, device="cuda")
h = torch.empty(n * 121, dtype=torch.uint8).pin_memory()
a_h = torch.empty(n * 16, dtype=torch.uint8).pin_memory(); a_d = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
for _ in range(3): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("D2H 127 MB pinned: %.3f ms = %.1f GB/s" % (dt * 1e3, n * 121 / dt / 1e9))
s2 = torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): a_d.copy_(a_h, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("D2H 127 MB + concurrent H2D 16.8 MB: %.3f ms" % (dt * 1e3))
