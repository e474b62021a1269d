import torch, time
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("H2D pinned 512 MiB: %.1f GB/s" % (n / e0.elapsed_time(e1) / 1e6))
e0.record(); h.copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("D2H pinned 512 MiB: %.1f GB/s" % (n / e0.elapsed_time(e1) / 1e6))
