"""Pinned-host -> device copy bandwidth of this box (the bound of bench.py's e2e leg: 151 MB of symbols per 16384-frame step)."""
import torch, time
n = 150994944
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"H2D {n / 1e6:.0f} MB pinned: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s")
for chunk in (4718592, 18874368):  # 512 and 2048 frames
    hc, dc = h[:chunk], d[:chunk]
    e0.record()
    for _ in range(20): dc.copy_(hc, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"H2D {chunk / 1e6:.1f} MB pinned: {ms:.3f} ms = {chunk / ms / 1e6:.1f} GB/s")
