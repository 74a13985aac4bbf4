import torch
dev = torch.device("cuda:0")
n = 3554496 // 4
h = torch.randn(n).pin_memory(); d = torch.empty(n, device=dev); h2 = torch.empty(n).pin_memory()
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h2.copy_(d, non_blocking=True))):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 10
    print(name, "3.55 MB pinned:", round(us, 1), "us ->", round(n * 4 / us / 1e3, 1), "GB/s")
