"""Does splitting the 3.55 MB host -> device copy over several streams (copy engines) raise its throughput?
Times back-to-back copies of the block as 1, 2 and 4 concurrent pieces, alone and with a device -> host copy stream running."""
import torch
dev = torch.device("cuda:0")
n = 3554496 // 4
h = torch.randn(n).pin_memory(); d = torch.empty(n, device=dev)
h2 = torch.empty(n).pin_memory(); d2 = torch.randn(n, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]
back = torch.cuda.Stream()

def run(parts, reps=300, with_d2h=False):
    torch.cuda.synchronize()
    bounds = [n * i // parts for i in range(parts + 1)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams[:parts]:
        s.wait_event(e0)
    back.wait_event(e0)
    for _ in range(reps):
        for i, s in enumerate(streams[:parts]):
            with torch.cuda.stream(s):
                d[bounds[i]:bounds[i + 1]].copy_(h[bounds[i]:bounds[i + 1]], non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(back):
                h2.copy_(d2, non_blocking=True)
    for s in streams[:parts]:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    return us

for with_d2h in (False, True):
    for parts in (1, 2, 4):
        run(parts, 20, with_d2h)
        us = run(parts, 300, with_d2h)
        print(f"H2D 3.55 MB in {parts} piece(s){' + concurrent D2H' if with_d2h else ''}: {us:.1f} us = {n * 4 / us / 1e3:.1f} GB/s")
