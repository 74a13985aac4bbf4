"""Micro-probe (not part of the product): how fast can 82 MB be written / 82 MB read on this B200 when
rotating over 4 buffers (same L2 regime as bench.py)?  Sets the practical floor for splat_fwd / bwd_rows."""
import torch
dev = torch.device("cuda:0")
N = 8 * 64 * 200 * 200
bufs = [torch.empty(N, device=dev) for _ in range(4)]
src = [torch.randn(N, device=dev) for _ in range(4)]

def t(fn, iters=200):
    for i in range(8): fn(i % 4)
    torch.cuda.synchronize()
    gs = []
    side = torch.cuda.Stream()
    for i in range(4):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn(i)
        gs.append(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): gs[i % 4].replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

print("memset 82MB          us", t(lambda i: bufs[i].zero_()))
print("fill kernel 82MB     us", t(lambda i: bufs[i].fill_(1.0)))
print("copy 82MB->82MB      us", t(lambda i: bufs[i].copy_(src[i])))
print("read-reduce 82MB     us", t(lambda i: src[i].sum()))
print("empty kernel launch  us", t(lambda i: bufs[i][:32].fill_(1.0)))
