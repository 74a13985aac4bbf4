#!/usr/bin/env python
"""Which memory format does the consumer of the BEV tensor prefer?  

The lift-splat hands its result to `bevencode.conv1` (7x7, stride 2, 64 -> 64, reference src/models.py:97-98,:118) and
receives that layer's input gradient.  This probe times, with CUDA events on the B200:
   conv1 + bn1 + relu   forward and forward+backward,  input NCHW-contiguous vs channels_last (module converted too)
   the whole BevEncode  forward+backward in both formats
for fp32 (cuDNN's default TF32 setting of this torch build) at cfg 2 (B=8, 64 x 200 x 200).
Writes gpurun_out/conv1_layout.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lss_carla_b200.trunk import BevEncode  # noqa: E402


def ev_time(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    out = {"cudnn_allow_tf32": torch.backends.cudnn.allow_tf32, "cudnn_benchmark": torch.backends.cudnn.benchmark}
    for B in (8,):
        x0 = torch.randn(B, 64, 200, 200, device=dev)
        for fmt_name, fmt in (("nchw", torch.contiguous_format), ("channels_last", torch.channels_last)):
            for bench_flag in (False, True):
                torch.backends.cudnn.benchmark = bench_flag
                enc = BevEncode(64, 1).to(dev).to(memory_format=fmt).train()
                stem = torch.nn.Sequential(enc.conv1, enc.bn1, enc.relu)
                x = x0.clone().contiguous(memory_format=fmt).requires_grad_(True)

                def stem_fwd():
                    with torch.no_grad():
                        return stem(x)

                def stem_step():
                    x.grad = None
                    y = stem(x)
                    y.backward(torch.ones_like(y))

                def enc_step():
                    x.grad = None
                    y = enc(x)
                    y.backward(torch.ones_like(y))

                key = f"B{B}_{fmt_name}_cudnnbench{int(bench_flag)}"
                r = {"stem_fwd_ms": ev_time(stem_fwd), "stem_fwd_bwd_ms": ev_time(stem_step), "bevencode_fwd_bwd_ms": ev_time(enc_step)}
                stem_step()
                r["input_grad_is_channels_last"] = bool(x.grad.is_contiguous(memory_format=torch.channels_last) and not x.grad.is_contiguous())
                r["input_grad_strides"] = list(x.grad.stride())
                out[key] = r
                print(key, json.dumps(r), flush=True)
        # NCHW-contiguous module fed a channels_last tensor (what install(bev_channels_last=True) gives an unconverted model)
        torch.backends.cudnn.benchmark = False
        enc = BevEncode(64, 1).to(dev).train()
        x = x0.clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)

        def enc_step2():
            x.grad = None
            y = enc(x)
            y.backward(torch.ones_like(y))

        r = {"bevencode_fwd_bwd_ms": ev_time(enc_step2)}
        enc_step2()
        r["input_grad_strides"] = list(x.grad.stride())
        out[f"B{B}_channels_last_input_nchw_module"] = r
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "conv1_layout.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
