#!/usr/bin/env python
"""Soak test of the flag protocol between the three forward grids (READY, zero_done, scratch reset): N graph replays of
lss_liftsplat_forward + backward over rotating buffer sets, the BEV poisoned with NaN before every step, every result compared
on the device with the one-launch forward from a separately built plan.  Prints the number of mismatching steps (must be 0).

    python scripts/soak_forward.py [cfg2] [replays]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lss_carla_b200 import ops  # noqa: E402
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad  # noqa: E402
from lss_carla_b200.tools import gen_dx_bx  # noqa: E402
from oracle.lss_oracle import create_frustum  # noqa: E402  (constants only)

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
replays = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
cfg = CONFIGS[name]
dev = torch.device("cuda:0")
dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
fH, fW = cfg.fHW
prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, dx, bx, nx)
fr = torch.from_numpy(create_frustum(cfg.final_dim, list(cfg.dbound))).to(dev)
sets = []
for i in range(4):
    b = make_batch(cfg, i, ("train", "full", "eval", "full")[i])
    cal = dict(trans=b["trans"].to(dev).reshape(-1, 3), post_trans=b["post_trans"].to(dev).reshape(-1, 3), rots=b["rots"].to(dev),
               intrins=b["intrins"].to(dev), post_rots=b["post_rots"].to(dev))
    dn = b["depthnet_out"].to(dev)
    gb = make_bev_grad(cfg, i).to(dev).contiguous(memory_format=torch.channels_last)
    ref_rp = ops.build_runplan(prob, fr, **cal)
    pr, ct = ops.lift_prepare(prob, dn)
    want = ops.splat_fwd_cl(prob, ref_rp, pr, ct, out=ops.bev_zero(prob, dev), precleared=True)
    want_g = ops.splat_bwd_cl(prob, ref_rp, gb, pr, ct)
    sets.append(dict(cal=cal, dn=dn, gb=gb, want=want, want_g=want_g, rp=ops.RunPlan(prob, dev),
                     bev=torch.empty(prob.bev_shape, device=dev).contiguous(memory_format=torch.channels_last), grad=torch.empty_like(dn),
                     lift=(torch.empty((2, prob.B * prob.N, prob.D, fH, fW), device=dev), torch.empty((prob.B * prob.N, fH * fW, prob.C), device=dev))))
bad = torch.zeros((), dtype=torch.int64, device=dev)


def step(s):
    s["bev"].fill_(float("nan"))
    _, pr, ct = ops.liftsplat_forward(prob, s["rp"], s["dn"], s["lift"], s["bev"], fr, **s["cal"])
    ops.splat_bwd_cl(prob, s["rp"], s["gb"], pr, ct, out=s["grad"])
    bad.add_(((s["bev"] != s["want"]).any() | (s["grad"] != s["want_g"]).any()).long())


side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for s in sets:
        step(s)
torch.cuda.synchronize()
graphs = []
for s in sets:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step(s)
    graphs.append(g)
for i in range(replays):
    graphs[i % 4].replay()
torch.cuda.synchronize()
print(f"{name}: {replays} replays, mismatching steps: {int(bad)}; scratch clean: {all(not s['rp'].scratch.any() for s in sets)}")
