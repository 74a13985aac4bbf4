#!/usr/bin/env python
"""Small driver for compute-sanitizer: every kernel of both plans, forward and backward, at the tiny and
cfg1 shapes, with programmatic dependent launch on and off, twice into the same workspaces (self-cleaning scratch).

    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python scripts/sanitize_case.py
One tool per gpurun call (B200_PROFILING.md).  Results are summarised in profiles/r02_sanitizer.md."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lss_carla_b200 import ops  # noqa: E402
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad  # noqa: E402
from lss_carla_b200.tools import gen_dx_bx  # noqa: E402

dev = torch.device("cuda:0")
for pdl in (1, 0):
    ops.set_option("pdl", pdl)
    for name in ("tiny", "cfg1"):
        cfg = CONFIGS[name]
        dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
        fH, fW = cfg.fHW
        prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, dx, bx, nx)
        ds = torch.arange(*cfg.dbound, dtype=torch.float)
        fr = torch.empty(ds.shape[0], fH, fW, 3)
        fr[..., 0] = torch.linspace(0, cfg.final_dim[1] - 1, fW).view(1, 1, fW)
        fr[..., 1] = torch.linspace(0, cfg.final_dim[0] - 1, fH).view(1, fH, 1)
        fr[..., 2] = ds.view(-1, 1, 1)
        fr = fr.to(dev)
        rp, tp = ops.RunPlan(prob, dev), ops.Plan(prob, dev)
        for seed, aug in ((0, "train"), (1, "full")):
            b = make_batch(cfg, seed, aug)
            cal = {k: b[k].to(dev) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
            dn = b["depthnet_out"].to(dev)
            gb = make_bev_grad(cfg, seed).to(dev)
            # run plan: two-launch forward (zero-fill || lift || index -> classify + gather), backward
            bev = torch.full(prob.bev_shape, float("nan"), device=dev).contiguous(memory_format=torch.channels_last)
            _, pr, ct = ops.liftsplat_forward(prob, rp, dn, None, bev, fr, cal["trans"].reshape(-1, 3), cal["post_trans"].reshape(-1, 3),
                                              rots=cal["rots"], intrins=cal["intrins"], post_rots=cal["post_rots"])
            g1 = ops.splat_bwd_cl(prob, rp, gb.contiguous(memory_format=torch.channels_last), pr, ct)
            # ... and with the prologue's own zero role + a pre-cleared forward
            bev1 = torch.full(prob.bev_shape, float("nan"), device=dev).contiguous(memory_format=torch.channels_last)
            ops.liftsplat_prologue(prob, None, None, bev1)
            ops.splat_fwd_cl(prob, rp, pr, ct, out=bev1, precleared=True)
            bev2 = torch.full(prob.bev_shape, float("nan"), device=dev).contiguous(memory_format=torch.channels_last)
            ops.splat_fwd_cl(prob, rp, pr, ct, out=bev2)           # zero CTAs inside the forward kernel
            assert torch.equal(bev, bev1) and torch.equal(bev, bev2)
            # tile plan: voxel index, scatter, sort, lift, gather + store, gradient rows + gather; atomic and red modes
            ops.build_plan_raw(prob, fr, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"], sorted=True, plan=tp)
            pr2, ct2 = ops.lift_prepare(prob, dn)
            b2 = ops.splat_fwd(prob, tp, pr2, ct2, "sorted", False)
            g2 = ops.splat_bwd(prob, tp, gb, pr2, ct2)
            ops.splat_fwd(prob, tp, pr2, ct2, "atomic", False)
            ops.splat_fwd(prob, tp, pr2, ct2, "red", True)
            torch.cuda.synchronize()
            assert torch.equal(bev.contiguous(), b2) and torch.equal(g1, g2)
            print(f"pdl={pdl} {name} seed={seed} {aug}: ok", flush=True)
ops.set_option("pdl", 1)
print("sanitize_case done")
