#!/usr/bin/env python
"""A few eager steps of the run-plan path at one workload, for ncu (scripts/gpu_profile.sh captures bench.py; this is the
short form used while tuning):  ncu --set full -k regex:'^k_' --launch-skip 9 --launch-count 3 python scripts/profile_step.py"""
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
mode = sys.argv[2] if len(sys.argv) > 2 else "forward"
cfg = CONFIGS[name]
dev = torch.device("cuda:0")
dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
fH, fW = cfg.fHW
prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, dx, bx, nx)
fr = torch.from_numpy(create_frustum(cfg.final_dim, list(cfg.dbound))).to(dev)
sets = []
for i in range(4):
    b = make_batch(cfg, i, "train")
    cal = dict(trans=b["trans"].to(dev).reshape(-1, 3), post_trans=b["post_trans"].to(dev).reshape(-1, 3), rots=b["rots"].to(dev),
               intrins=b["intrins"].to(dev), post_rots=b["post_rots"].to(dev))
    sets.append(dict(cal=cal, dn=b["depthnet_out"].to(dev), gb=make_bev_grad(cfg, i).to(dev).contiguous(memory_format=torch.channels_last),
                     rp=ops.RunPlan(prob, dev), bev=torch.empty(prob.bev_shape, device=dev).contiguous(memory_format=torch.channels_last)))
for it in range(8):
    s = sets[it % 4]
    if mode == "forward":
        bev, pr, ct = ops.liftsplat_forward(prob, s["rp"], s["dn"], None, s["bev"], fr, **s["cal"])
    else:       # separate pieces: prologue without zero-fill, one-launch forward on a pre-cleared tensor
        pr, ct = ops.liftsplat_prologue(prob, s["dn"], None, None, s["rp"], fr, **s["cal"])
        ops.splat_fwd_cl(prob, s["rp"], pr, ct, out=s["bev"], precleared=True)
    ops.splat_bwd_cl(prob, s["rp"], s["gb"], pr, ct)
torch.cuda.synchronize()
print("ok")
