#!/usr/bin/env python
"""DEV: do the zero-fill and the plan build of one captured step overlap in TIME?  (globaltimer stamps inside the kernels)"""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lss_carla_b200 import ops, _lib
from lss_carla_b200.synthetic import CONFIGS, make_batch
from lss_carla_b200.tools import gen_dx_bx
cfg = CONFIGS["cfg2"]; dev = torch.device("cuda:0")
dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound); fH, fW = cfg.fHW
prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, dx, bx, nx)
from lss_carla_b200 import api
ls = api.LiftSplat(cfg.grid_conf, cfg.data_aug_conf, device=dev)
b = make_batch(cfg, 0, "train"); cal = {k: b[k].to(dev) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
rp = ops.RunPlan(prob, dev); bev = torch.empty(prob.bev_shape, device=dev).contiguous(memory_format=torch.channels_last)
s1 = torch.cuda.Stream()
L = _lib.lib(); L.lss_debug_runplan_timeline.restype = C.c_int; L.lss_debug_runplan_timeline.argtypes = [C.c_int, C.c_void_p]
def plan(): ops.build_runplan(prob, ls.frustum, cal["trans"].reshape(-1, 3), cal["post_trans"].reshape(-1, 3), rots=cal["rots"], intrins=cal["intrins"], post_rots=cal["post_rots"], plan=rp)
def step(order):
    cur = torch.cuda.current_stream(); s1.wait_stream(cur)
    if order == 0:
        with torch.cuda.stream(s1): ops.bev_zero(prob, dev, out=bev)
        plan()
    else:
        plan()
        with torch.cuda.stream(s1): ops.bev_zero(prob, dev, out=bev)
    cur.wait_stream(s1)
for order in (0, 1):
    for mode in ("eager", "graph"):
        step(order); torch.cuda.synchronize()
        if mode == "graph":
            g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
            with torch.cuda.graph(g, stream=side): step(order)
            run = g.replay
        else:
            run = lambda: step(order)
        run(); torch.cuda.synchronize()
        L.lss_debug_runplan_timeline(1, None)
        run(); torch.cuda.synchronize()
        out = (C.c_ulonglong * 8)(); L.lss_debug_runplan_timeline(0, out)
        t0 = min(out[0], out[2], out[4])
        print(f"order={order} {mode}: zero [{(out[0]-t0)/1e3:.1f}, {(out[1]-t0)/1e3:.1f}] index [{(out[2]-t0)/1e3:.1f}, {(out[3]-t0)/1e3:.1f}] classify [{(out[4]-t0)/1e3:.1f}, {(out[5]-t0)/1e3:.1f}] us")
