#!/usr/bin/env python
"""Quick per-kernel and whole-step timing of the run-plan path at one workload (development aid; bench.py is the record).

    python scripts/bench_runplan_quick.py [cfg2] [iters]
Each item is captured as one CUDA graph per rotating buffer set (4 sets > L2) and replayed back to back."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lss_carla_b200 import ops  # noqa: E402
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad  # noqa: E402
from lss_carla_b200.tools import gen_dx_bx  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    cfg = CONFIGS[name]
    dev = torch.device("cuda:0")
    dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    fH, fW = cfg.fHW
    prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, dx, bx, nx)
    ds = torch.arange(*cfg.dbound, dtype=torch.float)
    fr = torch.empty(ds.shape[0], fH, fW, 3)
    fr[..., 0] = torch.linspace(0, cfg.final_dim[1] - 1, fW).view(1, 1, fW)
    fr[..., 1] = torch.linspace(0, cfg.final_dim[0] - 1, fH).view(1, fH, 1)
    fr[..., 2] = ds.view(-1, 1, 1)
    fr = fr.to(dev)

    class S:
        pass
    sets = []
    for i in range(4):
        b = make_batch(cfg, i, "train")
        s = S()
        s.cal = {k: b[k].to(dev) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
        s.dn = b["depthnet_out"].to(dev)
        s.gb = make_bev_grad(cfg, i).to(dev).contiguous(memory_format=torch.channels_last)
        s.rp = ops.RunPlan(prob, dev)
        s.bev = torch.empty(prob.bev_shape, device=dev).contiguous(memory_format=torch.channels_last)
        s.grad = torch.empty_like(s.dn)
        s.lift = (torch.empty((2, prob.B * prob.N, prob.D, fH, fW), device=dev), torch.empty((prob.B * prob.N, fH * fW, prob.C), device=dev))
        s.s1, s.s2 = torch.cuda.Stream(), torch.cuda.Stream()
        sets.append(s)

    def plan(s):
        ops.build_runplan(prob, fr, s.cal["trans"].reshape(-1, 3), s.cal["post_trans"].reshape(-1, 3), rots=s.cal["rots"],
                          intrins=s.cal["intrins"], post_rots=s.cal["post_rots"], plan=s.rp)

    def lift(s):
        s.pr, s.ct = ops.lift_prepare(prob, s.dn, out=s.lift)

    def zero(s):
        ops.bev_zero(prob, dev, out=s.bev)

    def gather(s):
        ops.splat_fwd_cl(prob, s.rp, s.pr, s.ct, out=s.bev, precleared=True)

    def fwd_serial(s):
        ops.splat_fwd_cl(prob, s.rp, s.pr, s.ct, out=s.bev, precleared=False)

    def bwd(s):
        ops.splat_bwd_cl(prob, s.rp, s.gb, s.pr, s.ct, out=s.grad)

    def fwd_op(s):            # zero || lift, then gather (plan cached)
        cur = torch.cuda.current_stream()
        s.s1.wait_stream(cur)
        with torch.cuda.stream(s.s1):
            zero(s)
        lift(s)
        cur.wait_stream(s.s1)
        gather(s)

    def step(s, upto=3):
        cur = torch.cuda.current_stream()
        s.s1.wait_stream(cur)
        s.s2.wait_stream(cur)
        with torch.cuda.stream(s.s1):
            zero(s)
        with torch.cuda.stream(s.s2):
            lift(s)
        plan(s)
        cur.wait_stream(s.s1)
        cur.wait_stream(s.s2)
        if upto >= 2:
            gather(s)
        if upto >= 3:
            bwd(s)

    NP = int(os.environ.get("ZPARTS", "2"))

    def step_split(s, upto=3):
        """zero slices chained on a side stream; slice k+1 starts when slice k AND the k-th plan kernel are done"""
        cur = torch.cuda.current_stream()
        s.s1.wait_stream(cur)
        s.s2.wait_stream(cur)
        with torch.cuda.stream(s.s2):
            lift(s)
        with torch.cuda.stream(s.s1):
            ops.bev_zero(prob, dev, out=s.bev, part=0, n_parts=NP)
        plan(s)
        with torch.cuda.stream(s.s1):
            for k in range(1, NP):
                ops.bev_zero(prob, dev, out=s.bev, part=k, n_parts=NP)
        cur.wait_stream(s.s1)
        cur.wait_stream(s.s2)
        if upto >= 2:
            gather(s)
        if upto >= 3:
            bwd(s)

    def prologue(s):
        s.pr, s.ct = ops.liftsplat_prologue(prob, s.dn, s.lift, s.bev, s.rp, fr, s.cal["trans"].reshape(-1, 3), s.cal["post_trans"].reshape(-1, 3),
                                            rots=s.cal["rots"], intrins=s.cal["intrins"], post_rots=s.cal["post_rots"])

    def prologue_cached(s):      # plan cached: zero + lift only
        s.pr, s.ct = ops.liftsplat_prologue(prob, s.dn, s.lift, s.bev)

    def fused_step(s, upto=3):
        prologue(s)
        if upto >= 2:
            gather(s)
        if upto >= 3:
            bwd(s)

    def fused_fwd_cached(s):
        prologue_cached(s)
        gather(s)

    def pair(a, b=None, c=None):
        def f(s):
            cur = torch.cuda.current_stream()
            s.s1.wait_stream(cur)
            with torch.cuda.stream(s.s1):
                a(s)
            if c is not None:
                s.s2.wait_stream(cur)
                with torch.cuda.stream(s.s2):
                    c(s)
                cur.wait_stream(s.s2)
            if b is not None:
                b(s)
            cur.wait_stream(s.s1)
        return f

    def timeit(fn):
        for s in sets:
            fn(s)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        graphs = []
        for s in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                fn(s)
            graphs.append(g)
        torch.cuda.synchronize()
        for g in graphs:
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            graphs[i % 4].replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    for s in sets:
        step(s)
    torch.cuda.synchronize()
    res = {"workload": name}
    for nm, fn in (("plan", plan), ("lift", lift), ("zero", zero), ("gather", gather), ("fwd_serial(zero+gather)", fwd_serial),
                   ("fwd_op(zero||lift->gather)", fwd_op), ("bwd", bwd), ("zero||plan", pair(zero, plan)), ("zero||lift", pair(zero, lift)),
                   ("lift||plan", pair(lift, plan)), ("zero||gather", pair(zero, gather)), ("zero||bwd", pair(zero, bwd)), ("zero_side_only", pair(zero)), ("step_upto_plan", lambda s: step(s, 1)),
                   ("step_upto_fwd", lambda s: step(s, 2)), ("step", step), ("prologue", prologue), ("prologue_cached(zero+lift)", prologue_cached), ("fused_upto_fwd", lambda s: fused_step(s, 2)),
                   ("fused_fwd_cached", fused_fwd_cached), ("fused_step", fused_step)):
        res[nm + "_us"] = round(timeit(fn), 2)
    res["mpoints_per_s"] = round(cfg.points / res["fused_step_us"], 1)
    # timeline of one step in the rotating (L2-cold) regime: globaltimer stamps inside the kernels
    import ctypes as C
    from lss_carla_b200 import _lib
    L = _lib.lib()
    if not hasattr(L, "lss_debug_runplan_timeline"):          # library built without -DLSS_RP_TIMELINE
        print(json.dumps(res))
        return
    L.lss_debug_runplan_timeline.restype = C.c_int
    L.lss_debug_runplan_timeline.argtypes = [C.c_int, C.c_void_p]
    side = torch.cuda.Stream()
    graphs = []
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fused_step(s)
        graphs.append(g)
    for i in range(12):
        graphs[i % 4].replay()
    torch.cuda.synchronize()
    L.lss_debug_runplan_timeline(1, None)
    graphs[0].replay()
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 8)()
    L.lss_debug_runplan_timeline(0, out)
    t0 = min(out[0], out[2])
    res["timeline_us"] = {nm: [round((out[2 * k] - t0) / 1e3, 1), round((out[2 * k + 1] - t0) / 1e3, 1)] for k, nm in enumerate(("zero", "index", "classify", "gather_col"))}
    res["counters"] = sets[0].rp.counters.cpu().tolist()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
