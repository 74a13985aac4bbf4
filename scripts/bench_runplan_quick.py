#!/usr/bin/env python
"""Quick per-kernel and whole-step timing of the run-plan path at one workload (development aid; bench.py is the record).

    [LSS_TIMELINE=1] python scripts/bench_runplan_quick.py [cfg2] [iters]
Each item is captured as one CUDA graph per rotating buffer set (4 sets > L2) and replayed back to back."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lss_carla_b200 import _lib, ops  # noqa: E402

if os.environ.get("LSS_TIMELINE"):      # a second build with globaltimer stamps inside the kernels
    _lib.SO_PATH = _lib.build_library(out=os.path.join(ROOT, "lss_carla_b200", "liblss_b200_timeline.so"), defines=("LSS_RP_TIMELINE",))
    _lib._lib = None
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
        sets.append(s)

    cal_args = lambda s: dict(trans=s.cal["trans"].reshape(-1, 3), post_trans=s.cal["post_trans"].reshape(-1, 3), rots=s.cal["rots"],  # noqa: E731
                              intrins=s.cal["intrins"], post_rots=s.cal["post_rots"])

    def plan(s):
        ops.build_runplan(prob, fr, plan=s.rp, **cal_args(s))

    def lift(s):
        s.pr, s.ct = ops.lift_prepare(prob, s.dn, out=s.lift)

    def zero(s):
        ops.bev_zero(prob, dev, out=s.bev)

    def fwd_precleared(s):       # classify + gather alone (timing only: the tensor is not cleared)
        ops.splat_fwd_cl(prob, s.rp, s.pr, s.ct, out=s.bev, precleared=True)

    def fwd_ordered(s):          # one launch from an existing plan: zero CTAs + classify + gather
        ops.splat_fwd_cl(prob, s.rp, s.pr, s.ct, out=s.bev)

    def bwd(s):
        ops.splat_bwd_cl(prob, s.rp, s.gb, s.pr, s.ct, out=s.grad)

    def prologue(s):             # lift + plan, no zero-fill
        s.pr, s.ct = ops.liftsplat_prologue(prob, s.dn, s.lift, None, s.rp, fr, **cal_args(s))

    def forward(s):              # lss_liftsplat_forward: zero-fill || lift || plan -> classify + gather
        if os.environ.get("QUICK_NOZERO"):       # measurement: the same chain without any zero-fill traffic
            prologue(s)
            fwd_precleared(s)
            return
        _, s.pr, s.ct = ops.liftsplat_forward(prob, s.rp, s.dn, s.lift, s.bev, fr, **cal_args(s))

    def forward_persistent(s):   # output tensor kept between steps: only the rows of the previous step are cleared
        _, s.pr, s.ct = ops.liftsplat_forward(prob, s.rp, s.dn, s.lift, s.bev, fr, persistent=True, **cal_args(s))

    def step_persistent(s):
        forward_persistent(s)
        bwd(s)

    def forward_kept(s):
        _, s.pr, s.ct = ops.liftsplat_forward(prob, s.rp, s.dn, s.lift, s.bev)

    if os.environ.get("QUICK_PDL") == "0":
        ops.set_option("pdl", 0)

    def step(s, upto=2):
        forward(s)
        if upto >= 2:
            bwd(s)

    def step_kept(s):
        forward_kept(s)
        bwd(s)

    import ctypes
    _rt = ctypes.CDLL("libcudart.so.12")
    _rt.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
    side = [torch.cuda.Stream() for _ in sets]

    def forward_fork(kind):      # experiment: the zero-fill as a memset / ATen fill on a forked branch, forward on a pre-cleared tensor
        def f(s):
            cur = torch.cuda.current_stream()
            st = side[sets.index(s)]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                if kind == "memset":
                    _rt.cudaMemsetAsync(s.bev.data_ptr(), 0, s.bev.numel() * 4, ctypes.c_void_p(st.cuda_stream))
                else:
                    s.bev.zero_()
            prologue(s)
            cur.wait_stream(st)
            fwd_precleared(s)
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
    for nm, fn in (("plan", plan), ("lift", lift), ("zero", zero), ("prologue(lift+plan)", prologue), ("fwd_precleared", fwd_precleared),
                   ("fwd_ordered", fwd_ordered), ("bwd", bwd), ("forward", forward), ("forward_fork_memset", forward_fork("memset")), ("forward_fork_fill", forward_fork("fill")), ("forward_kept_plan", forward_kept), ("step", step),
                   ("step_kept_plan", step_kept), ("forward_persistent", forward_persistent), ("step_persistent", step_persistent)):
        res[nm + "_us"] = round(timeit(fn), 2)
    res["mpoints_per_s"] = round(cfg.points / res["step_us"], 1)
    res["counters"] = sets[0].rp.counters.cpu().tolist()
    # timeline of one step in the rotating (L2-cold) regime: globaltimer stamps inside the kernels
    import ctypes as C
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
            {"ordered": fwd_ordered, "persistent": step_persistent}.get(os.environ.get("QUICK_TL"), step)(s)
        graphs.append(g)
    for i in range(12):
        graphs[i % 4].replay()
    torch.cuda.synchronize()
    L.lss_debug_runplan_timeline(1, None)
    graphs[0].replay()
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 10)()
    L.lss_debug_runplan_timeline(0, out)
    t0 = min(out[0], out[6]) if os.environ.get("QUICK_TL") == "ordered" else min(out[2], out[4])
    res["timeline_us"] = {nm: [round((out[2 * k] - t0) / 1e3, 1), round((out[2 * k + 1] - t0) / 1e3, 1)]
                          for k, nm in enumerate(("zero", "index", "lift", "fwd_columns", "fwd_shared"))}
    if hasattr(L, "lss_debug_runplan_marks"):    # phase stamps of the forward's column CTAs (thread 0): after READY, staged, after the
        n = prob.B * prob.N * prob.fW            # zero wait, exclusive loop done, barrier, shared voxels done
        m = (C.c_ulonglong * (8 * n))()
        L.lss_debug_runplan_marks.argtypes = [C.c_void_p, C.c_int]
        L.lss_debug_runplan_marks(m, n)
        import numpy as np
        mk = (np.array(m, dtype=np.float64).reshape(n, 8)[:, :4] - float(t0)) / 1e3
        names = ("start", "staged", "zero_ok", "excl_done")
        res["column_cta_marks_us"] = {nm: [round(float(np.percentile(mk[:, k], q)), 1) for q in (0, 50, 90, 100)] for k, nm in enumerate(names)}
        dur = np.diff(mk, axis=1)
        res["column_cta_phase_us"] = {f"{names[k]}->{names[k + 1]}": [round(float(np.percentile(dur[:, k], q)), 1) for q in (50, 90, 100)] for k in range(3)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
