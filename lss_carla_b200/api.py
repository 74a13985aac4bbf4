"""Public entry point of the path for callers that do not hold a `LiftSplatShoot` module.

    ls = LiftSplat(grid_conf, data_aug_conf, device="cuda:0")
    bev = ls(depthnet_out, rots, trans, intrins, post_rots, post_trans)     # [B, nz*C, nx, ny], autograd-aware

`depthnet_out` is what `CamEncode.depthnet` produces (reference src/models.py:56); the remaining
arguments are the calibration tensors of `LiftSplatShoot.forward` (models.py:256).  Inputs may live on
the host (pinned or not): they are copied to the device on the current stream.  With
`inverse_mode="reference"` and HOST calibration the two 3x3 inverses run on the host exactly as the
reference does (models.py:180,186) without any device->host round trip; `inverse_mode="device"` uses the
closed-form inverse kernel (no LAPACK call, graph-capturable).

The BEV comes back in torch.channels_last strides by default (same logical shape and values as the reference's
NCHW-contiguous tensor): that is the layout cuDNN wants for `bevencode.conv1` (measured on B200: BevEncode forward +
backward 7.8 ms against 11.6 ms from an NCHW input) and the one the sort-free run plan writes without a store pass;
`bev_channels_last=False` returns the reference's memory format through the tile-plan kernels.
"""
from __future__ import annotations

import collections
import hashlib

import torch

from . import models, ops
from .tools import gen_dx_bx


def calibration_key(rots, trans, intrins, post_rots, post_trans):
    """Key of a batch's calibration for the plan cache: the bytes of the five HOST tensors of LiftSplatShoot.forward
    (models.py:256) -- 33 floats per camera.  None if any of them lives on the device (hashing would cost a round trip)."""
    h = hashlib.blake2b(digest_size=16)
    for t in (rots, trans, intrins, post_rots, post_trans):
        if t.is_cuda:
            return None
        h.update(str(tuple(t.shape)).encode())
        h.update(t.detach().contiguous().float().numpy().tobytes())
    return h.digest()


class LiftSplat:
    """`plan_cache`: number of run plans kept, keyed by the bytes of the host calibration (0 = every call rebuilds its plan).
    The forward only reads a plan, so a batch whose calibration repeats -- evaluation (static resize / crop,
    src/data_simbev.py:135-143), or a repeated training batch -- skips the plan build; `plan_cache_stats()` reports hits."""

    def __init__(self, grid_conf, data_aug_conf, C=64, downsample=16, splat_mode="sorted",
                 inverse_mode="reference", bev_channels_last=True, device="cuda:0", tile_cols=0, copy_streams=False, plan_cache=0):
        self.device = torch.device(device)
        self.plan_cache_size = int(plan_cache)
        self._plans = collections.OrderedDict()
        self._hits = self._misses = 0
        dx, bx, nx = gen_dx_bx(grid_conf["xbound"], grid_conf["ybound"], grid_conf["zbound"])
        self.dx, self.bx, self.nx = dx, bx, nx
        ogfH, ogfW = data_aug_conf["final_dim"]
        fH, fW = ogfH // downsample, ogfW // downsample
        ds = torch.arange(*grid_conf["dbound"], dtype=torch.float)
        xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float)
        ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float)
        fr = torch.empty(ds.shape[0], fH, fW, 3)
        fr[..., 0], fr[..., 1], fr[..., 2] = xs.view(1, 1, fW), ys.view(1, fH, 1), ds.view(-1, 1, 1)
        self.frustum = fr.to(self.device)
        self.D, self.camC = ds.shape[0], C
        self.splat_mode, self.inverse_mode, self.bev_channels_last = splat_mode, inverse_mode, bev_channels_last
        self.tile_cols = tile_cols
        self._up = self._down = None
        self.copy_streams = bool(copy_streams)

    def _dev(self, t):
        return t if t.is_cuda else t.to(self.device, non_blocking=True)

    _BIG = 1 << 18     # elements: only transfers of this size are worth a second stream (event hand-offs cost host time)

    def upload(self, t):
        """Device copy of a (pinned) host tensor on the current stream.  With `copy_streams=True` large tensors go
        through a dedicated copy stream (upload of the next call next to the kernels of the previous one); measured
        on B200 this is SLOWER for this call sequence (952 vs 1177 Mpoints/s end to end at cfg 2), so it is off."""
        if t.is_cuda:
            return t
        if t.numel() < self._BIG or not self.copy_streams:
            return t.to(self.device, non_blocking=True)
        if self._up is None:
            self._up, self._down = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._up):
            d = t.to(self.device, non_blocking=True)
        cur.wait_stream(self._up)
        d.record_stream(cur)
        return d

    def download(self, t, out):
        """Copy a device result into the (pinned) host tensor `out`; large results leave through a dedicated copy
        stream that waits for the current one, so the copy overlaps whatever is enqueued next.  Synchronise the
        device (or `sync_downloads()`) before reading `out`."""
        if t.numel() < self._BIG or not self.copy_streams:
            out.copy_(t, non_blocking=True)
            return out
        if self._down is None:
            self._up, self._down = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
        self._down.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._down):
            out.copy_(t, non_blocking=True)
        t.record_stream(self._down)
        return out

    def plan_cache_stats(self):
        n = self._hits + self._misses
        return {"hits": self._hits, "misses": self._misses, "hit_rate": (self._hits / n) if n else None, "plans_kept": len(self._plans)}

    def sync_downloads(self):
        if self._down is not None:
            self._down.synchronize()

    def __call__(self, depthnet_out, rots, trans, intrins, post_rots, post_trans, plan=None):
        B, N = trans.shape[:2]
        fH, fW = depthnet_out.shape[-2:]
        prob = models._problem_for(self, B, N, fH, fW, depthnet_out.shape[1] - self.D)
        if plan is None and models._use_runplan(self, prob):
            key = calibration_key(rots, trans, intrins, post_rots, post_trans) if self.plan_cache_size > 0 else None
            if key is not None:
                key = (key, id(prob))
                if key in self._plans:                # the calibration repeats: no plan build, lift + forward only
                    self._plans.move_to_end(key)
                    self._hits += 1
                    return ops.lift_splat(self._dev(depthnet_out), prob, self._plans[key], self.splat_mode, True)
                self._misses += 1
            # a cached plan is never rebuilt in place (a graph that still needs it may be alive): fresh workspace per entry
            ws = ops.RunPlan(prob, self.device) if key is not None else models._cached_plan(self, prob, self.device, run=True)
            if self.inverse_mode == "reference":      # host tensors: LAPACK inverse where the data already is
                M1 = torch.inverse(post_rots.cpu() if post_rots.is_cuda else post_rots)
                M2h = torch.inverse(intrins.cpu() if intrins.is_cuda else intrins)
                M1, M2 = self._dev(M1), self._dev(rots).matmul(self._dev(M2h))
                build = dict(frustum=self.frustum, trans=self._dev(trans).reshape(-1, 3), post_trans=self._dev(post_trans).reshape(-1, 3),
                             M1=M1.reshape(-1, 3, 3), M2=M2.reshape(-1, 3, 3))
            else:
                build = models.runplan_build_args(self, prob, self._dev(rots), self._dev(trans), self._dev(intrins),
                                                  self._dev(post_rots), self._dev(post_trans))
            out = ops.lift_splat(self._dev(depthnet_out), prob, ws, self.splat_mode, True, build=build)   # plan build + lift + forward
            if key is not None:
                self._plans[key] = ws
                while len(self._plans) > self.plan_cache_size:
                    self._plans.popitem(last=False)
            return out
        if plan is None:
            if self.inverse_mode == "reference":
                # host tensors: LAPACK inverse where the data already is; device tensors: the reference's round trip
                M1 = torch.inverse(post_rots.cpu() if post_rots.is_cuda else post_rots)
                M2h = torch.inverse(intrins.cpu() if intrins.is_cuda else intrins)
                M1, M2 = self._dev(M1), self._dev(rots).matmul(self._dev(M2h))
                calib = (self.frustum, self._dev(post_trans).reshape(-1, 3), M1.reshape(-1, 3, 3), M2.reshape(-1, 3, 3),
                         self._dev(trans).reshape(-1, 3))
                plan = ops.build_plan(prob, calib=calib, sorted=(self.splat_mode == "sorted"),
                                      plan=models._cached_plan(self, prob, self.device))
            else:
                plan = models.plan_from_calibration(self, prob, self._dev(rots), self._dev(trans), self._dev(intrins),
                                                    self._dev(post_rots), self._dev(post_trans))
        return ops.lift_splat(self._dev(depthnet_out), prob, plan, self.splat_mode, self.bev_channels_last)


def _step_fn(ls, prob, x, cal, grad_bev, grad_out, probe_out, dev, tile_cols):
    """The kernels of one forward + backward step on preallocated device buffers (no autograd graph, capturable).
    channels_last + sorted: prologue (zero-fill || lift || run index) -> classify + gather in its shadow -> backward gather;
    otherwise: tile plan -> lift on a forked branch -> gather + store -> gradient rows + gather."""
    if models._use_runplan(ls, prob):
        rp = ops.RunPlan(prob, dev)
        bev = torch.empty(prob.bev_shape, dtype=torch.float32, device=dev).contiguous(memory_format=torch.channels_last)
        lift_out = (torch.empty((2, prob.B * prob.N, prob.D, prob.fH, prob.fW), dtype=torch.float32, device=dev),
                    torch.empty((prob.B * prob.N, prob.fH * prob.fW, prob.C), dtype=torch.float32, device=dev))
        gb = grad_bev if grad_bev.is_contiguous(memory_format=torch.channels_last) else grad_bev.contiguous(memory_format=torch.channels_last)
        flat = bev.permute(0, 2, 3, 1).reshape(-1)            # physical order: a view

        def step():
            rots, trans, intrins, post_rots, post_trans = cal
            _, pr, ct = ops.liftsplat_forward(prob, rp, x, lift_out, bev, ls.frustum, trans.reshape(-1, 3), post_trans.reshape(-1, 3),
                                              rots=rots, intrins=intrins, post_rots=post_rots)
            ops.splat_bwd_cl(prob, rp, gb, pr, ct, out=grad_out)
            probe_out.copy_(flat[:probe_out.numel()])
        return step, (rp, bev, lift_out, gb)
    ws = ops.Plan(prob, dev, tile_cols)
    vsum = torch.empty((ws.layout.n_rows_cap, prob.C), dtype=torch.float32, device=dev)
    rows = torch.empty((max(prob.n_voxels, ws.layout.n_rows_cap), prob.C), dtype=torch.float32, device=dev)
    side = torch.cuda.Stream(device=dev)           # lift_prepare next to the plan build, inside the graph

    def step():
        cur = torch.cuda.current_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            pr, ct = ops.lift_prepare(prob, x)
        plan = ops.build_plan_raw(prob, ls.frustum, *cal, sorted=(ls.splat_mode == "sorted"), plan=ws)
        cur.wait_stream(side)
        bev = ops.splat_fwd(prob, plan, pr, ct, ls.splat_mode, ls.bev_channels_last, voxel_sums=vsum)
        ops.splat_bwd(prob, plan, grad_bev, pr, ct, rows, out=grad_out)
        probe_out.copy_(bev.reshape(-1)[:probe_out.numel()])
    return step, (ws, vsum, rows, side)


class StepGraph:
    """One lift-splat forward + backward as a CUDA graph bound to a set of PINNED host buffers.

        g = StepGraph(ls, host, grad_bev, stream)      # host: dict of pinned tensors, see below
        ...write the next batch into host["depthnet_out"], host["rots"], ... ; g.replay() ; later: stream.synchronize()

    Every replay copies the step's inputs host -> device, builds the plan (device inverse mode), runs lift + splat,
    the backward against `grad_bev` (a device tensor, e.g. what BevEncode's backward produced) and copies the input
    gradient and a probe of the BEV back into host["grad_out"] / host["probe"] -- all as nodes of one graph, so the
    host pays one launch per step and graphs replayed on different streams overlap their copies with each other's
    kernels.  Keys of `host`: depthnet_out, rots, trans, intrins, post_rots, post_trans, grad_out, probe."""

    def __init__(self, ls: LiftSplat, host: dict, grad_bev, stream=None):
        if ls.inverse_mode != "device":
            raise ValueError("StepGraph needs inverse_mode='device' (the LAPACK inverse of the reference mode runs on the host)")
        self.ls, self.host, self.stream = ls, host, stream or torch.cuda.Stream(device=ls.device)
        B, N = host["trans"].shape[:2]
        fH, fW = host["depthnet_out"].shape[-2:]
        prob = models._problem_for(ls, B, N, fH, fW, host["depthnet_out"].shape[1] - ls.D)
        dev = ls.device
        x = torch.empty(host["depthnet_out"].shape, dtype=torch.float32, device=dev)
        cal = [torch.empty(host[k].shape, dtype=torch.float32, device=dev) for k in CALIB_KEYS]
        grad = torch.empty_like(x)
        probe = torch.empty(host["probe"].shape, dtype=torch.float32, device=dev)
        kernels, keep = _step_fn(ls, prob, x, cal, grad_bev, grad, probe, dev, ls.tile_cols)
        self._keep = (keep, x, cal, grad, probe)

        def step():
            # the kernels are called directly (no autograd graph inside the capture): forward, then the backward
            # of lift+splat against `grad_bev`, exactly what _LiftSplatFn.forward / backward do
            x.copy_(host["depthnet_out"], non_blocking=True)
            for t, k in zip(cal, CALIB_KEYS):
                t.copy_(host[k], non_blocking=True)
            kernels()
            host["grad_out"].copy_(grad, non_blocking=True)
            host["probe"].copy_(probe, non_blocking=True)

        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for _ in range(3):
                step()
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.pool = torch.cuda.graph_pool_handle()      # a memory pool of its own: graphs replay concurrently
        with torch.cuda.graph(self.graph, pool=self.pool, stream=self.stream):
            step()
        self.stream.synchronize()

    def replay(self):
        with torch.cuda.stream(self.stream):
            self.graph.replay()


CALIB_KEYS = ("rots", "trans", "intrins", "post_rots", "post_trans")
_CALIB_FLOATS = {"rots": 9, "trans": 3, "intrins": 9, "post_rots": 9, "post_trans": 3}


def pinned_step_buffers(B, N, channels, fH, fW, probe=1024):
    """Pinned host buffers of one step for `StepPipeline`: ONE input block (`in_block`: the depthnet output followed by
    the five calibration tensors, 33 floats per camera) and ONE output block (`out_block`: input gradient, BEV probe); the
    named entries are views, so a step is one copy host -> device and one back."""
    n_x = B * N * channels * fH * fW
    blk = torch.empty(n_x + B * N * 33, dtype=torch.float32).pin_memory()
    h = {"in_block": blk, "depthnet_out": blk[:n_x].view(B * N, channels, fH, fW)}
    off = n_x
    for k in CALIB_KEYS:
        n = _CALIB_FLOATS[k]
        h[k] = blk[off:off + B * N * n].view((B, N, 3, 3) if n == 9 else (B, N, 3))
        off += B * N * n
    out = torch.empty(n_x + probe, dtype=torch.float32).pin_memory()
    h["out_block"] = out
    h["grad_out"] = out[:n_x].view(B * N, channels, fH, fW)
    h["probe"] = out[n_x:]
    return h


class PipelineStreams:
    """The streams a group of `StepPipeline`s shares: copy-in, compute, copy-out.  `copy_in_streams` > 1 spreads the
    copy-in of consecutive steps over several streams (copy engines): two input blocks are then in flight over PCIe at
    the same time, which raises the host -> device throughput where a single copy is latency-bound."""

    def __init__(self, device, copy_in_streams=2):
        self.h2d_all = [torch.cuda.Stream(device=device) for _ in range(max(1, copy_in_streams))]
        self.h2d = self.h2d_all[0]
        self.compute, self.d2h = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        self._next = 0

    def next_h2d(self):
        s = self.h2d_all[self._next % len(self.h2d_all)]
        self._next += 1
        return s

    def all(self):
        return tuple(self.h2d_all) + (self.compute, self.d2h)


class _Event:
    """cudaEvent (no timing) owned by the C library: recorded / waited for inside lss_pipe_stage."""

    def __init__(self):
        from ._lib import lib
        self.h = lib().lss_pipe_event_create()
        if not self.h:
            raise RuntimeError("liblss_b200: lss_pipe_event_create failed")

    def synchronize(self):
        from ._lib import check, lib
        check(lib().lss_pipe_event_synchronize(self.h), "lss_pipe_event_synchronize")

    def __del__(self):
        try:
            from ._lib import lib
            lib().lss_pipe_event_destroy(self.h)
        except Exception:
            pass


class StepPipeline:
    """One lift-splat forward + backward step between PINNED host buffers, as a three-stage pipeline.

        streams = PipelineStreams(dev)
        steps = [StepPipeline(ls, pinned_step_buffers(...), grad_bev, streams) for _ in range(depth)]
        ... s = steps[i % depth]; s.done.synchronize(); write batch i into s.host[...]; s.run(); later read s.host["grad_out"]

    `run()` enqueues (1) the host -> device copy of this step's input block (depthnet output + calibration) on the copy-in
    stream, (2) ONE CUDA graph with the step's kernels (plan build with the device inverse, lift next to it on a forked
    branch, splat forward, backward against `grad_bev`) on the compute stream and (3) the device -> host copy of the output
    block (input gradient, BEV probe) on the copy-out stream, chained by events.  All instances of a group run their
    kernels on the SAME compute stream -- the kernels of different steps never interleave (co-scheduled steps were
    measured ~20 % slower per step than back-to-back ones) -- while the copies of the neighbouring steps overlap them on
    the two copy engines.  The host side of a step is ONE foreign-function call (lss_pipe_step: event waits, the two copies, the graph launch).
    Every instance owns its device buffers and plan workspace; it may be re-run once its previous results have been
    consumed (`done.synchronize()`).  `host` comes from `pinned_step_buffers`."""

    def __init__(self, ls: LiftSplat, host: dict, grad_bev, streams: PipelineStreams):
        import ctypes as C
        from ._lib import check, lib
        if ls.inverse_mode != "device":
            raise ValueError("StepPipeline needs inverse_mode='device' (the LAPACK inverse of the reference mode runs on the host)")
        if "in_block" not in host or "out_block" not in host:
            raise ValueError("StepPipeline needs the buffers of api.pinned_step_buffers (one pinned block per direction)")
        self.ls, self.host, self.streams = ls, host, streams
        self._lib, self._check = lib(), check
        dev = ls.device
        B, N = host["trans"].shape[:2]
        fH, fW = host["depthnet_out"].shape[-2:]
        prob = models._problem_for(ls, B, N, fH, fW, host["depthnet_out"].shape[1] - ls.D)
        n_x = host["depthnet_out"].numel()
        self.in_dev = torch.empty(host["in_block"].shape, dtype=torch.float32, device=dev)
        self.out_dev = torch.empty(host["out_block"].shape, dtype=torch.float32, device=dev)
        x = self.in_dev[:n_x].view(host["depthnet_out"].shape)
        cal, off = [], n_x
        for k in CALIB_KEYS:
            n = host[k].numel()
            cal.append(self.in_dev[off:off + n].view(host[k].shape))
            off += n
        grad_out = self.out_dev[:n_x].view(host["depthnet_out"].shape)
        probe_out = self.out_dev[n_x:]
        compute, keep = _step_fn(ls, prob, x, cal, grad_bev, grad_out, probe_out, dev, ls.tile_cols)
        self._keep = (keep, x, cal)

        self.ev_in, self.ev_c, self.done = _Event(), _Event(), _Event()
        self._s = tuple(C.c_void_p(st.cuda_stream) for st in (streams.next_h2d(), streams.compute, streams.d2h))
        cs = streams.compute
        cs.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cs):
            self.in_dev.copy_(host["in_block"], non_blocking=True)
            for _ in range(3):
                compute()
        cs.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=cs):
            compute()
        cs.synchronize()
        for ev in (self.ev_in, self.ev_c, self.done):  # "recorded and complete": the first run() must not wait
            check(self._lib.lss_pipe_stage(self._s[1], None, None, 0, None, None, None, ev.h), "lss_pipe_stage")
        cs.synchronize()
        # the whole of run() as ONE foreign call: (copy-in stream, compute stream, copy-out stream, the graph's cudaGraphExec_t,
        # events, input block, output block)
        self._step_args = (self._s[0], self._s[1], self._s[2], C.c_void_p(int(self.graph.raw_cuda_graph_exec())),
                           self.ev_in.h, self.ev_c.h, self.done.h,
                           C.c_void_p(self.in_dev.data_ptr()), C.c_void_p(host["in_block"].data_ptr()),
                           C.c_size_t(host["in_block"].numel() * 4),
                           C.c_void_p(host["out_block"].data_ptr()), C.c_void_p(self.out_dev.data_ptr()),
                           C.c_size_t(host["out_block"].numel() * 4))

    def run(self):
        # copy-in after the previous run of this instance has read its inputs; the kernels after the inputs have arrived and the
        # previous results have left the device; copy-out after the kernels (lss_pipe_step, include/lss_b200.h)
        st = self._lib.lss_pipe_step(*self._step_args)
        if st:
            self._check(st, "lss_pipe_step")
