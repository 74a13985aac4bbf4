"""Public entry point of the path for callers that do not hold a `LiftSplatShoot` module.

    ls = LiftSplat(grid_conf, data_aug_conf, device="cuda:0")
    bev = ls(depthnet_out, rots, trans, intrins, post_rots, post_trans)     # [B, nz*C, nx, ny], autograd-aware

`depthnet_out` is what `CamEncode.depthnet` produces (reference src/models.py:56); the remaining
arguments are the calibration tensors of `LiftSplatShoot.forward` (models.py:256).  Inputs may live on
the host (pinned or not): they are copied to the device on the current stream.  With
`inverse_mode="reference"` and HOST calibration the two 3x3 inverses run on the host exactly as the
reference does (models.py:180,186) without any device->host round trip; `inverse_mode="device"` uses the
closed-form inverse kernel (no LAPACK call, graph-capturable).
"""
from __future__ import annotations

import os

import torch

from . import models, ops
from .tools import gen_dx_bx


class LiftSplat:
    def __init__(self, grid_conf, data_aug_conf, C=64, downsample=16, splat_mode="sorted",
                 inverse_mode="reference", bev_channels_last=False, device="cuda:0", tile_cols=0, copy_streams=False):
        self.device = torch.device(device)
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
        self.copy_streams = bool(copy_streams or os.environ.get("LSS_API_COPY_STREAMS"))

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

    def sync_downloads(self):
        if self._down is not None:
            self._down.synchronize()

    def __call__(self, depthnet_out, rots, trans, intrins, post_rots, post_trans, plan=None):
        B, N = trans.shape[:2]
        fH, fW = depthnet_out.shape[-2:]
        prob = models._problem_for(self, B, N, fH, fW, depthnet_out.shape[1] - self.D)
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
        ws = ops.Plan(prob, ls.device, ls.tile_cols)       # a workspace of its own: graphs may overlap
        dev = ls.device

        vsum = torch.empty((ws.layout.n_rows_cap, prob.C), dtype=torch.float32, device=dev)
        rows = torch.empty((max(prob.n_voxels, ws.layout.n_rows_cap), prob.C), dtype=torch.float32, device=dev)
        self._keep = (ws, vsum, rows)

        def step():
            # the kernels are called directly (no autograd graph inside the capture): forward, then the backward
            # of lift+splat against `grad_bev`, exactly what _LiftSplatFn.forward / backward do
            x = host["depthnet_out"].to(dev, non_blocking=True)
            cal = [host[k].to(dev, non_blocking=True) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")]
            plan = ops.build_plan_raw(prob, ls.frustum, *cal, sorted=(ls.splat_mode == "sorted"), plan=ws)
            pr, ct = ops.lift_prepare(prob, x)
            bev = ops.splat_fwd(prob, plan, pr, ct, ls.splat_mode, ls.bev_channels_last, voxel_sums=vsum)
            grad = ops.splat_bwd(prob, plan, grad_bev, pr, ct, rows)
            host["grad_out"].copy_(grad, non_blocking=True)
            host["probe"].copy_(bev.reshape(-1)[: host["probe"].numel()], non_blocking=True)

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
