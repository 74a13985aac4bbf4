"""Torch-facing operators of the B200 lift-splat library.

PyTorch is plumbing here: it owns device memory, streams and autograd bookkeeping; all arithmetic of the
path runs in liblss_b200.so (hand-written CUDA for sm_100a) through the C ABI of include/lss_b200.h.
Nothing in this module falls back to ATen: CPU tensors or a missing library raise.

Reference functions replaced (paths under the reference root):
    LiftSplatShoot.get_geometry     src/models.py:170-190   -> geometry()
    CamEncode.get_depth_feat (lift) src/models.py:49-61     -> lift_prepare() (operands only)
    LiftSplatShoot.voxel_pooling    src/models.py:204-246   -> voxel_pooling() / lift_splat()
    QuickCumsum / cumsum_trick      src/tools.py:182-219    -> QuickCumsum / cumsum_trick
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import torch

from . import _lib

from ._lib import (LAYOUT_CHANNELS_LAST, LAYOUT_NCHW, SPLAT_MODES, VARIANTS, ZERO_ORDERED, ZERO_PRECLEARED, LssPlanLayout, LssProblem, LssRunplanLayout, check,
                   lib)


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------

def _nvtx(name):
    """NVTX range around a public operator (ranges around geometry / lift / splat for nsys timelines)."""
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapped(*a, **k):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def set_option(name: str, value: int):
    """Process-wide library options (lss_set_option): "pdl" = programmatic dependent launch of the kernel chains (default 1)."""
    check(lib().lss_set_option({"pdl": 0}[name], int(value)), "lss_set_option")


def _f32c(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the lift-splat path has no CPU implementation")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


@dataclass
class Problem:
    """Static description of one lift-splat problem (mirrors lss_problem)."""
    B: int
    N: int
    D: int
    fH: int
    fW: int
    C: int
    nx: tuple
    dx: tuple   # float32 values as python floats
    lo: tuple   # bx - dx/2 evaluated in float32 (models.py:212)

    def __post_init__(self):
        p = LssProblem()
        p.B, p.N, p.D, p.fH, p.fW, p.C = self.B, self.N, self.D, self.fH, self.fW, self.C
        p.nx, p.ny, p.nz = (int(v) for v in self.nx)
        for k in range(3):
            p.dx[k] = self.dx[k]
            p.lo[k] = self.lo[k]
        self.c = p

    @property
    def n_points(self):
        return self.B * self.N * self.D * self.fH * self.fW

    @property
    def n_voxels(self):
        return self.B * int(self.nx[0]) * int(self.nx[1]) * int(self.nx[2])

    @property
    def bev_shape(self):
        return (self.B, int(self.nx[2]) * self.C, int(self.nx[0]), int(self.nx[1]))

    @staticmethod
    def from_grid(B, N, D, fH, fW, C, dx, bx, nx):
        """dx, bx: float32 tensors (any device), nx: integer tensor -- the model's parameters
        (models.py:143-145).  `lo` is evaluated with float32 tensor ops exactly like models.py:212."""
        dxc, bxc = dx.detach().float().cpu(), bx.detach().float().cpu()
        lo = bxc - dxc / 2.
        return Problem(B, N, D, fH, fW, C, tuple(int(v) for v in nx.detach().cpu().tolist()),
                       tuple(float(v) for v in dxc.tolist()), tuple(float(v) for v in lo.tolist()))


class Plan:
    """Per-batch index structures (voxel ids, tile buckets) living in one workspace tensor."""

    def __init__(self, prob: Problem, device, tile_cols: int = 0):
        self.prob = prob
        self.layout = LssPlanLayout()
        check(lib().lss_plan_layout_init(C.byref(prob.c), tile_cols, C.byref(self.layout)), "lss_plan_layout_init")
        self.ws = torch.zeros(self.layout.bytes, dtype=torch.uint8, device=device)   # scratch counters start at 0
        self.sorted = False
        self.built = False
        self.busy = False     # True between a differentiable forward and its backward (or the death of its graph)
        self.generation = 0   # bumped by every build: a backward checks that its plan was not rebuilt in between

    def _view(self, off, n, dtype):
        item = torch.empty(0, dtype=dtype).element_size()
        return self.ws[off:off + n * item].view(dtype)

    def scratch_rows(self, name, rows, device):
        """Row workspace f32[rows, C] of the tile-plan kernels (compact voxel sums / gradient rows), allocated once per plan
        (the kernels of one plan are stream-ordered: forward and backward never use the same buffer at the same time)."""
        cache = self.__dict__.setdefault("_rows_cache", {})
        t = cache.get(name)
        if t is None or t.shape[0] < rows or t.device != device:
            t = cache[name] = torch.empty((rows, self.prob.C), dtype=torch.float32, device=device)
        return t

    @property
    def vox(self):
        return self._view(self.layout.off_vox, self.prob.n_points, torch.int32)

    @property
    def entries(self):
        return self._view(self.layout.off_entries, self.prob.n_points, torch.int32)

    @property
    def tile_start(self):
        return self._view(self.layout.off_tile_start, self.layout.n_tiles + 1, torch.int32)

    @property
    def tile_nseg(self):
        return self._view(self.layout.off_tile_nseg, self.layout.n_tiles, torch.int32)

    @property
    def tile_row0(self):
        """First compact row of every tile (sorted plans)."""
        return self._view(self.layout.off_tile_row0, self.layout.n_tiles, torch.int32)

    @property
    def n_rows(self):
        """Device scalar: number of non-empty voxels of the batch (sorted plans)."""
        return self._view(self.layout.off_counters, 1, torch.int32)

    def reset(self):
        check(lib().lss_plan_reset(C.byref(self.layout), _ptr(self.ws), _stream()), "lss_plan_reset")


# ------------------------------------------------------------------------------------------------
# geometry / plan
# ------------------------------------------------------------------------------------------------

def calib_matrices_reference(rots, intrins, post_rots):
    """M1 = inverse(post_rots), M2 = rots @ inverse(intrins) with the reference's own calls
    (models.py:180,186): LAPACK inverse on the host, product on the device.  Bit-identical inputs for
    the kernels, at the price of the reference's host round trip (one stream synchronisation)."""
    M1 = torch.inverse(post_rots.cpu()).to(post_rots.device)
    M2 = rots.matmul(torch.inverse(intrins.cpu()).to(rots.device))
    return M1.contiguous(), M2.contiguous()


def calib_matrices_device(rots, intrins, post_rots):
    """Same matrices from a closed-form 3x3 inverse on the device: no host round trip, graph-capturable;
    last-bit differences from LAPACK are possible ('inverse modes')."""
    rots, intrins, post_rots = (_f32c(t, n) for t, n in ((rots, "rots"), (intrins, "intrins"), (post_rots, "post_rots")))
    n_cams = rots.numel() // 9
    M1, M2 = torch.empty_like(rots), torch.empty_like(rots)
    check(lib().lss_calib_matrices(n_cams, _ptr(rots), _ptr(intrins), _ptr(post_rots), _ptr(M1), _ptr(M2), _stream()),
          "lss_calib_matrices")
    return M1, M2


@_nvtx("lss:geometry")
def geometry(prob: Problem, frustum, post_trans, M1, M2, trans):
    """get_geometry (models.py:170-190) given prepared matrices -> f32[B,N,D,fH,fW,3]."""
    frustum, post_trans, M1, M2, trans = (_f32c(t, n) for t, n in (
        (frustum, "frustum"), (post_trans, "post_trans"), (M1, "M1"), (M2, "M2"), (trans, "trans")))
    out = torch.empty((prob.B, prob.N, prob.D, prob.fH, prob.fW, 3), dtype=torch.float32, device=frustum.device)
    check(lib().lss_geometry(C.byref(prob.c), _ptr(frustum), _ptr(post_trans), _ptr(M1), _ptr(M2), _ptr(trans),
                             _ptr(out), _stream()), "lss_geometry")
    return out


def voxel_index(prob: Problem, geom=None, calib=None, want=("vox", "idx", "kept", "rank")):
    """Per-point quantisation dump (models.py:212-229).  `geom` f32[...,3] or `calib` =
    (frustum, post_trans, M1, M2, trans).  Returns dict of tensors."""
    dev = geom.device if geom is not None else calib[0].device
    n = prob.n_points
    out = {}
    if "vox" in want:
        out["vox"] = torch.empty(n, dtype=torch.int32, device=dev)
    if "idx" in want:
        out["idx"] = torch.empty((n, 3), dtype=torch.int64, device=dev)
    if "kept" in want:
        out["kept"] = torch.empty(n, dtype=torch.uint8, device=dev)
    if "rank" in want:
        out["rank"] = torch.empty(n, dtype=torch.int64, device=dev)
    g = _f32c(geom, "geom") if geom is not None else None
    cal = [_f32c(t, "calib") for t in calib] if calib is not None else [None] * 5
    check(lib().lss_voxel_index(C.byref(prob.c), _ptr(g), *[_ptr(t) for t in cal], _ptr(out.get("vox")),
                                _ptr(out.get("idx")), _ptr(out.get("kept")), _ptr(out.get("rank")), _stream()),
          "lss_voxel_index")
    return out


@_nvtx("lss:plan_build")
def build_plan(prob: Problem, geom=None, calib=None, sorted: bool = True, plan: Plan | None = None,
               tile_cols: int = 0) -> Plan:
    """Voxel ids + tile buckets (+ in-bucket sort) for one batch; replaces models.py:212-231."""
    dev = geom.device if geom is not None else calib[0].device
    if plan is None:
        plan = Plan(prob, dev, tile_cols)
    g = _f32c(geom, "geom") if geom is not None else None
    cal = [_f32c(t, "calib") for t in calib] if calib is not None else [None] * 5
    check(lib().lss_plan_build(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), _ptr(g),
                               *[_ptr(t) for t in cal], 1 if sorted else 0, _stream()), "lss_plan_build")
    plan.sorted = bool(sorted)
    plan.built = True
    plan.generation += 1
    plan._keepalive = (g, cal)
    return plan


@_nvtx("lss:plan_build")
def build_plan_raw(prob: Problem, frustum, rots, trans, intrins, post_rots, post_trans, sorted: bool = True,
                   plan: Plan | None = None, tile_cols: int = 0) -> Plan:
    """build_plan straight from the raw calibration (device inverse mode): the 3x3 inverses are evaluated inside
    the voxel-index kernel, no separate calib launch.  Same bits as calib_matrices_device + build_plan."""
    ts = [_f32c(t, n) for t, n in ((frustum, "frustum"), (rots, "rots"), (trans, "trans"), (intrins, "intrins"),
                                   (post_rots, "post_rots"), (post_trans, "post_trans"))]
    if plan is None:
        plan = Plan(prob, ts[0].device, tile_cols)
    check(lib().lss_plan_build_raw(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), *[_ptr(t) for t in ts],
                                   1 if sorted else 0, _stream()), "lss_plan_build_raw")
    plan.sorted, plan.built, plan._keepalive = bool(sorted), True, ts
    plan.generation += 1
    return plan


def reference_order(plan: Plan):
    """The reference's `sorts` (models.py:230) as flat point indices, int64[n_kept]."""
    prob = plan.prob
    if not plan.sorted:
        raise RuntimeError("reference_order needs a plan built with sorted=True")
    scratch = torch.empty(prob.n_voxels + 1, dtype=torch.int32, device=plan.ws.device)
    order = torch.empty(prob.n_points, dtype=torch.int64, device=plan.ws.device)
    nk = torch.empty(1, dtype=torch.int32, device=plan.ws.device)
    check(lib().lss_plan_reference_order(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), _ptr(scratch),
                                         _ptr(order), _ptr(nk), _stream()), "lss_plan_reference_order")
    return order[: int(nk.item())]


# ------------------------------------------------------------------------------------------------
# lift + splat
# ------------------------------------------------------------------------------------------------

@_nvtx("lss:lift_prepare")
def lift_prepare(prob: Problem, depthnet_out, out=None):
    """softmax over depth + pixel-major context (models.py:49-61) -> (prob [BN,D,fH,fW], ctx_t [BN,HW,C]).
    `out`: optional preallocated (f32[2,BN,D,fH,fW], f32[BN,HW,C]) pair, e.g. when the call runs on a side stream."""
    bf16 = depthnet_out.dtype == torch.bfloat16     # autocast: widened on load, float32 from there on (lss_lift_prepare_bf16)
    if bf16:
        if not depthnet_out.is_cuda:
            raise RuntimeError("depthnet_out must be a CUDA tensor: the lift-splat path has no CPU implementation")
        x = depthnet_out.contiguous()
    else:
        x = _f32c(depthnet_out, "depthnet_out")
    BN, HW = prob.B * prob.N, prob.fH * prob.fW
    if tuple(x.shape) != (BN, prob.D + prob.C, prob.fH, prob.fW):
        raise ValueError(f"depthnet_out has shape {tuple(x.shape)}, expected {(BN, prob.D + prob.C, prob.fH, prob.fW)}")
    if out is not None:
        both, ct = out
    else:
        both = torch.empty((2, BN, prob.D, prob.fH, prob.fW), dtype=torch.float32, device=x.device)
        ct = torch.empty((BN, HW, prob.C), dtype=torch.float32, device=x.device)
    pr = both[0]                                   # [BN, D, fH, fW]; both[1] holds the column-major copy [BN, fW, D, fH]
    fn = lib().lss_lift_prepare_bf16 if bf16 else lib().lss_lift_prepare
    check(fn(C.byref(prob.c), _ptr(x), _ptr(pr), _ptr(ct), _ptr(both[1]), _stream()), "lss_lift_prepare")
    return pr, ct


def _prob_col(pr):
    """The column-major copy that lift_prepare wrote right behind `pr` (None for foreign tensors)."""
    base = pr._base if pr._base is not None else None
    if base is not None and base.dim() == 5 and base.shape[0] == 2 and base.data_ptr() == pr.data_ptr() and pr.is_contiguous():
        return base[1]
    return None


def _empty_bev(prob: Problem, device, channels_last: bool):
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    return torch.empty(prob.bev_shape, dtype=torch.float32, device=device, memory_format=fmt)


def _bev_layout(t):
    """Layout code of a BEV-shaped tensor, making it dense in one of the two supported formats."""
    if t.is_contiguous():
        return t, LAYOUT_NCHW
    if t.is_contiguous(memory_format=torch.channels_last):
        return t, LAYOUT_CHANNELS_LAST
    return t.contiguous(), LAYOUT_NCHW


def bev_clear(prob: Problem, device, channels_last=False):
    """A zeroed BEV tensor (models.py:240), e.g. issued early on a side stream for the `scatter` variant."""
    bev = _empty_bev(prob, device, channels_last)
    check(lib().lss_bev_clear(C.byref(prob.c), _ptr(bev), _stream()), "lss_bev_clear")
    return bev


@_nvtx("lss:splat_fwd")
def splat_fwd(prob: Problem, plan: Plan, pr, ct, mode="sorted", channels_last=False, variant="auto", out=None,
              voxel_sums=None, batch_range=(0, 0), precleared=None):
    """`out`: optional output tensor (pre-zeroed, from bev_clear, for mode 'red').  `voxel_sums`: optional
    workspace f32[plan.layout.n_rows_cap, C] of the two-kernel GROUP variant (allocated when omitted).
    `batch_range` = (b0, b1): only these samples (GROUP variant; samples are independent)."""
    if mode == "sorted" and not plan.sorted:
        raise RuntimeError("mode='sorted' needs a plan built with sorted=True")
    bev = out if out is not None else _empty_bev(prob, pr.device, channels_last)
    if voxel_sums is None and mode == "sorted" and variant != "warp" and prob.C in (32, 64, 128):
        voxel_sums = plan.scratch_rows("voxel_sums", plan.layout.n_rows_cap, pr.device)    # plan-sized: kept with the plan, not per call
    check(lib().lss_splat_fwd(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), _ptr(pr), _ptr(ct),
                              _ptr(_prob_col(pr)), _ptr(voxel_sums), _ptr(bev),
                              SPLAT_MODES[mode], LAYOUT_CHANNELS_LAST if channels_last else LAYOUT_NCHW,
                              VARIANTS[variant], int(out is not None if precleared is None else precleared),
                              int(batch_range[0]), int(batch_range[1]), _stream()), "lss_splat_fwd")
    return bev


@_nvtx("lss:splat_bwd")
def splat_bwd(prob: Problem, plan: Plan, grad_bev, pr, ct, grad_rows=None, prob_col=None, out=None, stage=0,
              batch_range=(0, 0)):
    g, layout = _bev_layout(_f32c_keep(grad_bev))
    if grad_rows is None and (layout == LAYOUT_NCHW or plan.sorted):
        grad_rows = plan.scratch_rows("grad_rows", max(prob.n_voxels, plan.layout.n_rows_cap), g.device)
    if out is None:
        out = torch.empty((prob.B * prob.N, prob.D + prob.C, prob.fH, prob.fW), dtype=torch.float32, device=g.device)
    if prob_col is None:
        prob_col = _prob_col(pr)
    check(lib().lss_splat_bwd(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), _ptr(g), layout, _ptr(pr), _ptr(ct),
                              _ptr(prob_col), _ptr(grad_rows), _ptr(out), 1 if plan.sorted else 0, int(stage),
                              int(batch_range[0]), int(batch_range[1]), _stream()), "lss_splat_bwd")
    return out


# ------------------------------------------------------------------------------------------------
# run plan: channels_last fast path (include/lss_b200.h, "Run plan")
# ------------------------------------------------------------------------------------------------

def runplan_supported(prob: Problem) -> bool:
    """Shapes the run plan takes: a run (the fH image rows of one camera column and depth bin) fits a warp, C = 32/64/128."""
    lay = LssRunplanLayout()
    return lib().lss_runplan_layout_init(C.byref(prob.c), C.byref(lay)) == 0


def runplan_raw_supported(prob: Problem) -> bool:
    """Whether build_runplan / liftsplat_prologue take the raw calibration (closed-form inverses inside the index kernel)."""
    return lib().lss_runplan_raw_supported(C.byref(prob.c)) == 0


class RunPlan:
    """Per-batch index structures of the channels_last path: voxel row per point, sub-run nodes, per-voxel list heads.
    The forward only reads them (its scratch -- progress counters, pool -- is left clean), so a built plan can be kept."""

    def __init__(self, prob: Problem, device):
        self.prob = prob
        self.layout = LssRunplanLayout()
        check(lib().lss_runplan_layout_init(C.byref(prob.c), C.byref(self.layout)), "lss_runplan_layout_init")
        self.ws = torch.zeros(self.layout.bytes, dtype=torch.uint8, device=device)   # epoch 0, empty lists
        self.built = False
        self.busy = False
        self.generation = 0
        self._bev_ref = self._bev_version = None      # the output tensor liftsplat_forward(persistent=True) is paired with

    def _view(self, off, n, dtype):
        item = torch.empty(0, dtype=dtype).element_size()
        return self.ws[off:off + n * item].view(dtype)

    @property
    def prow(self):
        """int32[B, N, fW, D, fH]: voxel row ((b*nx+ix)*ny+iy)*nz+iz of every point, -1 where dropped."""
        p = self.prob
        return self._view(self.layout.off_prow, p.n_points, torch.int32).view(p.B, p.N, p.fW, p.D, p.fH)

    @property
    def sub(self):
        """int32[B, N, fW, D, fH, 2]: at the first point of a sub-run {previous sub-run on the voxel's list (flat
        camera-column-major point index + 1, 0 = none), image-row mask}; {0, 0} elsewhere."""
        p = self.prob
        return self._view(self.layout.off_sub, 2 * p.n_points, torch.int32).view(p.B, p.N, p.fW, p.D, p.fH, 2)

    @property
    def head(self):
        """int64[n_voxels]: (build epoch << 32) | (point index + 1) of the last sub-run pushed on the voxel's list."""
        return self._view(self.layout.off_head, self.layout.n_voxels, torch.int64)

    @property
    def counters(self):
        """int32[8]: [0] epoch of the last build, [6] / [7] voxels shared by
        several sub-runs / voxels with >= 64 points summed by the last forward, the rest scratch (zero between launches)."""
        return self._view(self.layout.off_counters, 8, torch.int32)

    @property
    def scratch(self):
        """What a forward uses as scratch (all-zero between launches): counters [1..3] and [8..], the zero-fill progress counters
        and the READY flags."""
        n = (self.layout.off_head - self.layout.off_zero_done) // 4
        c = self._view(self.layout.off_counters, 64, torch.int32)
        return torch.cat((c[1:6], c[8:], self._view(self.layout.off_zero_done, n, torch.int32)))

    def lists(self):
        """The per-voxel lists of the current build as numpy arrays (synchronises; tests): (rows, points) -- for every
        sub-run on a list of this epoch, the voxel row it belongs to and its number of points -- plus the number of voxels
        whose list holds more than one sub-run."""
        import numpy as np
        sub = self.sub.reshape(-1, 2).cpu().numpy()
        prow = self.prow.reshape(-1).cpu().numpy()
        head = self.head.cpu().numpy()
        epoch = int(self.counters[0])
        live = np.flatnonzero((head >> 32) == epoch)
        cur = (head[live] & 0xFFFFFFFF).astype(np.int64)
        rows, pts, nodes = [], [], np.zeros(live.size, dtype=np.int64)
        voxel = live.copy()
        while cur.size:
            node = sub[cur - 1]
            assert (node[:, 1] != 0).all() and (prow[cur - 1] == voxel).all()
            rows.append(voxel)
            pts.append(np.array([bin(int(m) & 0xFFFFFFFF).count("1") for m in node[:, 1]], dtype=np.int64))
            keep = node[:, 0] != 0
            cur, voxel = node[keep, 0].astype(np.int64), voxel[keep]
        rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        pts = np.concatenate(pts) if pts else np.zeros(0, np.int64)
        n_shared = int((np.bincount(rows, minlength=1) > 1).sum()) if rows.size else 0
        return rows, pts, n_shared

    def reset(self):
        self._bev_ref = self._bev_version = None
        check(lib().lss_runplan_reset(C.byref(self.layout), _ptr(self.ws), _stream()), "lss_runplan_reset")


@_nvtx("lss:runplan_build")
def build_runplan(prob: Problem, frustum, trans, post_trans, M1=None, M2=None, rots=None, intrins=None, post_rots=None,
                  plan: RunPlan | None = None) -> RunPlan:
    """Run plan of one batch from calibration; replaces models.py:170-190 + :212-231.  With M1 / M2 (the reference's
    own host inverses, `calib_matrices_reference`) the voxel rows are bit-exact; without them the closed-form
    inverses are evaluated inside the kernel from rots / intrins / post_rots (no host round trip, graph-capturable)."""
    raw = M1 is None or M2 is None
    ts = [_f32c(t, n) for t, n in ((frustum, "frustum"), (post_trans, "post_trans"), (trans, "trans"))]
    mats = [_f32c(t, n) for t, n in ((rots, "rots"), (intrins, "intrins"), (post_rots, "post_rots"))] if raw else \
           [_f32c(t, n) for t, n in ((M1, "M1"), (M2, "M2"))]
    if plan is None:
        plan = RunPlan(prob, ts[0].device)
    null = C.c_void_p(0)
    args = (null, null, _ptr(ts[2]), *[_ptr(t) for t in mats]) if raw else (_ptr(mats[0]), _ptr(mats[1]), _ptr(ts[2]), null, null, null)
    check(lib().lss_runplan_build(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), _ptr(ts[0]), _ptr(ts[1]), *args, _stream()),
          "lss_runplan_build")
    plan.built, plan._keepalive = True, (ts, mats)
    plan.generation += 1
    plan._bev_ref = plan._bev_version = None           # rebuilt without clearing anybody's rows
    return plan


def _prologue_args(prob: Problem, depthnet_out, lift_out, plan, frustum, trans, post_trans, M1, M2, rots, intrins, post_rots):
    """ctypes arguments shared by lss_liftsplat_prologue / lss_liftsplat_forward: (calibration x8, lift x4), (pr, ct), keepalive."""
    null = C.c_void_p(0)
    build = plan is not None and frustum is not None
    cal = [null] * 8
    keep = []
    if build:
        raw = M1 is None or M2 is None
        ts = [_f32c(t, n) for t, n in ((frustum, "frustum"), (post_trans, "post_trans"), (trans, "trans"))]
        mats = [_f32c(t, n) for t, n in ((rots, "rots"), (intrins, "intrins"), (post_rots, "post_rots"))] if raw else \
               [_f32c(t, n) for t, n in ((M1, "M1"), (M2, "M2"))]
        keep = [ts, mats]
        cal = [_ptr(ts[0]), _ptr(ts[1])] + ([null, null, _ptr(ts[2])] + [_ptr(t) for t in mats] if raw else
                                          [_ptr(mats[0]), _ptr(mats[1]), _ptr(ts[2]), null, null, null])
    lift_args, res = [null] * 4, None
    if depthnet_out is not None:
        x = _f32c(depthnet_out, "depthnet_out")
        BN, HW = prob.B * prob.N, prob.fH * prob.fW
        if tuple(x.shape) != (BN, prob.D + prob.C, prob.fH, prob.fW):
            raise ValueError(f"depthnet_out has shape {tuple(x.shape)}, expected {(BN, prob.D + prob.C, prob.fH, prob.fW)}")
        both, ct = lift_out if lift_out is not None else (
            torch.empty((2, BN, prob.D, prob.fH, prob.fW), dtype=torch.float32, device=x.device),
            torch.empty((BN, HW, prob.C), dtype=torch.float32, device=x.device))
        lift_args, res = [_ptr(x), _ptr(both[0]), _ptr(ct), _ptr(both[1])], (both[0], ct)
        keep.append(x)
    return build, cal + lift_args, res, keep


@_nvtx("lss:prologue(lift+index)")
def liftsplat_prologue(prob: Problem, depthnet_out=None, lift_out=None, bev=None, plan: RunPlan | None = None, frustum=None, trans=None,
                       post_trans=None, M1=None, M2=None, rots=None, intrins=None, post_rots=None):
    """Fused prologue in ONE launch: run plan from the calibration (`plan` and `frustum` given; arguments as build_runplan),
    lift operands of `depthnet_out` (float32; None: off) and, optionally, a zero-fill of `bev` (channels_last; None: off) as
    independent CTA roles of one grid.  Returns (pr, ct) of the lift (or None).
    Then: splat_fwd_cl(..., out=bev), with precleared=True if `bev` was zero-filled here.  (liftsplat_forward does it all.)"""
    build, args, res, keep = _prologue_args(prob, depthnet_out, lift_out, plan, frustum, trans, post_trans, M1, M2, rots, intrins, post_rots)
    if bev is not None and not bev.is_contiguous(memory_format=torch.channels_last):
        raise RuntimeError("liftsplat_prologue zero-fills channels_last tensors only")
    lay = C.byref(plan.layout) if plan is not None else None
    check(lib().lss_liftsplat_prologue(C.byref(prob.c), lay, _ptr(plan.ws) if plan is not None else C.c_void_p(0), *args,
                                       _ptr(bev), _stream()), "lss_liftsplat_prologue")
    if build:
        plan.built, plan._keepalive = True, keep
        plan.generation += 1
        plan._bev_ref = plan._bev_version = None
    return res


@_nvtx("lss:forward(zero || lift+index || classify+gather)")
def liftsplat_forward(prob: Problem, plan: RunPlan, depthnet_out, lift_out=None, out=None, frustum=None, trans=None, post_trans=None,
                      M1=None, M2=None, rots=None, intrins=None, post_rots=None, _allow_unbuilt=False, persistent=False):
    """The whole forward of a step (lss_liftsplat_forward): three launches that run side by side -- zero-fill with progress
    counters, lift || plan build, classify + gather polling both.  `frustum` None: the plan in `plan` is kept (static calibration).
    Returns (bev, pr, ct); `bev` is channels_last, the same bits as splat_fwd(mode="sorted").

    `persistent`: the output tensor is kept between calls (`out`, or a tensor the plan owns) and only the rows the previous call
    wrote are cleared (lss_liftsplat_forward_persistent: no zero-fill of the whole tensor).  The plan remembers the tensor it is
    paired with and holds a reference to it; the first call with another tensor, or after the plan was rebuilt by anything else,
    takes the regular path once.  Nobody else may write to the tensor in between (in-place torch ops on it are detected)."""
    build, args, res, keep = _prologue_args(prob, depthnet_out, lift_out, plan, frustum, trans, post_trans, M1, M2, rots, intrins, post_rots)
    if not build and not plan.built and not _allow_unbuilt:
        raise RuntimeError("liftsplat_forward without calibration needs a built plan")
    if persistent and out is None:
        out = plan._bev_ref if plan._bev_ref is not None else _empty_bev(prob, depthnet_out.device, True)
    bev = out if out is not None else _empty_bev(prob, depthnet_out.device, True)
    if not bev.is_contiguous(memory_format=torch.channels_last):
        raise RuntimeError("liftsplat_forward writes channels_last tensors only")
    paired = (persistent and plan._bev_ref is not None and plan._bev_ref.data_ptr() == bev.data_ptr()
              and plan._bev_ref.shape == bev.shape and plan._bev_version == bev._version)
    fn, what = (lib().lss_liftsplat_forward_persistent, "lss_liftsplat_forward_persistent") if paired else \
               (lib().lss_liftsplat_forward, "lss_liftsplat_forward")
    check(fn(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), *args, _ptr(bev), _stream()), what)
    if build:
        plan.built, plan._keepalive = True, keep
        plan.generation += 1
    plan._bev_ref, plan._bev_version = (bev, bev._version) if persistent else (None, None)
    return bev, res[0], res[1]


@_nvtx("lss:bev_zero")
def bev_zero(prob: Problem, device, out=None, part=0, n_parts=1):
    """A zeroed channels_last BEV tensor (models.py:240) through the bulk-copy kernel; may be issued on a side stream
    next to the plan build and handed to splat_fwd_cl(..., out=bev, precleared=True).  (part, n_parts): only that slice."""
    bev = out if out is not None else _empty_bev(prob, device, True)
    check(lib().lss_bev_zero(C.byref(prob.c), _ptr(bev), int(part), int(n_parts), _stream()), "lss_bev_zero")
    return bev


@_nvtx("lss:splat_fwd_cl(zero+classify+gather)")
def splat_fwd_cl(prob: Problem, plan: RunPlan, pr, ct, out=None, precleared=False):
    """Deterministic forward into a channels_last BEV tensor from an existing plan and existing lift operands, ONE launch
    (zero-fill + classify + gather): same bits as splat_fwd(mode="sorted").  `precleared`: `out` is already all-zero."""
    pc = _prob_col(pr)
    if pc is None:
        raise RuntimeError("splat_fwd_cl needs the (prob, ctx) pair of lift_prepare (column-major weights)")
    bev = out if out is not None else _empty_bev(prob, pr.device, True)
    if not bev.is_contiguous(memory_format=torch.channels_last):
        raise RuntimeError("splat_fwd_cl writes channels_last tensors only")
    check(lib().lss_liftsplat_fwd_cl(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), _ptr(pc), _ptr(ct), _ptr(bev),
                                     ZERO_PRECLEARED if precleared else ZERO_ORDERED, _stream()), "lss_liftsplat_fwd_cl")
    return bev


@_nvtx("lss:splat_bwd_cl")
def splat_bwd_cl(prob: Problem, plan: RunPlan, grad_bev, pr, ct, prob_col=None, out=None):
    g = _f32c_keep(grad_bev)
    if not g.is_contiguous(memory_format=torch.channels_last):
        g = g.contiguous(memory_format=torch.channels_last)      # an NCHW gradient is transposed once (cuDNN's NHWC conv1 hands us channels_last)
    if prob_col is None:
        prob_col = _prob_col(pr)
    if out is None:
        out = torch.empty((prob.B * prob.N, prob.D + prob.C, prob.fH, prob.fW), dtype=torch.float32, device=g.device)
    check(lib().lss_liftsplat_bwd_cl(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), _ptr(g), _ptr(prob_col), _ptr(ct),
                                     _ptr(out), _stream()), "lss_liftsplat_bwd_cl")
    return out


def _f32c_keep(t):
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError("expected a float32 CUDA tensor")
    return t


def _release(plan):
    plan.busy = False


class _LiftSplatFn(torch.autograd.Function):
    """depthnet output [B*N, D+C, fH, fW] -> BEV [B, nz*C, nx, ny]; backward = fused gather."""

    @staticmethod
    def forward(ctx, depthnet_out, prob, plan, mode, channels_last, bev_out, build):
        if isinstance(plan, RunPlan):
            if depthnet_out.dtype == torch.float32 and bev_out is None:    # the whole forward, plan build included
                bev, pr, ct = liftsplat_forward(prob, plan, depthnet_out, **(build or {}))
            else:
                if build:
                    build_runplan(prob, plan=plan, **build)
                pr, ct = lift_prepare(prob, depthnet_out)
                bev = splat_fwd_cl(prob, plan, pr, ct, out=bev_out, precleared=bev_out is not None)
        else:
            pr, ct = lift_prepare(prob, depthnet_out)
            bev = splat_fwd(prob, plan, pr, ct, mode, channels_last)
        ctx.prob, ctx.plan, ctx.in_dtype, ctx.generation = prob, plan, depthnet_out.dtype, plan.generation
        ctx.save_for_backward(pr, ct, _prob_col(pr))
        if ctx.needs_input_grad[0]:
            plan.busy = True                       # released by the backward, or when the graph dies without one
            weakref.finalize(ctx, _release, plan)
        return bev

    @staticmethod
    def backward(ctx, grad_bev):
        pr, ct, pc = ctx.saved_tensors
        if ctx.plan.generation != ctx.generation:
            raise RuntimeError("the plan of this lift-splat forward was rebuilt before its backward ran "
                               "(backward(retain_graph=True) followed by another forward on the same plan?)")
        if isinstance(ctx.plan, RunPlan):
            out = splat_bwd_cl(ctx.prob, ctx.plan, grad_bev.float(), pr, ct, prob_col=pc)
        else:
            out = splat_bwd(ctx.prob, ctx.plan, grad_bev.float(), pr, ct, prob_col=pc)
        ctx.plan.busy = False
        if ctx.in_dtype != torch.float32:          # bfloat16 input: the float32 gradient is rounded once, at the very end
            out = out.to(ctx.in_dtype)
        return out, None, None, None, None, None, None


def lift_splat(depthnet_out, prob: Problem, plan, mode="sorted", channels_last=False, bev_out=None, build=None):
    """Fused lift + splat of the depthnet output through an existing plan (differentiable w.r.t.
    `depthnet_out`; geometry carries no gradient, tools.py:207).  `plan`: a tile `Plan`, or a `RunPlan`
    (deterministic, channels_last output; `bev_out`: optional pre-zeroed output from `bev_zero`; `build`: the
    calibration arguments of `build_runplan` as a dict -- the plan is then (re)built in the same launch as the lift
    operands; None: `plan` is used as it is)."""
    if build is not None and not isinstance(plan, RunPlan):
        raise RuntimeError("lift_splat(build=...) needs a RunPlan")
    return _LiftSplatFn.apply(depthnet_out, prob, plan, mode, channels_last, bev_out, build)


class _VoxelPoolingFn(torch.autograd.Function):
    """Operator-level voxel_pooling(geom_feats, x) with a materialised x (models.py:204-246)."""

    @staticmethod
    def forward(ctx, x, prob, plan, mode, channels_last):
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("x must be a float32 CUDA tensor")
        strides = (C.c_int64 * 6)(*x.stride())
        bev = _empty_bev(prob, x.device, channels_last)
        check(lib().lss_voxel_pooling_fwd(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), _ptr(x), strides,
                                          _ptr(bev), SPLAT_MODES[mode],
                                          LAYOUT_CHANNELS_LAST if channels_last else LAYOUT_NCHW, 0, 0, _stream()),
              "lss_voxel_pooling_fwd")
        ctx.prob, ctx.plan, ctx.xshape = prob, plan, tuple(x.shape)
        return bev

    @staticmethod
    def backward(ctx, grad_bev):
        prob, plan = ctx.prob, ctx.plan
        g, layout = _bev_layout(_f32c_keep(grad_bev))
        rows = None
        if layout == LAYOUT_NCHW:
            rows = torch.empty((max(prob.n_voxels, plan.layout.n_rows_cap), prob.C), dtype=torch.float32, device=g.device)
        gx = torch.empty(ctx.xshape, dtype=torch.float32, device=g.device)
        check(lib().lss_voxel_pooling_bwd(C.byref(prob.c), C.byref(plan.layout), _ptr(plan.ws), _ptr(g), layout,
                                          _ptr(rows), _ptr(gx), _stream()), "lss_voxel_pooling_bwd")
        return gx, None, None, None, None


def voxel_pooling(geom_feats, x, dx, bx, nx, mode="sorted", channels_last=False, plan=None):
    """Drop-in for LiftSplatShoot.voxel_pooling(geom_feats, x) (models.py:204-246).

    geom_feats f32[B,N,D,H,W,3], x f32[B,N,D,H,W,C] (any strides) -> f32[B, nz*C, nx, ny]."""
    B, N, D, H, W, Cc = x.shape
    prob = Problem.from_grid(B, N, D, H, W, Cc, dx, bx, nx)
    plan = build_plan(prob, geom=geom_feats, sorted=(mode == "sorted"), plan=plan)
    return _VoxelPoolingFn.apply(x, prob, plan, mode, channels_last)


# ------------------------------------------------------------------------------------------------
# QuickCumsum / cumsum_trick (tools.py:182-219)
# ------------------------------------------------------------------------------------------------

class QuickCumsum(torch.autograd.Function):
    """Same signature and return values as the reference's `QuickCumsum.apply(x, geom_feats, ranks)`:
    x f32[n, C] and geom_feats i64[n, 4] sorted by `ranks` i64[n] -> (sums f32[V, C], geom i64[V, 4]),
    V = number of distinct ranks; each run keeps the coordinates of its LAST point (tools.py:198-200)."""

    @staticmethod
    def forward(ctx, x, geom_feats, ranks):
        if not x.is_cuda:
            raise RuntimeError("QuickCumsum: CUDA tensors required (no CPU implementation)")
        n, Cc = x.shape
        x = x.float()
        if x.stride(1) != 1:
            x = x.contiguous()
        geom_feats = geom_feats.contiguous().long()
        ranks = ranks.contiguous().long()
        L = lib()
        run_id = torch.empty(max(n, 1), dtype=torch.int32, device=x.device)
        n_runs = torch.zeros(1, dtype=torch.int32, device=x.device)
        scratch = torch.empty(L.lss_quickcumsum_scratch_elems(n), dtype=torch.int32, device=x.device)
        check(L.lss_quickcumsum_runs(n, _ptr(ranks), _ptr(run_id), _ptr(n_runs), _ptr(scratch), _stream()),
              "lss_quickcumsum_runs")
        V = int(n_runs.item())      # data-dependent output size: same sync as the reference's x[kept]
        sums = torch.empty((V, Cc), dtype=torch.float32, device=x.device)
        gout = torch.empty((V, geom_feats.shape[1]), dtype=torch.int64, device=x.device)
        if geom_feats.shape[1] != 4:
            raise ValueError("geom_feats must have 4 columns (ix, iy, iz, b)")
        check(L.lss_quickcumsum_fwd(n, Cc, _ptr(x), x.stride(0), _ptr(geom_feats), _ptr(scratch), V, _ptr(sums),
                                    _ptr(gout), _stream()), "lss_quickcumsum_fwd")
        ctx.save_for_backward(run_id)
        ctx.n, ctx.Cc = n, Cc
        ctx.mark_non_differentiable(gout)
        return sums, gout

    @staticmethod
    def backward(ctx, gradx, gradgeom):
        run_id, = ctx.saved_tensors
        g = gradx.contiguous().float()
        out = torch.empty((ctx.n, ctx.Cc), dtype=torch.float32, device=g.device)
        check(lib().lss_quickcumsum_bwd(ctx.n, ctx.Cc, _ptr(g), _ptr(run_id), _ptr(out), _stream()),
              "lss_quickcumsum_bwd")
        return out, None, None


def cumsum_trick(x, geom_feats, ranks):
    """tools.py:182-190.  Same kernels as QuickCumsum (the reference's two variants differ only in how
    autograd derives the backward; the gather backward is exact for both)."""
    return QuickCumsum.apply(x, geom_feats, ranks)
