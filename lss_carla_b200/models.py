"""API-compatible `LiftSplatShoot` whose lift-splat path runs on the B200 CUDA library.

Mirrors the public surface of the reference's `src/models.py` for the path:

    compile_model(grid_conf, data_aug_conf, outC)                      models.py:262-263
    LiftSplatShoot.{dx,bx,nx,frustum}  (no-grad Parameters, same state_dict keys)   models.py:143-145,168
    LiftSplatShoot.create_frustum / get_geometry / get_cam_feats / voxel_pooling / get_voxels / forward
                                                                        models.py:157-259
    LiftSplatShoot.use_quickcumsum                                      models.py:155

plus `install(model)`, which rebinds the three hot methods of an EXISTING reference instance -- the
drop-in inside `src/models.py` -- leaving its trunk, parameters and state_dict untouched.

The EfficientNet trunk and BevEncode stay in PyTorch (north_star); see `trunk.py` for the offline
stand-in of `efficientnet_pytorch`.
"""
from __future__ import annotations

import types

import torch
from torch import nn

from . import ops
from .tools import gen_dx_bx

INVERSE_MODES = ("reference", "device")


def _fused_get_voxels(self, x, rots, trans, intrins, post_rots, post_trans):
    """get_voxels (models.py:248-254) without materialising geometry or frustum features."""
    B, N, _, imH, imW = x.shape
    ce = self.camencode
    feat = ce.get_eff_depth(x.view(B * N, x.shape[2], imH, imW))          # PyTorch trunk (models.py:53)
    dn = ce.depthnet(ce.dropout(feat))                                     # models.py:55-56
    return lift_splat_from_depthnet(self, dn, rots, trans, intrins, post_rots, post_trans)


def lift_splat_from_depthnet(model, depthnet_out, rots, trans, intrins, post_rots, post_trans, plan=None):
    """Geometry + lift + splat for a depthnet output [B*N, D+C, fH, fW] -> BEV [B, nz*C, nx, ny].

    `bev_channels_last` + `splat_mode="sorted"` take the run-plan path (ops.RunPlan, ops.liftsplat_forward): the BEV is
    zero-filled by the bulk-copy engine while the plan is built and the voxels are summed; every non-empty voxel row is
    written once."""
    B, N = trans.shape[:2]
    fH, fW = depthnet_out.shape[-2:]
    C = depthnet_out.shape[1] - model.D
    prob = _problem_for(model, B, N, fH, fW, C)
    mode = getattr(model, "splat_mode", "sorted")
    if plan is None and _use_runplan(model, prob):
        ws = _cached_plan(model, prob, depthnet_out.device, run=True)
        build = runplan_build_args(model, prob, rots, trans, intrins, post_rots, post_trans)
        return ops.lift_splat(depthnet_out, prob, ws, mode, True, build=build)     # ops.liftsplat_forward: plan + lift + forward
    if plan is None:
        plan = plan_from_calibration(model, prob, rots, trans, intrins, post_rots, post_trans)
    return ops.lift_splat(depthnet_out, prob, plan, mode, getattr(model, "bev_channels_last", False))


def _use_runplan(model, prob):
    return (getattr(model, "bev_channels_last", False) and getattr(model, "splat_mode", "sorted") == "sorted"
            and ops.runplan_supported(prob))


def runplan_build_args(model, prob, rots, trans, intrins, post_rots, post_trans):
    """Calibration arguments of ops.build_runplan / ops.liftsplat_prologue for the batch of `forward` (models.py:256).
    inverse_mode "reference": M1/M2 from the reference's own torch calls (bit-identical geometry); "device": closed-form
    inverses inside the kernel, no host round trip (tiny cameras, where a CTA would span too many of them: separate kernel)."""
    args = dict(frustum=model.frustum.detach(), trans=trans.reshape(-1, 3), post_trans=post_trans.reshape(-1, 3))
    if getattr(model, "inverse_mode", "reference") == "device":
        if ops.runplan_raw_supported(prob):
            return dict(args, rots=rots, intrins=intrins, post_rots=post_rots)
        M1, M2 = ops.calib_matrices_device(rots, intrins, post_rots)
    else:
        M1, M2 = ops.calib_matrices_reference(rots, intrins, post_rots)
    return dict(args, M1=M1.reshape(-1, 3, 3), M2=M2.reshape(-1, 3, 3))


def runplan_from_calibration(model, prob, rots, trans, intrins, post_rots, post_trans):
    """Run plan (voxel row per point, sub-run lists per voxel) from the calibration of `forward` (models.py:256)."""
    ws = _cached_plan(model, prob, rots.device, run=True)
    return ops.build_runplan(prob, plan=ws, **runplan_build_args(model, prob, rots, trans, intrins, post_rots, post_trans))


def plan_from_calibration(model, prob, rots, trans, intrins, post_rots, post_trans):
    """Voxel ids + buckets for one batch from the calibration of `forward` (models.py:256).  inverse_mode "device"
    evaluates the 3x3 inverses inside the plan build (one launch less, no host round trip); "reference" prepares
    M1/M2 with the reference's own torch calls (bit-identical geometry)."""
    mode = getattr(model, "splat_mode", "sorted")
    ws = _cached_plan(model, prob, rots.device)
    frustum = model.frustum.detach()
    if getattr(model, "inverse_mode", "reference") == "device":
        try:
            return ops.build_plan_raw(prob, frustum, rots, trans, intrins, post_rots, post_trans, sorted=(mode == "sorted"), plan=ws)
        except RuntimeError as e:                  # tiny cameras (LSS_ERR_UNSUPPORTED): separate calibration kernel
            if "status -3" not in str(e):
                raise
    M1, M2 = _calib_matrices(model, rots, intrins, post_rots)
    calib = (frustum, post_trans.reshape(-1, 3), M1.reshape(-1, 3, 3), M2.reshape(-1, 3, 3), trans.reshape(-1, 3))
    return ops.build_plan(prob, calib=calib, sorted=(mode == "sorted"), plan=ws)


def _calib_matrices(model, rots, intrins, post_rots):
    if getattr(model, "inverse_mode", "reference") == "device":
        return ops.calib_matrices_device(rots, intrins, post_rots)
    return ops.calib_matrices_reference(rots, intrins, post_rots)


def _problem_for(model, B, N, fH, fW, C):
    key = (B, N, fH, fW, C)
    cache = model.__dict__.setdefault("_lss_problems", {})
    if key not in cache:
        cache[key] = ops.Problem.from_grid(B, N, model.D, fH, fW, C, model.dx, model.bx, model.nx)
    return cache[key]


def _cached_plan(model, prob, device, run=False):
    """Small pool of reusable plan workspaces per (problem, device, kind).  A plan handed to a forward that
    needs gradients stays busy until its backward ran or its graph died (`ops._LiftSplatFn`); if every
    pooled plan is busy a fresh one is allocated."""
    pool = model.__dict__.setdefault("_lss_plans", {})
    lst = pool.setdefault((id(prob), str(device), run), [])
    for pl in lst:
        if not pl.busy:
            return pl
    pl = ops.RunPlan(prob, device) if run else ops.Plan(prob, device, getattr(model, "tile_cols", 0))
    if len(lst) < 4:
        lst.append(pl)
    return pl


def _get_geometry(self, rots, trans, intrins, post_rots, post_trans):
    """get_geometry (models.py:170-190) -> f32[B,N,D,fH,fW,3] on the CUDA path."""
    B, N, _ = trans.shape
    D, fH, fW, _ = self.frustum.shape
    prob = _problem_for(self, B, N, fH, fW, getattr(self, "camC", 64))
    M1, M2 = _calib_matrices(self, rots, intrins, post_rots)
    return ops.geometry(prob, self.frustum.detach(), post_trans.reshape(-1, 3), M1.reshape(-1, 3, 3),
                        M2.reshape(-1, 3, 3), trans.reshape(-1, 3))


def _voxel_pooling(self, geom_feats, x):
    """voxel_pooling (models.py:204-246) on the CUDA path; x may be any strided view."""
    mode = getattr(self, "splat_mode", "sorted")
    return ops.voxel_pooling(geom_feats, x, self.dx, self.bx, self.nx, mode=mode,
                             channels_last=getattr(self, "bev_channels_last", False))


def install(model, splat_mode="sorted", inverse_mode="reference", bev_channels_last=False, fused=True):
    """Rebind the lift-splat methods of an existing (reference) `LiftSplatShoot` instance to the CUDA path.

    The instance keeps its class, sub-modules, parameters and state_dict.  `use_quickcumsum` keeps its
    meaning as "which backward derivation" only: both settings run the same kernels (the gather backward
    is exact), so toggling it does not change results.

    `bev_channels_last=False` (default) returns the reference's NCHW-contiguous tensor (tile-plan kernels);
    `True` returns the same logical tensor in torch.channels_last strides through the sort-free run plan -- the
    faster path end to end on B200 (cuDNN runs `bevencode.conv1`, models.py:97-98, on NHWC data: BevEncode
    forward + backward 7.8 ms against 11.6 ms, profiles/r02_conv1_layout.json) and what bench.py measures.
    """
    if inverse_mode not in INVERSE_MODES:
        raise ValueError(f"inverse_mode must be one of {INVERSE_MODES}")
    if splat_mode not in ops.SPLAT_MODES:
        raise ValueError(f"splat_mode must be one of {tuple(ops.SPLAT_MODES)}")
    model.splat_mode, model.inverse_mode, model.bev_channels_last = splat_mode, inverse_mode, bev_channels_last
    model.get_geometry = types.MethodType(_get_geometry, model)
    model.voxel_pooling = types.MethodType(_voxel_pooling, model)
    if fused:
        model.get_voxels = types.MethodType(_fused_get_voxels, model)
    return model


class LiftSplatShoot(nn.Module):
    """Stand-alone API-compatible model (for environments where the reference cannot be imported)."""

    def __init__(self, grid_conf, data_aug_conf, outC, camencode=None, bevencode=None,
                 splat_mode="sorted", inverse_mode="reference", bev_channels_last=False):
        super().__init__()
        self.grid_conf = grid_conf
        self.data_aug_conf = data_aug_conf
        dx, bx, nx = gen_dx_bx(grid_conf["xbound"], grid_conf["ybound"], grid_conf["zbound"])
        self.dx = nn.Parameter(dx, requires_grad=False)
        self.bx = nn.Parameter(bx, requires_grad=False)
        self.nx = nn.Parameter(nx, requires_grad=False)
        self.downsample = 16
        self.camC = 64
        self.frustum = self.create_frustum()
        self.D = self.frustum.shape[0]
        if camencode is None or bevencode is None:
            from .trunk import BevEncode, CamEncode
            camencode = camencode or CamEncode(self.D, self.camC, self.downsample)
            bevencode = bevencode or BevEncode(inC=self.camC * int(nx[2]), outC=outC)
        self.camencode = camencode
        self.bevencode = bevencode
        self.use_quickcumsum = True
        self.splat_mode, self.inverse_mode, self.bev_channels_last = splat_mode, inverse_mode, bev_channels_last

    def create_frustum(self):
        """Pixel-centre / depth-bin lattice, f32[D, fH, fW, 3] = (x_px, y_px, depth).  Built once with the
        same torch constructors as the reference so that state_dicts are interchangeable bit for bit."""
        ogfH, ogfW = self.data_aug_conf["final_dim"]
        fH, fW = ogfH // self.downsample, ogfW // self.downsample
        ds = torch.arange(*self.grid_conf["dbound"], dtype=torch.float)
        xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float)
        ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float)
        D = ds.shape[0]
        fr = torch.empty(D, fH, fW, 3)
        fr[..., 0] = xs.view(1, 1, fW)
        fr[..., 1] = ys.view(1, fH, 1)
        fr[..., 2] = ds.view(D, 1, 1)
        return nn.Parameter(fr, requires_grad=False)

    get_geometry = _get_geometry
    voxel_pooling = _voxel_pooling
    get_voxels = _fused_get_voxels

    def get_cam_feats(self, x):
        """B x N x D x fH x fW x C lifted features, materialised (API compatibility; the fused
        `get_voxels` never calls this)."""
        B, N, Cin, imH, imW = x.shape
        x = self.camencode(x.view(B * N, Cin, imH, imW))
        x = x.view(B, N, self.camC, self.D, imH // self.downsample, imW // self.downsample)
        return x.permute(0, 1, 3, 4, 5, 2)

    def forward(self, x, rots, trans, intrins, post_rots, post_trans):
        x = self.get_voxels(x, rots, trans, intrins, post_rots, post_trans)
        return self.bevencode(x)


def compile_model(grid_conf, data_aug_conf, outC, **kw):
    return LiftSplatShoot(grid_conf, data_aug_conf, outC, **kw)
