"""Torch-CPU port of the reference's lift-splat path  --  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The reference is pure PyTorch; its hot path is a chain of stock ATen ops.  It
cannot travel to the GPU box (`/root/reference` is absent there), so this module restates that op
chain with the same ATen calls in the same order, function by function.  It serves two purposes:

  * `bench.py`'s `cpu_baseline` leg and `bench.py --impl reference` time it on the host cores
    (ATen's intra-op thread pool = all cores, like the reference itself would use);
  * `tests/test_oracle_golden.py` pins it bit-for-bit against fixtures produced by the real reference.

It is never imported by the product package.  Reference lines followed:
    geometry        src/models.py:170-190      lift     src/models.py:49-61, 192-202
    voxel_pooling   src/models.py:204-246      cumsum   src/tools.py:182-219
"""
from __future__ import annotations

import torch


class RunSum(torch.autograd.Function):
    """Segment sum by global prefix sum + difference, gather backward (tools.py:193-219)."""

    @staticmethod
    def forward(ctx, feats, coords, ranks):
        pref = feats.cumsum(0)
        last = torch.ones(pref.shape[0], device=pref.device, dtype=torch.bool)
        last[:-1] = ranks[1:] != ranks[:-1]
        pref, coords = pref[last], coords[last]
        sums = torch.cat((pref[:1], pref[1:] - pref[:-1]))
        ctx.save_for_backward(last)
        ctx.mark_non_differentiable(coords)
        return sums, coords

    @staticmethod
    def backward(ctx, g_sums, g_coords):
        last, = ctx.saved_tensors
        run = torch.cumsum(last, 0)
        run[last] -= 1
        return g_sums[run], None, None


def run_sum_autograd(feats, coords, ranks):
    """tools.py:182-190, the autograd-traced variant (use_quickcumsum=False)."""
    pref = feats.cumsum(0)
    last = torch.ones(pref.shape[0], device=pref.device, dtype=torch.bool)
    last[:-1] = ranks[1:] != ranks[:-1]
    pref, coords = pref[last], coords[last]
    return torch.cat((pref[:1], pref[1:] - pref[:-1])), coords


def _host_inverse(m):
    """`torch.inverse(m.cpu()).cuda()` of models.py:180,186: LAPACK on the host, result back on m's device."""
    return torch.inverse(m.cpu()).to(m.device)


def geometry(frustum, rots, trans, intrins, post_rots, post_trans):
    """models.py:170-190.  The two 3x3 inverses run on the host as in the reference (identity hop for
    CPU tensors), so the same function also times the stock-ATen path on a GPU."""
    B, N, _ = trans.shape
    pts = frustum - post_trans.view(B, N, 1, 1, 1, 3)
    pts = _host_inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1))
    pts = torch.cat((pts[..., :2, :] * pts[..., 2:3, :], pts[..., 2:3, :]), 5)
    combine = rots.matmul(_host_inverse(intrins))
    pts = combine.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1)
    pts += trans.view(B, N, 1, 1, 1, 3)
    return pts


def lift(depthnet_out, B, N, D, C):
    """models.py:58-59 and 199-200: returns the permuted VIEW [B,N,D,fH,fW,C] like get_cam_feats."""
    depth = depthnet_out[:, :D].softmax(dim=1)
    feat = depth.unsqueeze(1) * depthnet_out[:, D:D + C].unsqueeze(2)
    fH, fW = depthnet_out.shape[-2:]
    return feat.view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)


def voxel_pooling(geom, x, dx, bx, nx, quick=True):
    """models.py:204-246."""
    B, N, D, H, W, C = x.shape
    Np = B * N * D * H * W
    x = x.reshape(Np, C)
    g = ((geom - (bx - dx / 2.)) / dx).long().view(Np, 3)
    bix = torch.cat([torch.full([Np // B, 1], i, device=x.device, dtype=torch.long) for i in range(B)])
    g = torch.cat((g, bix), 1)
    keep = (g[:, 0] >= 0) & (g[:, 0] < nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < nx[1]) \
        & (g[:, 2] >= 0) & (g[:, 2] < nx[2])
    x, g = x[keep], g[keep]
    ranks = g[:, 0] * (nx[1] * nx[2] * B) + g[:, 1] * (nx[2] * B) + g[:, 2] * B + g[:, 3]
    order = ranks.argsort()
    x, g, ranks = x[order], g[order], ranks[order]
    x, g = RunSum.apply(x, g, ranks) if quick else run_sum_autograd(x, g, ranks)
    out = torch.zeros((B, C, int(nx[2]), int(nx[0]), int(nx[1])), device=x.device)
    out[g[:, 3], :, g[:, 2], g[:, 0], g[:, 1]] = x
    return torch.cat(out.unbind(dim=2), 1)


def liftsplat_forward(depthnet_out, frustum, calib, dx, bx, nx, C, quick=True):
    """models.py:248-254 minus the camera trunk: geometry -> lift -> voxel_pooling."""
    B, N = calib["trans"].shape[:2]
    D = frustum.shape[0]
    geom = geometry(frustum, calib["rots"], calib["trans"], calib["intrins"],
                    calib["post_rots"], calib["post_trans"])
    x = lift(depthnet_out, B, N, D, C)
    return voxel_pooling(geom, x, dx, bx, nx, quick=quick)


def liftsplat_step(depthnet_out, frustum, calib, dx, bx, nx, C, grad_bev):
    """One forward + backward of the path: returns (bev, grad w.r.t. depthnet_out)."""
    inp = depthnet_out.detach().requires_grad_(True)
    bev = liftsplat_forward(inp, frustum, calib, dx, bx, nx, C)
    bev.backward(grad_bev)
    return bev.detach(), inp.grad
