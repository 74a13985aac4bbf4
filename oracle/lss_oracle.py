"""CPU oracle for the lift-splat hot path of shdragron/LSS-Carla  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's algorithm for the path named by
BASELINE.json `north_star`.  It is imported only by `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu-baseline / `--impl reference` legs, always as the
checker or the timed baseline -- never by the product package `lss_carla_b200`, which has no CPU
fallback and raises when the CUDA library is missing.

Parity status: PINNED AGAINST THE REFERENCE RUN HERE.  The reference ships no tests, golden
vectors or known-answer fixtures for this path, so the oracle is pinned against
outputs of the reference's own code executed in the build container (`tests/golden/make_golden.py`
imports the real `src/models.py` / `src/tools.py` with third-party stubs and writes
`tests/golden/*.npz`); `tests/test_oracle_golden.py` checks every function below against them.

Reference lines restated (paths relative to the reference root):
    gen_dx_bx                  src/tools.py:174-179
    create_frustum             src/models.py:157-168
    geometry                   src/models.py:170-190
    lift (softmax (x) context) src/models.py:49-61, layout src/models.py:192-202
    voxel index / kept / rank  src/models.py:204-231
    cumsum trick fwd / bwd     src/tools.py:182-219
    griddify + collapse Z      src/models.py:239-244

Arithmetic conventions that make integer results bit-exact (measured against the reference on CPU):
  * all per-point arithmetic is IEEE binary32, one rounding per operation, no FMA contraction;
  * the two 3x3 * 3x1 products (models.py:180,187) evaluate each row as (a0*v0 + a1*v1) + a2*v2;
  * `torch.linspace` (float32) uses a float32 step=(end-start)/(steps-1), fills the upper half from
    `end`, and fuses the multiply-add (one rounding per element);
  * voxel index = C-style truncation toward zero of (geom - (bx - dx/2)) / dx (models.py:212);
  * `argsort` ties keep ascending flat (b,n,d,h,w) order (stable), SURVEY.md section 7.3 H3.  Measured:
    ATen's CPU `argsort()` is stable (radix path) for >= ~32k int64 keys, i.e. every BASELINE config;
    for smaller inputs its tie order is an unspecified permutation, so small fixtures are compared
    as sorted rank sequence + per-voxel point sets;
  * the CPU `cumsum` of the reference accumulates float32 inputs in float64 and rounds every prefix to
    float32 (measured: bit-identical to that on 100 000 x 8 random inputs).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ------------------------------------------------------------------------------------------------
# grid constants and frustum
# ------------------------------------------------------------------------------------------------

def gen_dx_bx(xbound, ybound, zbound):
    """tools.py:174-179.  dx = step, bx = first bin centre, nx = bin count (float -> int64 truncation)."""
    rows = (xbound, ybound, zbound)
    dx = np.array([r[2] for r in rows], dtype=np.float64).astype(F32)
    bx = np.array([r[0] + r[2] / 2.0 for r in rows], dtype=np.float64).astype(F32)
    nx = np.array([int((r[1] - r[0]) / r[2]) for r in rows], dtype=np.int64)
    return dx, bx, nx


def torch_linspace_f32(start, end, steps):
    """float32 `torch.linspace(start, end, steps)` as ATen's vectorised CPU kernel evaluates it: a
    float32 step = (end-start)/(steps-1); the lower half is fma(step, i, start), the upper half is
    fma(-step, steps-1-i, end) -- a single rounding per element (measured: the un-fused form differs in
    1-5 elements for 30..100 steps).  step*i is exact in float64, so one float64 -> float32 rounding
    reproduces the fused result.  Pinned by tests/golden/constants.npz."""
    start, end = F32(start), F32(end)
    if steps == 1:
        return np.array([start], dtype=F32)
    step = np.float64(F32((end - start) / F32(steps - 1)))
    i = np.arange(steps, dtype=np.float64)
    lo = np.float64(start) + step * i
    hi = np.float64(end) - step * (steps - 1 - i)
    return np.where(i < steps // 2, lo, hi).astype(F32)


def torch_arange_f32(start, end, step):
    """float32 `torch.arange(start, end, step, dtype=torch.float)`: length ceil((end-start)/step) in
    double, element i = float32(start + i*step) evaluated in double."""
    n = int(np.ceil((float(end) - float(start)) / float(step)))
    return (float(start) + np.arange(n, dtype=np.float64) * float(step)).astype(F32)


def create_frustum(final_dim, dbound, downsample=16):
    """models.py:157-168 -> f32[D, fH, fW, 3] holding (x_pixel, y_pixel, depth)."""
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample, ogfW // downsample
    ds = torch_arange_f32(*dbound)
    xs = torch_linspace_f32(0, ogfW - 1, fW)
    ys = torch_linspace_f32(0, ogfH - 1, fH)
    D = ds.shape[0]
    fr = np.empty((D, fH, fW, 3), dtype=F32)
    fr[..., 0] = xs[None, None, :]
    fr[..., 1] = ys[None, :, None]
    fr[..., 2] = ds[:, None, None]
    return fr


# ------------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------------

def calib_matrices_torch(rots, intrins, post_rots):
    """The two host-side matrix preparations of models.py:180,186, through the same third-party calls
    the reference makes (`torch.inverse` = LAPACK on the CPU, `Tensor.matmul`).  The LU inverse is
    not restated: its bits depend on the LAPACK build, and every function below takes M1/M2 as
    inputs so that the per-point arithmetic is pinned independently of it.
    Returns M1 = inverse(post_rots), M2 = rots @ inverse(intrins), float32 [B,N,3,3]."""
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F32))
    M1 = torch.inverse(t(post_rots))
    M2 = t(rots).matmul(torch.inverse(t(intrins)))
    return M1.numpy(), M2.numpy()


def _matvec_unfused(M, v):
    """Row-wise (a0*v0 + a1*v1) + a2*v2 in float32; M [B,N,3,3] broadcast over v [B,N,D,H,W,3]."""
    M = M[:, :, None, None, None]
    out = np.empty_like(v)
    for i in range(3):
        t = M[..., i, 0] * v[..., 0]
        t = t + M[..., i, 1] * v[..., 1]
        out[..., i] = t + M[..., i, 2] * v[..., 2]
    return out


def geometry(frustum, post_trans, M1, M2, trans):
    """models.py:170-190 given the prepared matrices.  -> f32[B,N,D,fH,fW,3] ego-frame xyz."""
    frustum = np.asarray(frustum, F32)
    post_trans, trans = np.asarray(post_trans, F32), np.asarray(trans, F32)
    M1, M2 = np.asarray(M1, F32), np.asarray(M2, F32)
    p = frustum[None, None] - post_trans[:, :, None, None, None, :]          # models.py:179
    q = _matvec_unfused(M1, p)                                              # models.py:180
    r = np.empty_like(q)                                                    # models.py:183-185
    r[..., 0] = q[..., 0] * q[..., 2]
    r[..., 1] = q[..., 1] * q[..., 2]
    r[..., 2] = q[..., 2]
    s = _matvec_unfused(M2, r)                                              # models.py:187
    s = s + trans[:, :, None, None, None, :]                                # models.py:188
    return s.astype(F32, copy=False)


# ------------------------------------------------------------------------------------------------
# voxel index, kept mask, ranks, sort
# ------------------------------------------------------------------------------------------------

def voxel_index(geom, dx, bx, nx):
    """models.py:212-221.  Returns (idx int64[Nprime,3], kept bool[Nprime]).

    Truncation toward zero like `Tensor.long()`.  Non-finite or out-of-int64-range quotients convert
    to INT64_MIN on x86 (`cvttss2si` indefinite value), i.e. are never kept."""
    dx, bx = np.asarray(dx, F32), np.asarray(bx, F32)
    lo = bx - dx / F32(2.0)
    u = ((np.asarray(geom, F32) - lo) / dx).reshape(-1, 3)
    bad = ~np.isfinite(u) | (np.abs(u) >= F32(2.0 ** 63))
    with np.errstate(invalid="ignore"):
        idx = np.trunc(np.where(bad, 0, u)).astype(np.int64)
    idx[bad] = np.iinfo(np.int64).min
    nx = np.asarray(nx, np.int64)
    kept = np.all((idx >= 0) & (idx < nx[None, :]), axis=1)
    return idx, kept


def ranks_and_sort(idx, kept, B, nx):
    """models.py:214-231.  Returns dict with
        flat      int64[Nk]   flat (b,n,d,h,w) index of each kept point, in compacted order
        geom4     int64[Nk,4] (ix, iy, iz, b) of the kept points (models.py:216,223)
        ranks     int64[Nk]   models.py:226-229 (x-major, batch fastest)
        sorts     int64[Nk]   stable argsort of ranks (indices into the compacted arrays)
    """
    Nprime = idx.shape[0]
    batch_ix = np.repeat(np.arange(B, dtype=np.int64), Nprime // B)
    flat = np.nonzero(kept)[0].astype(np.int64)
    g4 = np.concatenate([idx[flat], batch_ix[flat, None]], axis=1)
    nx = np.asarray(nx, np.int64)
    ranks = g4[:, 0] * (nx[1] * nx[2] * B) + g4[:, 1] * (nx[2] * B) + g4[:, 2] * B + g4[:, 3]
    sorts = np.argsort(ranks, kind="stable")
    return {"flat": flat, "geom4": g4, "ranks": ranks, "sorts": sorts}


# ------------------------------------------------------------------------------------------------
# lift
# ------------------------------------------------------------------------------------------------

def depth_softmax(depthnet_out, D):
    """models.py:49-50,58: softmax over the first D channels.  float32 in, float32 out; the exponent
    is evaluated in float64 and rounded once, so this is the *correctly rounded* softmax the CUDA
    and torch results are compared against with a tolerance (they differ in the last ulps)."""
    z = np.asarray(depthnet_out[:, :D], np.float64)
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(F32)


def lift(depthnet_out, B, N, D, C, prob=None):
    """models.py:58-59 + 199-200: frustum features x[B,N,D,fH,fW,C] = prob (x) context (float32
    product, one rounding).  Only for sizes where that tensor is affordable."""
    BN, _, fH, fW = depthnet_out.shape
    if prob is None:
        prob = depth_softmax(depthnet_out, D)
    ctx = np.asarray(depthnet_out[:, D:D + C], F32)
    x = prob[:, None] * ctx[:, :, None]                  # [BN, C, D, fH, fW]
    return np.ascontiguousarray(x.reshape(B, N, C, D, fH, fW).transpose(0, 1, 3, 4, 5, 2))


# ------------------------------------------------------------------------------------------------
# splat
# ------------------------------------------------------------------------------------------------

def cumsum_trick(x_sorted, geom_sorted, ranks_sorted):
    """tools.py:182-201 (forward of `cumsum_trick` and `QuickCumsum`): global prefix sum, keep the
    last point of every run of equal rank, first-difference.  Prefix accumulated in float64 and
    rounded to float32 per element, as the reference's CPU `cumsum` does."""
    x_sorted = np.asarray(x_sorted, F32)
    n = x_sorted.shape[0]
    if n == 0:
        return x_sorted.copy(), geom_sorted.copy(), np.zeros(0, bool)
    pref = np.cumsum(x_sorted.astype(np.float64), axis=0).astype(F32)
    kept = np.ones(n, dtype=bool)
    kept[:-1] = ranks_sorted[1:] != ranks_sorted[:-1]
    xs, gs = pref[kept], geom_sorted[kept]
    out = np.concatenate([xs[:1], xs[1:] - xs[:-1]], axis=0)
    return out, gs, kept


def quickcumsum_backward(grad_out, kept):
    """tools.py:212-219: every sorted point receives the gradient row of its run (pure gather)."""
    back = np.cumsum(kept.astype(np.int64))
    back[kept] -= 1
    return grad_out[back]


def griddify(vox_feats, geom4, B, C, nx):
    """models.py:240-244: scatter run sums to (B, C, Z, X, Y), then fold Z into channels (z*C + c)."""
    X, Y, Z = (int(v) for v in nx)
    final = np.zeros((B, C, Z, X, Y), dtype=F32)
    final[geom4[:, 3], :, geom4[:, 2], geom4[:, 0], geom4[:, 1]] = vox_feats
    return np.ascontiguousarray(final.reshape(B, C * Z, X, Y)) if Z == 1 else \
        np.ascontiguousarray(final.transpose(0, 2, 1, 3, 4).reshape(B, Z * C, X, Y))


def voxel_pooling_reference(geom, x, dx, bx, nx):
    """models.py:204-246 end to end with the cumsum trick -> f32[B, Z*C, X, Y]."""
    B, N, D, H, W, C = x.shape
    idx, kept = voxel_index(geom, dx, bx, nx)
    rs = ranks_and_sort(idx, kept, B, nx)
    xk = x.reshape(-1, C)[rs["flat"]][rs["sorts"]]
    g4 = rs["geom4"][rs["sorts"]]
    rk = rs["ranks"][rs["sorts"]]
    vf, vg, _ = cumsum_trick(xk, g4, rk)
    return griddify(vf, vg, B, C, nx)


def voxel_linear_id(idx, kept, B, nx):
    """Dense voxel id used by the sequential/exact splats: ((b*Z + iz)*X + ix)*Y + iy, -1 if dropped."""
    Nprime = idx.shape[0]
    X, Y, Z = (int(v) for v in nx)
    b = np.repeat(np.arange(B, dtype=np.int64), Nprime // B)
    safe = np.where(kept[:, None], idx, 0)
    v = ((b * Z + safe[:, 2]) * X + safe[:, 0]) * Y + safe[:, 1]
    return np.where(kept, v, -1)


def splat_from_prob(prob, ctx_t, vox, B, C, nx, dtype=F32):
    """The deterministic segmented sum the CUDA `sorted` mode implements, stated without any sort:

        bev[v, c] = sum over kept points p with voxel v, in ASCENDING flat index p, of
                    float32(prob[p] * ctx[pixel(p), c])            (one rounding for the product)
        accumulated sequentially in `dtype` (float32: one rounding per add; float64: "truth").

    prob  f32[B*N, D, fH, fW] (flat index == point index), ctx_t f32[B*N, fH*fW, C], vox int64[Nprime].
    Returns f32[B, Z*C, X, Y].  `np.add.at` applies the updates one by one in index order."""
    X, Y, Z = (int(v) for v in nx)
    BN, D, fH, fW = prob.shape
    p = np.nonzero(vox >= 0)[0]
    bn = p // (D * fH * fW)
    hw = p % (fH * fW)
    contrib = (prob.reshape(-1)[p, None] * ctx_t[bn, hw, :]).astype(F32)   # float32 product
    acc = np.zeros((B * Z * X * Y, C), dtype=dtype)
    np.add.at(acc, vox[p], contrib.astype(dtype))
    acc = acc.astype(F32).reshape(B, Z, X, Y, C)
    return np.ascontiguousarray(acc.transpose(0, 1, 4, 2, 3).reshape(B, Z * C, X, Y))


def liftsplat_backward(grad_bev, depthnet_out, prob, vox, B, N, D, C, nx, dtype=np.float64):
    """Gradient of sum(bev * grad_bev) w.r.t. the depthnet output (logits and context), i.e. the
    composition of: griddify/QuickCumsum backward = gather (tools.py:212-219), outer-product backward
    (models.py:59) and softmax backward (models.py:50,58).  Evaluated in `dtype`, returned as float32.

        g[p, :]      = grad_bev[b, iz*C:(iz+1)*C, ix, iy]       for kept p, else 0
        gp[p]        = sum_c g[p, c] * ctx[pixel(p), c]
        d_ctx[pix,c] = sum_d prob[pix, d] * g[(pix, d), c]
        d_logit[p]   = prob[p] * (gp[p] - sum_d' prob[pix, d'] * gp[(pix, d')])
    """
    X, Y, Z = (int(v) for v in nx)
    BN, _, fH, fW = depthnet_out.shape
    HW = fH * fW
    gb = np.asarray(grad_bev, dtype).reshape(B, Z, C, X, Y).transpose(0, 1, 3, 4, 2).reshape(-1, C)
    ctx = np.asarray(depthnet_out[:, D:D + C], dtype).reshape(BN, C, HW)
    pr = np.asarray(prob, dtype).reshape(BN, D, HW)
    voxr = vox.reshape(BN, D, HW)
    gp = np.zeros((BN, D, HW), dtype)
    dctx = np.zeros((BN, C, HW), dtype)
    for d in range(D):
        v = voxr[:, d, :]                                  # [BN, HW]
        k = v >= 0
        g = np.where(k[..., None], gb[np.where(k, v, 0)], 0)   # [BN, HW, C]
        gp[:, d, :] = np.einsum("bpc,bcp->bp", g, ctx)
        dctx += pr[:, d, None, :] * g.transpose(0, 2, 1)
    dlogit = pr * (gp - (pr * gp).sum(axis=1, keepdims=True))
    out = np.zeros(depthnet_out.shape, F32)
    out[:, :D] = dlogit.reshape(BN, D, fH, fW).astype(F32)
    out[:, D:D + C] = dctx.reshape(BN, C, fH, fW).astype(F32)
    return out


def ctx_transposed(depthnet_out, D, C):
    """Context channels as [B*N, fH*fW, C] (pixel-major), the layout the splat gathers from."""
    BN, _, fH, fW = depthnet_out.shape
    return np.ascontiguousarray(
        np.asarray(depthnet_out[:, D:D + C], F32).reshape(BN, C, fH * fW).transpose(0, 2, 1))


def liftsplat_forward(depthnet_out, frustum, calib, dx, bx, nx, C, M1=None, M2=None, dtype=F32):
    """Whole path (models.py:248-254 minus the trunk): geometry -> voxel ids -> lift -> sequential
    per-voxel sum.  `calib` = dict(rots, trans, intrins, post_rots, post_trans) of numpy arrays.
    Returns (bev f32[B,Z*C,X,Y], aux dict)."""
    B, N = calib["trans"].shape[:2]
    D = frustum.shape[0]
    if M1 is None or M2 is None:
        M1, M2 = calib_matrices_torch(calib["rots"], calib["intrins"], calib["post_rots"])
    geom = geometry(frustum, calib["post_trans"], M1, M2, calib["trans"])
    idx, kept = voxel_index(geom, dx, bx, nx)
    vox = voxel_linear_id(idx, kept, B, nx)
    prob = depth_softmax(depthnet_out, D)
    ctx_t = ctx_transposed(depthnet_out, D, C)
    bev = splat_from_prob(prob, ctx_t, vox, B, C, nx, dtype=dtype)
    return bev, {"geom": geom, "idx": idx, "kept": kept, "vox": vox, "prob": prob, "M1": M1, "M2": M2}
