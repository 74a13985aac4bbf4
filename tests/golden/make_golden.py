"""Generate tests/golden/*.npz by executing the REAL reference in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only mount)

The reference ships no golden vectors for the lift-splat path, so these fixtures
are produced by importing its own `src/models.py` / `src/tools.py` (third-party imports stubbed,
see `_ref_import.py`) and running the unmodified methods on seeded synthetic inputs:

    LiftSplatShoot.__init__ / create_frustum / gen_dx_bx     models.py:134-168, tools.py:174-179
    LiftSplatShoot.get_geometry                                models.py:170-190
    CamEncode.get_depth_feat (trunk/depthnet = identity)       models.py:52-61
    LiftSplatShoot.get_cam_feats / voxel_pooling / get_voxels  models.py:192-254
    QuickCumsum / cumsum_trick (fwd + bwd through autograd)    tools.py:182-219

The camera trunk is not part of the path: `get_eff_depth`, `dropout` and `depthnet` are replaced by
the identity and `downsample` is set to 1 *after* construction so that `get_cam_feats` accepts a
depthnet-shaped tensor [B, N, D+C, fH, fW]; every line of lift/splat code that runs is the
reference's.  `voxel_pooling` does not return its integer intermediates (`geom_feats`, `kept`,
`ranks`, `sorts`), so `traced_voxel_pooling` re-evaluates those statements (models.py:212-231) with
the model's own `dx/bx/nx` parameters on the same inputs, next to the untouched method call.

Small cases store full tensors; cfg2/cfg4-sized cases store inputs + SHA-256 digests of the
integer / geometry arrays plus sampled values, to keep the repository small.
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

import _ref_import as R  # noqa: E402
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_model(models, cfg):
    m = models.LiftSplatShoot(cfg.grid_conf, cfg.data_aug_conf, outC=1)
    m.camencode.get_eff_depth = lambda x: x          # trunk is out of scope
    m.camencode.dropout = torch.nn.Identity()
    m.camencode.depthnet = torch.nn.Identity()
    m.downsample = 1                                 # frustum already built with 16
    m.camC = cfg.C                                   # models.py:148 hard-codes 64; attribute only
    m.camencode.C = cfg.C
    return m


def traced_voxel_pooling(model, geom, x):
    """Run the reference `voxel_pooling` and recover its integer intermediates by replaying the public,
    deterministic statements on the same inputs with the reference's own parameters (models.py:212-231).
    The BEV output itself comes from the untouched method."""
    out = model.voxel_pooling(geom, x)
    B = x.shape[0]
    Np = geom.numel() // 3
    g = ((geom - (model.bx - model.dx / 2.)) / model.dx).long().view(Np, 3)
    bix = torch.arange(B).repeat_interleave(Np // B).view(Np, 1)
    g4 = torch.cat((g, bix), 1)
    kept = (g4[:, 0] >= 0) & (g4[:, 0] < model.nx[0]) & (g4[:, 1] >= 0) & (g4[:, 1] < model.nx[1]) \
        & (g4[:, 2] >= 0) & (g4[:, 2] < model.nx[2])
    gk = g4[kept]
    ranks = gk[:, 0] * (model.nx[1] * model.nx[2] * B) + gk[:, 1] * (model.nx[2] * B) + gk[:, 2] * B + gk[:, 3]
    sorts = ranks.argsort()
    return out, g, kept, ranks, sorts


def run_case(models, tools, cfg_name, aug, seed, full):
    cfg = CONFIGS[cfg_name]
    m = build_model(models, cfg)
    batch = make_batch(cfg, seed, aug)
    B, N = cfg.B, cfg.N
    fH, fW = cfg.fHW
    D, C = cfg.D, cfg.C
    assert m.D == D, (m.D, D)
    calib = [batch[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans")]
    geom = R.reference_geometry_cpu(m, *calib).detach()
    M1 = torch.inverse(batch["post_rots"])
    M2 = batch["rots"].matmul(torch.inverse(batch["intrins"]))

    x_in = batch["depthnet_out"].view(B, N, D + C, fH, fW).clone().requires_grad_(True)
    feats = m.get_cam_feats(x_in)                                   # real lift (models.py:58-59,192-202)
    bev, g3, kept, ranks, sorts = traced_voxel_pooling(m, geom, feats)
    gbev = make_bev_grad(cfg, seed)
    bev.backward(gbev)
    grad_quick = x_in.grad.detach().clone()

    # autograd variant (use_quickcumsum = False, models.py:234-235)
    m.use_quickcumsum = False
    x_in2 = batch["depthnet_out"].view(B, N, D + C, fH, fW).clone().requires_grad_(True)
    bev2 = m.voxel_pooling(geom, m.get_cam_feats(x_in2))
    bev2.backward(gbev)
    m.use_quickcumsum = True

    out = {
        "cfg": np.array(cfg_name), "aug": np.array(aug), "seed": np.array(seed),
        "rots": batch["rots"].numpy(), "trans": batch["trans"].numpy(), "intrins": batch["intrins"].numpy(),
        "post_rots": batch["post_rots"].numpy(), "post_trans": batch["post_trans"].numpy(),
        "M1": M1.numpy(), "M2": M2.numpy(),
        "dx": m.dx.detach().numpy(), "bx": m.bx.detach().numpy(), "nx": m.nx.detach().numpy(),
        "frustum": m.frustum.detach().numpy(),
        "n_points": np.array(geom.numel() // 3), "n_kept": np.array(int(kept.sum())),
        "n_voxels_hit": np.array(int(torch.unique(ranks).numel())),
        "sha_geom": np.array(sha(geom.numpy())), "sha_idx": np.array(sha(g3.numpy())),
        "sha_kept": np.array(sha(kept.numpy())), "sha_ranks": np.array(sha(ranks.numpy())),
        "sha_sorts": np.array(sha(sorts.numpy())),
        "bev_quick_equals_autograd": np.array(bool(torch.equal(bev, bev2))),
        "grad_quick_vs_autograd_maxabs": np.array(float((grad_quick - x_in2.grad).abs().max())),
    }
    # sampled values (every case)
    rs = np.random.RandomState(7)
    pick = rs.choice(geom.numel() // 3, size=min(2048, geom.numel() // 3), replace=False)
    out["sample_points"] = pick.astype(np.int64)
    out["sample_geom"] = geom.reshape(-1, 3).numpy()[pick]
    out["sample_idx"] = g3.numpy()[pick]
    out["sample_kept"] = kept.numpy()[pick]
    bev_np = bev.detach().numpy()
    out["bev_sum"] = np.array(bev_np.astype(np.float64).sum())
    out["bev_abs_sum"] = np.array(np.abs(bev_np.astype(np.float64)).sum())
    if full:
        out["depthnet_out"] = batch["depthnet_out"].numpy()
        out["grad_bev_seed"] = np.array(seed)
        out["geom"] = geom.numpy()
        out["idx"] = g3.numpy().astype(np.int32) if int(g3.abs().max()) < 2 ** 31 else g3.numpy()
        out["kept"] = kept.numpy()
        out["ranks"] = ranks.numpy()
        out["sorts"] = sorts.numpy().astype(np.int32)
        nz = np.argwhere(np.abs(bev_np).sum(axis=1) > 0)            # (b, x, y) columns with any value
        out["bev_cols"] = nz.astype(np.int32)
        out["bev_vals"] = bev_np[nz[:, 0], :, nz[:, 1], nz[:, 2]]
        out["grad_in"] = grad_quick.numpy().reshape(B * N, D + C, fH, fW)
    else:
        # digest-only: still keep a slice of the BEV and of the input gradient
        out["grad_in_sha"] = np.array(sha(grad_quick.numpy()))
        out["grad_in_sample"] = grad_quick.numpy().reshape(-1)[::997].copy()
        out["bev_sample"] = bev_np.reshape(-1)[::4099].copy()
    return out


CASES = [
    # (cfg, aug, seed, store-full?)
    ("tiny", "train", 0, True),
    ("tiny", "full", 1, True),
    ("tiny_c32", "eval", 0, True),
    ("cfg1", "train", 0, True),
    ("cfg1", "eval", 1, False),
    ("cfg1", "full", 2, False),
    ("cfg2", "train", 0, False),
    ("cfg2", "full", 3, False),
    ("cfg4", "train", 0, False),
]


def main():
    assert R.reference_available(), "needs the reference checkout"
    torch.manual_seed(0)
    models, tools = R.import_reference()
    # gen_dx_bx / arange / linspace corner cases (tools.py:174-179, models.py:161-164)
    extra = {}
    for i, (xb, yb, zb) in enumerate([([-50., 50., .5], [-50., 50., .5], [-10., 10., 20.]),
                                      ([-30., 30., .3], [-15., 15., .15], [-5., 3., 2.5]),
                                      ([-51.2, 51.2, .8], [-51.2, 51.2, .8], [-10., 10., 2.5])]):
        dx, bx, nx = tools.gen_dx_bx(xb, yb, zb)
        extra[f"gdb{i}_in"] = np.array([xb, yb, zb])
        extra[f"gdb{i}_dx"], extra[f"gdb{i}_bx"], extra[f"gdb{i}_nx"] = dx.numpy(), bx.numpy(), nx.numpy()
    for i, (fd, db) in enumerate([((128, 352), [4., 45., 1.]), ((256, 704), [1., 60., .5]),
                                  ((64, 176), [2., 18., .7]), ((900, 1600), [2., 58., .5]),
                                  ((224, 480), [1., 10.05, .35])]):
        class _M:  # minimal carrier for the unbound reference method
            data_aug_conf = {"final_dim": fd}
            grid_conf = {"dbound": db}
            downsample = 16
        fr = models.LiftSplatShoot.create_frustum(_M())
        extra[f"fr{i}_final_dim"], extra[f"fr{i}_dbound"] = np.array(fd), np.array(db)
        fr = fr.detach().numpy()
        # the frustum is separable: store its three axis vectors (and check that it really is)
        assert (fr[..., 0] == fr[0, 0, :, 0][None, None, :]).all()
        assert (fr[..., 1] == fr[0, :, 0, 1][None, :, None]).all()
        assert (fr[..., 2] == fr[:, 0, 0, 2][:, None, None]).all()
        extra[f"fr{i}_xs"], extra[f"fr{i}_ys"], extra[f"fr{i}_ds"] = fr[0, 0, :, 0], fr[0, :, 0, 1], fr[:, 0, 0, 2]
    np.savez_compressed(os.path.join(HERE, "constants.npz"), **extra)
    print("constants.npz")
    for cfg_name, aug, seed, full in CASES:
        out = run_case(models, tools, cfg_name, aug, seed, full)
        fn = os.path.join(HERE, f"{cfg_name}_{aug}_s{seed}.npz")
        np.savez_compressed(fn, **out)
        print(os.path.basename(fn), os.path.getsize(fn) // 1024, "KiB",
              "kept", int(out["n_kept"]), "/", int(out["n_points"]), "voxels", int(out["n_voxels_hit"]))


if __name__ == "__main__":
    main()
