"""GPU parity tests of the run-plan path (channels_last BEV, default of bench.py): ops.build_runplan, splat_fwd_cl, splat_bwd_cl.

Everything goes through the C ABI.  The forward must reproduce, BIT FOR BIT and over the WHOLE tensor (cfg 2 / cfg 4 / the
DDP shard sizes included), the sequential definition of the deterministic mode -- per voxel, ascending flat (b,n,d,h,w) index,
float32 adds (oracle.splat_from_prob) -- and therefore the tile-plan path; voxel rows must equal the oracle's indices exactly;
gradients are compared with the float64 analytic gradient at the north_star tolerance (rtol 1e-4 / atol 1e-5).
Nothing here reads /root/reference."""
import dataclasses

import numpy as np
import pytest
import torch

from conftest import load_golden
from lss_carla_b200 import models, ops
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad, make_depthnet_out
from oracle import lss_oracle as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-5
ALL_CASES = ["tiny_train_s0", "tiny_full_s1", "tiny_c32_eval_s0", "cfg1_train_s0", "cfg1_eval_s1",
             "cfg1_full_s2", "cfg2_train_s0", "cfg2_full_s3", "cfg4_train_s0"]


def dev():
    return torch.device("cuda:0")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def problem_of(cfg, g, B=None):
    fH, fW = cfg.fHW
    return ops.Problem.from_grid(B or cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, torch.from_numpy(g["dx"]),
                                 torch.from_numpy(g["bx"]), torch.from_numpy(g["nx"]))


def rows_of_vox(vox, prob):
    """oracle dense id ((b*Z+iz)*X+ix)*Y+iy  ->  run-plan row ((b*X+ix)*Y+iy)*Z+iz, laid out [B,N,fW,D,fH]."""
    X, Y, Z = (int(v) for v in prob.nx)
    v = vox.astype(np.int64)
    iy = v % Y
    ix = (v // Y) % X
    iz = (v // (Y * X)) % Z
    b = v // (Y * X * Z)
    row = np.where(v >= 0, ((b * X + ix) * Y + iy) * Z + iz, -1)
    return row.reshape(prob.B, prob.N, prob.D, prob.fH, prob.fW).transpose(0, 1, 4, 2, 3).astype(np.int32)


def runplan_from_golden(prob, g, plan=None):
    return ops.build_runplan(prob, cu(g["frustum"]), cu(g["trans"]).reshape(-1, 3), cu(g["post_trans"]).reshape(-1, 3),
                             M1=cu(g["M1"]).reshape(-1, 3, 3), M2=cu(g["M2"]).reshape(-1, 3, 3), plan=plan)


def oracle_vox(g, B):
    geom = O.geometry(g["frustum"], g["post_trans"], g["M1"], g["M2"], g["trans"])
    idx, kept = O.voxel_index(geom, g["dx"], g["bx"], g["nx"])
    return O.voxel_linear_id(idx, kept, B, g["nx"])


@pytest.mark.parametrize("case", ALL_CASES)
def test_runplan_whole_tensor_parity(case):
    g = load_golden(case)
    cfg = CONFIGS[str(g["cfg"])]
    seed = int(g["seed"])
    prob = problem_of(cfg, g)
    assert ops.runplan_supported(prob)
    rp = runplan_from_golden(prob, g)
    vox = oracle_vox(g, cfg.B)
    # ---- integer results: voxel row of every point, exact
    assert np.array_equal(rp.prow.cpu().numpy(), rows_of_vox(vox, prob))
    # every kept point sits on exactly one list (the voxel's), and the lists are as long as the oracle's voxel populations
    n_kept = int((vox >= 0).sum())
    rows, pts, n_shared = rp.lists()
    assert int(pts.sum()) == n_kept
    want_rows = rows_of_vox(vox, prob).reshape(-1)
    want_cnt = np.bincount(want_rows[want_rows >= 0], minlength=prob.n_voxels)
    assert np.array_equal(np.bincount(rows, weights=pts, minlength=prob.n_voxels).astype(np.int64), want_cnt)
    # ---- forward: whole tensor, bit-exact against the sequential definition and against the tile-plan path
    dn = (torch.from_numpy(g["depthnet_out"]) if "depthnet_out" in g else make_depthnet_out(cfg, seed)).to(dev())
    pr, ct = ops.lift_prepare(prob, dn)
    bev = ops.splat_fwd_cl(prob, rp, pr, ct)
    assert bev.is_contiguous(memory_format=torch.channels_last) and tuple(bev.shape) == prob.bev_shape
    want = O.splat_from_prob(pr.cpu().numpy(), ct.cpu().numpy(), vox, cfg.B, cfg.C, g["nx"])
    assert np.array_equal(bev.cpu().numpy(), want)
    tp = ops.build_plan(prob, calib=(cu(g["frustum"]), cu(g["post_trans"]).reshape(-1, 3), cu(g["M1"]).reshape(-1, 3, 3),
                                      cu(g["M2"]).reshape(-1, 3, 3), cu(g["trans"]).reshape(-1, 3)), sorted=True)
    assert torch.equal(bev.contiguous(), ops.splat_fwd(prob, tp, pr, ct, "sorted", False))
    # pre-zeroed output issued separately (what the model path does on a side stream)
    assert not rp.scratch.any()                                   # the forward leaves its scratch clean
    assert int(rp.counters[6]) == n_shared                        # voxels that went through the list walk
    z = ops.bev_zero(prob, dev())
    assert not z.any()
    assert torch.equal(ops.splat_fwd_cl(prob, rp, pr, ct, out=z, precleared=True), bev)
    z.fill_(float("nan"))                                         # the two-launch forward on the kept plan, then with a rebuild
    assert torch.equal(ops.liftsplat_forward(prob, rp, dn, out=z)[0], bev)
    z.fill_(float("nan"))
    out, pr2, ct2 = ops.liftsplat_forward(prob, rp, dn, None, z, cu(g["frustum"]), cu(g["trans"]).reshape(-1, 3),
                                          cu(g["post_trans"]).reshape(-1, 3), M1=cu(g["M1"]).reshape(-1, 3, 3), M2=cu(g["M2"]).reshape(-1, 3, 3))
    assert torch.equal(out, bev) and torch.equal(pr2, pr) and torch.equal(ct2, ct) and not rp.scratch.any()
    # ---- backward: whole tensor vs the float64 analytic gradient; same bits as the tile-plan kernels
    gb = make_bev_grad(cfg, seed).to(dev())
    gr = ops.splat_bwd_cl(prob, rp, gb.contiguous(memory_format=torch.channels_last), pr, ct)
    want_g = O.liftsplat_backward(gb.cpu().numpy(), dn.cpu().numpy(), pr.cpu().numpy(), vox, cfg.B, cfg.N, cfg.D, cfg.C, g["nx"])
    np.testing.assert_allclose(gr.cpu().numpy(), want_g, rtol=RTOL, atol=ATOL)
    assert torch.equal(gr, ops.splat_bwd_cl(prob, rp, gb, pr, ct))                   # NCHW gradient: transposed once
    assert torch.equal(gr, ops.splat_bwd(prob, tp, gb, pr, ct))
    if "grad_in" in g:
        np.testing.assert_allclose(gr.cpu().numpy(), g["grad_in"], rtol=RTOL, atol=ATOL)   # the reference's autograd
    # ---- rebuilding into the same workspace gives the same plan and the same bits
    p0, epoch = rp.prow.clone(), int(rp.counters[0])
    assert epoch == 2                                             # built twice so far
    runplan_from_golden(prob, g, plan=rp)
    assert torch.equal(p0, rp.prow) and int(rp.counters[0]) == epoch + 1 and not rp.scratch.any()
    rows2, pts2, n_shared2 = rp.lists()
    assert int(pts2.sum()) == n_kept and n_shared2 == n_shared
    assert torch.equal(ops.splat_fwd_cl(prob, rp, pr, ct), bev)
    assert torch.equal(ops.splat_fwd_cl(prob, rp, pr, ct), bev)                      # the forward only reads the plan


@pytest.mark.parametrize("name,aug", [("tiny", "train"), ("tiny", "full"), ("cfg1", "eval"), ("cfg2", "full")])
def test_runplan_raw_build_equals_matrix_build(name, aug):
    """Closed-form inverses inside k_run_index == lss_calib_matrices + the matrix build, bit for bit."""
    cfg = CONFIGS[name]
    b = make_batch(cfg, 7, aug)
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    g = {"dx": dx, "bx": bx, "nx": nx}
    prob = problem_of(cfg, g)
    fr = cu(O.create_frustum(cfg.final_dim, list(cfg.dbound)))
    cal = {k: b[k].to(dev()) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
    M1, M2 = ops.calib_matrices_device(cal["rots"], cal["intrins"], cal["post_rots"])
    a = ops.build_runplan(prob, fr, cal["trans"].reshape(-1, 3), cal["post_trans"].reshape(-1, 3), M1=M1, M2=M2)
    r = ops.build_runplan(prob, fr, cal["trans"].reshape(-1, 3), cal["post_trans"].reshape(-1, 3), rots=cal["rots"],
                          intrins=cal["intrins"], post_rots=cal["post_rots"])
    assert torch.equal(a.prow, r.prow)
    (ra, pa, sa), (rr, pr_, sr) = a.lists(), r.lists()
    assert sa == sr and np.array_equal(np.bincount(ra, weights=pa), np.bincount(rr, weights=pr_))


def test_runplan_long_voxels_and_all_dropped():
    """A 2 x 2 grid: every voxel holds thousands of points (CTA path, shared-memory and in-place global sort);
    then a grid nobody hits: the BEV is all zero and the gradient is zero."""
    cfg = dataclasses.replace(CONFIGS["tiny"], name="coarse", xbound=(-40.0, 40.0, 40.0), ybound=(-40.0, 40.0, 40.0),
                              zbound=(-10.0, 10.0, 20.0))
    for case_cfg, expect_hits in ((cfg, True), (dataclasses.replace(cfg, xbound=(500.0, 580.0, 40.0)), False)):
        b = make_batch(case_cfg, 1, "train")
        dx, bx, nx = O.gen_dx_bx(case_cfg.xbound, case_cfg.ybound, case_cfg.zbound)
        g = {"dx": dx, "bx": bx, "nx": nx}
        prob = problem_of(case_cfg, g)
        fr = O.create_frustum(case_cfg.final_dim, list(case_cfg.dbound))
        M1, M2 = O.calib_matrices_torch(b["rots"].numpy(), b["intrins"].numpy(), b["post_rots"].numpy())
        rp = ops.build_runplan(prob, cu(fr), b["trans"].to(dev()).reshape(-1, 3), b["post_trans"].to(dev()).reshape(-1, 3),
                               M1=cu(M1).reshape(-1, 3, 3), M2=cu(M2).reshape(-1, 3, 3))
        geom = O.geometry(fr, b["post_trans"].numpy(), M1, M2, b["trans"].numpy())
        idx, kept = O.voxel_index(geom, dx, bx, nx)
        vox = O.voxel_linear_id(idx, kept, case_cfg.B, nx)
        assert np.array_equal(rp.prow.cpu().numpy(), rows_of_vox(vox, prob))
        pr, ct = ops.lift_prepare(prob, b["depthnet_out"].to(dev()))
        bev = ops.splat_fwd_cl(prob, rp, pr, ct)
        want = O.splat_from_prob(pr.cpu().numpy(), ct.cpu().numpy(), vox, case_cfg.B, case_cfg.C, nx)
        assert np.array_equal(bev.cpu().numpy(), want)
        cnt = rp.counters.cpu().numpy()
        assert not rp.scratch.any()
        if expect_hits:
            assert cnt[7] >= 1 and np.bincount(vox[vox >= 0]).max() > 1024      # long voxels beyond the shared-memory sort
            assert torch.equal(ops.splat_fwd_cl(prob, rp, pr, ct), bev)
        else:
            assert not kept.any() and not bev.any() and cnt[6] == 0 and cnt[7] == 0
        gb = make_bev_grad(case_cfg, 1).to(dev())
        gr = ops.splat_bwd_cl(prob, rp, gb, pr, ct)
        want_g = O.liftsplat_backward(gb.cpu().numpy(), b["depthnet_out"].numpy(), pr.cpu().numpy(), vox, case_cfg.B,
                                      case_cfg.N, case_cfg.D, case_cfg.C, nx)
        np.testing.assert_allclose(gr.cpu().numpy(), want_g, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("B", [16, 32, 64])
def test_runplan_ddp_shard_sizes(B):
    """cfg 5 per-GPU shards (global batch 64 over 4 / 2 / 1 GPUs): whole-tensor parity at B = 16 / 32 / 64."""
    cfg = dataclasses.replace(CONFIGS["cfg2"], name=f"cfg5_B{B}", B=B)
    b = make_batch(cfg, 11, "train")
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    prob = problem_of(cfg, {"dx": dx, "bx": bx, "nx": nx})
    fr = O.create_frustum(cfg.final_dim, list(cfg.dbound))
    M1, M2 = O.calib_matrices_torch(b["rots"].numpy(), b["intrins"].numpy(), b["post_rots"].numpy())
    rp = ops.build_runplan(prob, cu(fr), b["trans"].to(dev()).reshape(-1, 3), b["post_trans"].to(dev()).reshape(-1, 3),
                           M1=cu(M1).reshape(-1, 3, 3), M2=cu(M2).reshape(-1, 3, 3))
    geom = O.geometry(fr, b["post_trans"].numpy(), M1, M2, b["trans"].numpy())
    idx, kept = O.voxel_index(geom, dx, bx, nx)
    vox = O.voxel_linear_id(idx, kept, B, nx)
    assert np.array_equal(rp.prow.cpu().numpy(), rows_of_vox(vox, prob))
    pr, ct = ops.lift_prepare(prob, b["depthnet_out"].to(dev()))
    bev = ops.splat_fwd_cl(prob, rp, pr, ct)
    want = O.splat_from_prob(pr.cpu().numpy(), ct.cpu().numpy(), vox, B, cfg.C, nx)
    assert np.array_equal(bev.cpu().numpy(), want)
    # the per-voxel sum does not depend on the batch size: sample 0 equals the B = 1 result
    cfg1 = dataclasses.replace(cfg, B=1)
    p1 = problem_of(cfg1, {"dx": dx, "bx": bx, "nx": nx})
    rp1 = ops.build_runplan(p1, cu(fr), b["trans"][:1].to(dev()).reshape(-1, 3), b["post_trans"][:1].to(dev()).reshape(-1, 3),
                            M1=cu(M1[:1]).reshape(-1, 3, 3), M2=cu(M2[:1]).reshape(-1, 3, 3))
    pr1, ct1 = ops.lift_prepare(p1, b["depthnet_out"][:cfg.N].to(dev()))
    assert torch.equal(ops.splat_fwd_cl(p1, rp1, pr1, ct1)[0], bev[0])


def test_runplan_autograd_and_model_path():
    """ops.lift_splat with a RunPlan (autograd) and the model-level path with bev_channels_last=True (side-stream zero-fill)."""
    g = load_golden("cfg1_train_s0")
    cfg = CONFIGS["cfg1"]
    prob = problem_of(cfg, g)
    rp = runplan_from_golden(prob, g)
    x = cu(g["depthnet_out"]).requires_grad_(True)
    bev = ops.lift_splat(x, prob, rp, "sorted", True)
    assert rp.busy
    gb = make_bev_grad(cfg, 0).to(dev())
    bev.backward(gb)
    assert not rp.busy
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["grad_in"], rtol=RTOL, atol=ATOL)
    # a forward that never reaches backward releases its plan when the graph dies
    y = ops.lift_splat(x, prob, rp, "sorted", True)
    assert rp.busy
    del y
    assert not rp.busy
    # a plan rebuilt between forward and backward is detected
    y = ops.lift_splat(x, prob, rp, "sorted", True)
    runplan_from_golden(prob, g, plan=rp)
    with pytest.raises(RuntimeError, match="rebuilt"):
        y.backward(gb)
    # model level
    m = models.LiftSplatShoot(cfg.grid_conf, cfg.data_aug_conf, outC=1, camencode=torch.nn.Identity(), bevencode=torch.nn.Identity(),
                              splat_mode="sorted", inverse_mode="reference", bev_channels_last=True).to(dev())
    cal = [cu(g[k]) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")]
    x2 = cu(g["depthnet_out"]).requires_grad_(True)
    out = models.lift_splat_from_depthnet(m, x2, *cal)
    assert out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out.detach(), bev.detach())
    out.backward(gb)
    assert torch.equal(x2.grad, x.grad)


@pytest.mark.parametrize("cfg_name", ["tiny", "cfg1"])
def test_runplan_graph_replay_with_changing_calibration(cfg_name):
    """One captured step (lift || plan build -> zero-fill || classify + gather, launched programmatically -> backward)
    replayed with a different calibration every time: no stale plan data may leak from one replay into the next."""
    cfg = CONFIGS[cfg_name]
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    prob = problem_of(cfg, {"dx": dx, "bx": bx, "nx": nx})
    fr = cu(O.create_frustum(cfg.final_dim, list(cfg.dbound)))
    keys = ("rots", "trans", "intrins", "post_rots", "post_trans")
    batches = [make_batch(cfg, s, aug) for s, aug in ((0, "train"), (1, "full"), (2, "eval"), (3, "full"))]
    cal = {k: batches[0][k].to(dev()).clone() for k in keys}
    dn = batches[0]["depthnet_out"].to(dev()).clone()
    gb = make_bev_grad(cfg, 0).to(dev()).contiguous(memory_format=torch.channels_last)
    rp = ops.RunPlan(prob, dev())
    bev = torch.empty(prob.bev_shape, device=dev()).contiguous(memory_format=torch.channels_last)
    grad = torch.empty_like(dn)
    s_main = torch.cuda.Stream()
    lift_out = (torch.empty((2, prob.B * prob.N, prob.D, prob.fH, prob.fW), device=dev()),
                torch.empty((prob.B * prob.N, prob.fH * prob.fW, prob.C), device=dev()))

    def step():      # the step of bench.py / api.StepPipeline: zero-fill || lift || index -> classify + gather -> backward
        _, pr, ct = ops.liftsplat_forward(prob, rp, dn, lift_out, bev, fr, cal["trans"].reshape(-1, 3), cal["post_trans"].reshape(-1, 3),
                                          rots=cal["rots"], intrins=cal["intrins"], post_rots=cal["post_rots"])
        ops.splat_bwd_cl(prob, rp, gb, pr, ct, out=grad)

    with torch.cuda.stream(s_main):
        for _ in range(2):
            step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s_main):
        step()
    for it in range(8):
        b = batches[it % len(batches)]
        for k in keys:
            cal[k].copy_(b[k])
        dn.copy_(b["depthnet_out"])
        bev.fill_(float("nan"))
        graph.replay()
        torch.cuda.synchronize()
        ref_rp = ops.build_runplan(prob, fr, cal["trans"].reshape(-1, 3), cal["post_trans"].reshape(-1, 3), rots=cal["rots"],
                                   intrins=cal["intrins"], post_rots=cal["post_rots"])
        pr, ct = ops.lift_prepare(prob, dn)
        assert torch.equal(bev, ops.splat_fwd_cl(prob, ref_rp, pr, ct)), it
        assert torch.equal(grad, ops.splat_bwd_cl(prob, ref_rp, gb, pr, ct)), it


@pytest.mark.parametrize("one_launch", [False, True])
@pytest.mark.parametrize("cfg_name,B", [("cfg2", 8), ("cfg2", 3), ("cfg1", 1)])
def test_fused_forward_zero_fill_ordering_stress(cfg_name, B, one_launch):
    """The zero-fill runs NEXT TO the classify + gather CTAs (per-sample progress counters) -- in the prologue grid of
    liftsplat_forward, whose successor starts on the READY flag, or as the first CTAs of splat_fwd_cl: 40 back-to-back replays over a tensor poisoned with NaN before every step must all give the reference bits --
    a zero landing after a voxel row, or a row written before its zeros, shows up as a mismatch."""
    cfg = dataclasses.replace(CONFIGS[cfg_name], B=B)
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    prob = problem_of(cfg, {"dx": dx, "bx": bx, "nx": nx})
    fr = cu(O.create_frustum(cfg.final_dim, list(cfg.dbound)))
    b = make_batch(cfg, 5, "train")
    cal = {k: b[k].to(dev()) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
    dn = b["depthnet_out"].to(dev())
    args = dict(trans=cal["trans"].reshape(-1, 3), post_trans=cal["post_trans"].reshape(-1, 3), rots=cal["rots"],
                intrins=cal["intrins"], post_rots=cal["post_rots"])
    ref_rp = ops.build_runplan(prob, fr, **args)
    pr0, ct0 = ops.lift_prepare(prob, dn)
    want = ops.splat_fwd_cl(prob, ref_rp, pr0, ct0, out=ops.bev_zero(prob, dev()), precleared=True)
    rp = ops.RunPlan(prob, dev())
    bev = torch.empty(prob.bev_shape, device=dev()).contiguous(memory_format=torch.channels_last)
    lift_out = (torch.empty((2, prob.B * prob.N, prob.D, prob.fH, prob.fW), device=dev()),
                torch.empty((prob.B * prob.N, prob.fH * prob.fW, prob.C), device=dev()))
    bad = torch.zeros((), dtype=torch.int64, device=dev())
    s_main = torch.cuda.Stream()

    def step():
        bev.fill_(float("nan"))
        if one_launch:
            pr, ct = ops.liftsplat_prologue(prob, dn, lift_out, None, rp, fr, **args)
            ops.splat_fwd_cl(prob, rp, pr, ct, out=bev)
        else:
            ops.liftsplat_forward(prob, rp, dn, lift_out, bev, fr, **args)
        bad.add_((bev != want).sum())                     # NaN != x counts too

    with torch.cuda.stream(s_main):
        step()
    torch.cuda.synchronize()
    assert int(bad) == 0
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s_main):
        step()
    for _ in range(40):
        graph.replay()
    torch.cuda.synchronize()
    assert int(bad) == 0 and not rp.scratch.any()


def test_api_plan_cache_on_repeating_host_calibration():
    """api.LiftSplat(plan_cache=n): a batch whose HOST calibration repeats reuses its plan (the forward only reads it); same bits
    and same gradient as a rebuild, LRU eviction, hit statistics."""
    from lss_carla_b200 import api
    cfg = CONFIGS["cfg1"]
    ls = api.LiftSplat(cfg.grid_conf, cfg.data_aug_conf, C=cfg.C, inverse_mode="device", device=dev(), plan_cache=2)
    plain = api.LiftSplat(cfg.grid_conf, cfg.data_aug_conf, C=cfg.C, inverse_mode="device", device=dev())
    batches = [make_batch(cfg, s, "train") for s in (0, 1, 2)]
    gb = make_bev_grad(cfg, 0).to(dev())
    for i in (0, 0, 1, 0, 2, 1, 0):
        b = batches[i]
        cal = [b[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans")]          # host tensors: hashable without a round trip
        x1 = b["depthnet_out"].to(dev()).requires_grad_(True)
        x2 = b["depthnet_out"].to(dev()).requires_grad_(True)
        o1, o2 = ls(x1, *cal), plain(x2, *cal)
        assert torch.equal(o1, o2)
        o1.backward(gb)
        o2.backward(gb)
        assert torch.equal(x1.grad, x2.grad)
    st = ls.plan_cache_stats()
    assert (st["hits"], st["misses"], st["plans_kept"]) == (2, 5, 2) and abs(st["hit_rate"] - 2 / 7) < 1e-12
    assert plain.plan_cache_stats()["hit_rate"] is None
    dev_cal = [batches[0][k].to(dev()) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")]
    ls(batches[0]["depthnet_out"].to(dev()), *dev_cal)                                            # device calibration: not hashed, not cached
    assert ls.plan_cache_stats()["misses"] == 5


def test_forward_on_an_unbuilt_plan_is_all_zero_and_does_not_hang():
    """The C entry point cannot see whether the workspace holds a plan: a forward without calibration on a fresh (all-zero) workspace
    must come back -- READY is raised for epoch 0 too -- with an all-zero BEV (empty lists), not spin until the 2 s trap."""
    cfg = CONFIGS["tiny"]
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    prob = problem_of(cfg, {"dx": dx, "bx": bx, "nx": nx})
    rp = ops.RunPlan(prob, dev())
    dn = make_batch(cfg, 0, "train")["depthnet_out"].to(dev())
    bev = torch.full(prob.bev_shape, float("nan"), device=dev()).contiguous(memory_format=torch.channels_last)
    ops.liftsplat_forward(prob, rp, dn, out=bev, _allow_unbuilt=True)
    torch.cuda.synchronize()
    assert not bev.any() and not rp.scratch.any()


@pytest.mark.parametrize("cfg_name,B", [("cfg2", 8), ("cfg1", 1), ("tiny", 2), ("cfg4", 1)])
def test_persistent_output_forward_clears_only_the_previous_rows(cfg_name, B):
    """lss_liftsplat_forward_persistent: the output tensor is kept between calls and only the rows the previous call wrote are
    zeroed (by the plan build, from the plan it is about to overwrite).  Over a sequence of different calibrations and depthnet
    outputs -- rebuilds, kept-plan calls, a tensor nobody has paired with the plan yet, an in-place modification by somebody else,
    a plan rebuilt behind its back -- every result must be the regular forward's, bit for bit over the whole tensor."""
    cfg = dataclasses.replace(CONFIGS[cfg_name], B=B)
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    prob = problem_of(cfg, {"dx": dx, "bx": bx, "nx": nx})
    fr = cu(O.create_frustum(cfg.final_dim, list(cfg.dbound)))
    rp, ref_rp = ops.RunPlan(prob, dev()), ops.RunPlan(prob, dev())
    bev = torch.full(prob.bev_shape, float("nan"), device=dev()).contiguous(memory_format=torch.channels_last)
    calls = []

    def run(seed, mode, rebuild=True):
        b = make_batch(cfg, seed, mode)
        args = dict(trans=b["trans"].to(dev()).reshape(-1, 3), post_trans=b["post_trans"].to(dev()).reshape(-1, 3),
                    rots=b["rots"].to(dev()), intrins=b["intrins"].to(dev()), post_rots=b["post_rots"].to(dev()))
        dn = b["depthnet_out"].to(dev())
        paired = rp._bev_ref is not None and rp._bev_ref.data_ptr() == bev.data_ptr() and rp._bev_version == bev._version
        calls.append(paired)
        if rebuild:
            got, _, _ = ops.liftsplat_forward(prob, rp, dn, None, bev, fr, persistent=True, **args)
            want, _, _ = ops.liftsplat_forward(prob, ref_rp, dn, None, None, fr, **args)
        else:               # plan kept: only the depthnet output changes
            got, _, _ = ops.liftsplat_forward(prob, rp, dn, None, bev, persistent=True)
            want, _, _ = ops.liftsplat_forward(prob, ref_rp, dn, None, None)
        assert got.data_ptr() == bev.data_ptr()
        assert torch.equal(got, want), (seed, mode, rebuild)
        assert not rp.scratch.any()

    run(0, "train")                      # not paired yet: the regular path, zero-fill of the NaN tensor included
    run(1, "full")                       # paired: rows of call 0 cleared by the build of call 1
    run(2, "eval")
    run(3, "train", rebuild=False)       # kept plan: nothing to clear
    run(4, "full")
    assert calls == [False, True, True, True, True]
    bev.add_(1.0)                        # somebody else wrote to the tensor: detected, one regular call
    run(5, "train")
    run(6, "full")
    b9 = make_batch(cfg, 9, "train")       # the plan rebuilt by somebody else: detected, one regular call
    ops.build_runplan(prob, fr, b9["trans"].to(dev()).reshape(-1, 3), b9["post_trans"].to(dev()).reshape(-1, 3), rots=b9["rots"].to(dev()),
                      intrins=b9["intrins"].to(dev()), post_rots=b9["post_rots"].to(dev()), plan=rp)
    run(7, "eval")
    run(8, "train")
    assert calls[5:] == [False, True, False, True]
