#!/usr/bin/env python
"""SURVEY.md Appendix B on the B200: the REAL reference (baseline/_ref, stubbed imports) next to liblss_b200.

    python scripts/gpu_reference_probe.py [--out gpurun_out/gpu_reference_probe.json] [--quick]

Answers, per config / seed / augmentation mode:
  1. is the reference's get_geometry ON THE GPU (cuBLAS bmm, models.py:180,187) bit-equal to the library's geometry
     (un-fused fp32, the association measured on the CPU)?  How many voxel indices differ?
  2. is `rots.matmul(inverse(intrins).cuda())` bit-equal between GPU and CPU (models.py:186)?
  3. is the GPU `ranks.argsort()` (models.py:230) the stable order?  Is the reference run-to-run bit-identical?
  4. CUDA-event timings of the reference get_geometry / lift / voxel_pooling, forward and forward+backward.
  5. reference-on-GPU and library BEV against the float64 sum.
Measurement infrastructure only; nothing here is on the product path."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from baseline import refload  # noqa: E402
from lss_carla_b200 import ops  # noqa: E402
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad  # noqa: E402


def ev_time(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def ref_index(model, geom, B):
    """models.py:212-231 replayed with the model's own parameters: idx, kept, ranks, argsort (GPU tensors)."""
    Np = geom.numel() // 3
    g = ((geom - (model.bx - model.dx / 2.)) / model.dx).long().view(Np, 3)
    bix = torch.cat([torch.full([Np // B, 1], ix, device=geom.device, dtype=torch.long) for ix in range(B)])
    g = torch.cat((g, bix), 1)
    kept = (g[:, 0] >= 0) & (g[:, 0] < model.nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < model.nx[1]) \
        & (g[:, 2] >= 0) & (g[:, 2] < model.nx[2])
    gk = g[kept]
    ranks = gk[:, 0] * (model.nx[1] * model.nx[2] * B) + gk[:, 1] * (model.nx[2] * B) + gk[:, 2] * B + gk[:, 3]
    return g, kept, ranks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_reference_probe.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    assert torch.cuda.is_available() and refload.reference_available()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    models, tools = refload.import_reference()
    props = torch.cuda.get_device_properties(dev)
    res = {"device": {"name": props.name, "sms": props.multi_processor_count, "l2_bytes": props.L2_cache_size,
                      "host_cpus": os.cpu_count(), "torch": torch.__version__},
           "allow_tf32_matmul": torch.backends.cuda.matmul.allow_tf32, "cases": [], "timings_ms": {}}

    cfg_names = ["tiny", "cfg1", "cfg2"] + ([] if args.quick else ["cfg4"])
    seeds = range(2 if args.quick else 5)
    for name in cfg_names:
        cfg = CONFIGS[name]
        model = refload.build_liftsplat_model(models, cfg, dev)
        fH, fW = cfg.fHW
        prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, model.dx, model.bx, model.nx)
        for aug in ("train", "eval", "full"):
            for seed in seeds:
                b = make_batch(cfg, seed, aug)
                cal = {k: b[k].to(dev) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
                with torch.no_grad():
                    g_ref = model.get_geometry(cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"])
                    M1 = torch.inverse(cal["post_rots"].cpu()).to(dev)
                    M2_gpu = cal["rots"].matmul(torch.inverse(cal["intrins"].cpu()).to(dev))
                    M2_cpu = cal["rots"].cpu().matmul(torch.inverse(cal["intrins"].cpu()))
                    g_lib = ops.geometry(prob, model.frustum.detach(), cal["post_trans"].reshape(-1, 3), M1.reshape(-1, 3, 3),
                                         M2_gpu.reshape(-1, 3, 3), cal["trans"].reshape(-1, 3))
                    g_lib_cpuM2 = ops.geometry(prob, model.frustum.detach(), cal["post_trans"].reshape(-1, 3), M1.reshape(-1, 3, 3),
                                               M2_cpu.to(dev).reshape(-1, 3, 3), cal["trans"].reshape(-1, 3))
                    neq = int((g_ref.view(torch.int32) != g_lib.view(torch.int32)).sum())
                    idx_r, kept_r, ranks_r = ref_index(model, g_ref, cfg.B)
                    idx_l, kept_l, _ = ref_index(model, g_lib, cfg.B)
                    idx_c, kept_c, _ = ref_index(model, g_lib_cpuM2, cfg.B)
                    # device inverse mode of the library (closed-form adjugate inside the plan build)
                    M1d, M2d = ops.calib_matrices_device(cal["rots"], cal["intrins"], cal["post_rots"])
                    g_dev = ops.geometry(prob, model.frustum.detach(), cal["post_trans"].reshape(-1, 3), M1d.reshape(-1, 3, 3),
                                         M2d.reshape(-1, 3, 3), cal["trans"].reshape(-1, 3))
                    idx_d, kept_d, _ = ref_index(model, g_dev, cfg.B)
                    stable = torch.argsort(ranks_r, stable=True)
                    plain = ranks_r.argsort()
                    case = {"cfg": name, "aug": aug, "seed": int(seed), "n_points": int(g_ref.numel() // 3),
                            "geom_coords_bit_unequal_ref_gpu_vs_lib": neq,
                            "M2_bit_unequal_gpu_vs_cpu": int((M2_gpu.cpu().view(torch.int32) != M2_cpu.view(torch.int32)).sum()),
                            "voxel_idx_mismatch_ref_gpu_vs_lib": int(((idx_r != idx_l).any(1) & (kept_r | kept_l)).sum()),
                            "kept_mismatch_ref_gpu_vs_lib": int((kept_r != kept_l).sum()),
                            "voxel_idx_mismatch_ref_gpu_vs_lib_cpuM2": int(((idx_r != idx_c).any(1) & (kept_r | kept_c)).sum()),
                            "voxel_idx_mismatch_ref_gpu_vs_lib_device_inverse": int(((idx_r != idx_d).any(1) & (kept_r | kept_d)).sum()),
                            "n_kept": int(kept_r.sum()),
                            "argsort_equals_stable": bool(torch.equal(stable, plain))}
                res["cases"].append(case)
                print(json.dumps(case), flush=True)
        # ---- timings + value comparison at seed 0, train augmentation
        b = make_batch(cfg, 0, "train")
        cal = {k: b[k].to(dev) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
        calt = tuple(cal[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans"))
        dn = b["depthnet_out"].to(dev).view(cfg.B, cfg.N, cfg.D + cfg.C, fH, fW)
        gb = make_bev_grad(cfg, 0).to(dev)

        def f_geom():
            with torch.no_grad():
                return model.get_geometry(*calt)

        def f_lift():
            with torch.no_grad():
                return model.get_cam_feats(dn)

        geom = f_geom()
        xfeat = f_lift()

        def f_pool():
            with torch.no_grad():
                return model.voxel_pooling(geom, xfeat)

        def f_fwd():
            with torch.no_grad():
                return model.get_voxels(dn, *calt)

        def f_step():
            x = dn.detach().requires_grad_(True)
            out = model.get_voxels(x, *calt)
            out.backward(gb)
            return out, x.grad

        iters = 5 if name == "cfg4" else 10
        torch.cuda.reset_peak_memory_stats()
        t = {"get_geometry": ev_time(f_geom, iters), "lift(get_cam_feats)": ev_time(f_lift, iters),
             "voxel_pooling_fwd": ev_time(f_pool, iters), "get_voxels_fwd": ev_time(f_fwd, iters),
             "get_voxels_fwd_bwd": ev_time(f_step, iters)}
        t["peak_mem_MB"] = torch.cuda.max_memory_allocated() / 1e6
        t["mpoints_per_s_fwd_bwd"] = cfg.points / t["get_voxels_fwd_bwd"] / 1e3
        # run-to-run identity of the reference on the GPU
        o1, g1 = f_step()
        o2, g2 = f_step()
        t["ref_run_to_run_bit_identical"] = bool(torch.equal(o1, o2) and torch.equal(g1, g2))
        # library (sorted mode, reference inverse) vs reference on the same GPU, and both vs the float64 sum
        M1 = torch.inverse(cal["post_rots"].cpu()).to(dev)
        M2 = cal["rots"].matmul(torch.inverse(cal["intrins"].cpu()).to(dev))
        plan = ops.build_plan(prob, calib=(model.frustum.detach(), cal["post_trans"].reshape(-1, 3), M1.reshape(-1, 3, 3),
                                           M2.reshape(-1, 3, 3), cal["trans"].reshape(-1, 3)), sorted=True)
        x = b["depthnet_out"].to(dev).requires_grad_(True)
        bev = ops.lift_splat(x, prob, plan, "sorted", False)
        bev.backward(gb)
        with torch.no_grad():
            pr = dn[:, :, :cfg.D].reshape(cfg.B * cfg.N, cfg.D, fH, fW).double().softmax(1)
            feat = (pr.unsqueeze(1) * dn[:, :, cfg.D:].reshape(cfg.B * cfg.N, cfg.C, fH, fW).double().unsqueeze(2))
            feat = feat.view(cfg.B, cfg.N, cfg.C, cfg.D, fH, fW).permute(0, 1, 3, 4, 5, 2).reshape(-1, cfg.C)
            vox = plan.vox.long()
            X, Y, Z = (int(v) for v in model.nx)
            truth = torch.zeros(cfg.B * Z * X * Y, cfg.C, dtype=torch.float64, device=dev)
            truth.index_add_(0, vox[vox >= 0], feat[vox >= 0])
            truth = truth.view(cfg.B, Z, X, Y, cfg.C).permute(0, 1, 4, 2, 3).reshape(cfg.B, Z * cfg.C, X, Y)
            t["max_abs_ref_gpu_vs_fp64"] = float((o1.double() - truth).abs().max())
            t["max_abs_lib_vs_fp64"] = float((bev.double() - truth).abs().max())
            t["max_abs_lib_vs_ref_gpu"] = float((bev - o1).abs().max())
            t["max_abs_grad_lib_vs_ref_gpu"] = float((x.grad - g1.view_as(x.grad)).abs().max())
            t["grad_scale"] = float(g1.abs().max())
        res["timings_ms"][name] = t
        print(name, json.dumps(t), flush=True)
        del model
        torch.cuda.empty_cache()

    # host-CPU timing of the reference voxel_pooling (cfg1 / cfg2) with the core count
    torch.set_num_threads(os.cpu_count() or 1)
    cpu = {}
    for name in ("cfg1", "cfg2"):
        cfg = CONFIGS[name]
        model = refload.build_liftsplat_model(models, cfg)
        fH, fW = cfg.fHW
        b = make_batch(cfg, 0, "train")
        from oracle import ref_torch_cpu as T  # CPU geometry (the reference's .cuda() hops need a GPU tensor round trip)
        geom = T.geometry(model.frustum, b["rots"], b["trans"], b["intrins"], b["post_rots"], b["post_trans"])
        with torch.no_grad():
            xfeat = model.get_cam_feats(b["depthnet_out"].view(cfg.B, cfg.N, cfg.D + cfg.C, fH, fW))
            model.voxel_pooling(geom, xfeat)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                model.voxel_pooling(geom, xfeat)
                best = min(best, time.perf_counter() - t0)
        cpu[name] = {"voxel_pooling_fwd_ms": best * 1e3, "mpoints_per_s": cfg.points / best / 1e6, "threads": torch.get_num_threads()}
    res["host_cpu_reference"] = cpu
    print(json.dumps(cpu), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
