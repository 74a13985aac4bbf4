#!/usr/bin/env python
"""Benchmark of the lift-splat hot path (BASELINE.json: BEV-pool Mpoints/s fwd+bwd & LSS training samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
                    [--layout channels_last|nchw] [--mode sorted|atomic|red] [--no-train] [--no-gpu-reference]

One "step" = one pass of the path over one synthetic SimBEV-shaped batch, cold plan included:
    channels_last (default, run plan):  k_prologue   zero-fill of the BEV || lift operands || run index (geometry, voxel rows)
                                        k_run_classify, k_fwd_gather_cl (every non-empty voxel row written once)
                                        k_bwd_gather_px (gradient rows read straight from the channels_last BEV gradient)
    nchw (the reference's memory format, tile plan): voxel index, scatter, sort, lift, gather, store, gradient rows, gather
points/step = B*N*D*fH*fW (every frustum point, kept or not).  The default layout is the one the consumer of the BEV tensor
runs fastest on (cuDNN conv1 of BevEncode: 7.8 ms vs 11.6 ms forward+backward, profiles/r02_conv1_layout.json).

Numbers in the ONE JSON line (rank 0)
    value        Mpoints/s, whole job (all ranks), inputs resident in HBM, steps replayed as CUDA graphs over 4 rotating
                 buffer sets (> 126 MB L2), max over ranks of the device time.
    e2e          the same metric through the public host-buffer API (`lss_carla_b200.api.StepPipeline.run`): per step the pinned
                 depthnet output + calibration block copied in, the step's kernels as one CUDA graph, the input gradient and a
                 1024-float probe of the BEV copied out.  The BEV itself (81.9 MB) and the upstream gradient stay on the device:
                 their consumer / producer is BevEncode on the same GPU.
    roofline     the FUSED FORWARD OP (IN + G algorithmic bytes) against MEASURED_PEAKS.json hbm_gbs, timed live
                 with CUDA events; `step_frac` for the whole step, per-kernel details and ncu DRAM traffic beside it.
    cpu_baseline oracle/ref_torch_cpu.py (the reference's ATen op chain) on the host cores, bounded sample.
    gpu_reference the UNMODIFIED reference classes (baseline/_ref, staged by scripts/install_reference.py) on this GPU: stock
                 ATen get_geometry + lift + voxel_pooling forward + autograd backward, CUDA-event timed, same inputs.
    train        second BASELINE metric: LSS training samples/s (train_simbev.py:231-248 step, DDP over the ranks, NCCL gradient
                 all-reduce): weak scaling (8 samples per GPU) and strong scaling (global batch 64), with the DDP bucket count
                 and the exposed all-reduce time (step time minus the same step under no_sync()).

`--impl reference` times the CPU port as the reference arm (the reference is pure Python/PyTorch and `/root/reference` does
not exist on the GPU box).  Under torchrun only rank 0 runs it.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad  # noqa: E402

METRIC = "bev_pool_mpoints_per_s_fwd_bwd"
UNIT = "Mpoints/s"
CALIB_KEYS = ("rots", "trans", "intrins", "post_rots", "post_trans")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
# clock / throttle sampling during the timed region (NVML)
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown",
                 "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown",
                 "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self.stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, attr in names.items():
                    if r & getattr(nv, attr, 0):
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's ATen op chain on the host cores
# ------------------------------------------------------------------------------------------------

def cpu_reference_run(cfg, steps, warmup, seed=0):
    from oracle import lss_oracle as O
    from oracle import ref_torch_cpu as T
    torch.set_num_threads(os.cpu_count() or 1)
    b = make_batch(cfg, seed, "train")
    dx, bx, nx = (torch.from_numpy(a) for a in O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound))
    frustum = torch.from_numpy(O.create_frustum(cfg.final_dim, list(cfg.dbound)))
    calib = {k: b[k] for k in CALIB_KEYS}
    gb = make_bev_grad(cfg, seed)
    for _ in range(warmup):
        T.liftsplat_step(b["depthnet_out"], frustum, calib, dx, bx, nx, cfg.C, gb)
    t0 = time.perf_counter()
    for _ in range(steps):
        T.liftsplat_step(b["depthnet_out"], frustum, calib, dx, bx, nx, cfg.C, gb)
    dt = time.perf_counter() - t0
    return cfg.points * steps / dt / 1e6, dt / steps * 1e3, torch.get_num_threads()


def run_reference_arm(args, cfg, rank, world):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))          # bounded: ~0.2 s of CPU work per step at cfg2
    warm = max(1, min(args.warmup, 2))
    val, ms, cores = cpu_reference_run(cfg, steps, warm)
    sample = f"{steps} full {cfg.name} fwd+bwd steps (B={cfg.B}) of the reference ATen op chain, torch CPU, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "where": "host CPU (oracle/ref_torch_cpu.py port of the reference path)"},
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def gpu_reference_run(cfg, dev, iters=3):
    """The UNMODIFIED reference (baseline/_ref) on this GPU: get_voxels (get_geometry + lift + voxel_pooling with QuickCumsum)
    forward + autograd backward with the camera trunk replaced by the identity, CUDA-event timed."""
    try:
        from baseline import refload
        if not refload.reference_available():
            return {"unavailable": "baseline/_ref not staged (scripts/install_reference.py runs in the build container)"}
        models, _ = refload.import_reference()
        model = refload.build_liftsplat_model(models, cfg, dev)
        fH, fW = cfg.fHW
        b = make_batch(cfg, 0, "train")
        cal = tuple(b[k].to(dev) for k in CALIB_KEYS)
        dn = b["depthnet_out"].to(dev).view(cfg.B, cfg.N, cfg.D + cfg.C, fH, fW)
        gb = make_bev_grad(cfg, 0).to(dev)

        def step():
            x = dn.detach().requires_grad_(True)
            model.get_voxels(x, *cal).backward(gb)

        def fwd():
            with torch.no_grad():
                model.get_voxels(dn, *cal)

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        ms, ms_f = timed(step), timed(fwd)
        return {"value": round(cfg.points / ms / 1e3, 3), "unit": UNIT, "ms_per_step": round(ms, 2), "forward_ms": round(ms_f, 2),
                "what": "unmodified reference LiftSplatShoot.get_voxels (stock ATen, QuickCumsum) fwd + autograd bwd on this GPU, "
                        f"{iters} iterations after one warm-up, same synthetic inputs", "source": "baseline/_ref (verbatim copy of src/models.py, src/tools.py)"}
    except Exception as e:      # measurement aid only: never fail the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------------
# training metric (DDP)
# ------------------------------------------------------------------------------------------------

def train_measure(cfg, dev, local, rank, world, per_gpu, steps, warm, mode, inverse, channels_last=True, splat="ours", amp=False):
    """samples/s of the training step (train_simbev.py:231-248: fwd + BCE + bwd + clip 5.0 + Adam) at `per_gpu` samples per
    rank, data-parallel over the ranks (DDP, NCCL all-reduce of the gradients; no collective inside the lift-splat)."""
    import contextlib
    import torch.distributed as dist
    from lss_carla_b200.dist import max_over_ranks
    from lss_carla_b200.harness import TrainStep, make_train_batch
    override = None
    if splat == "aten":          # baseline arm: the reference's stock-ATen lift-splat (QuickCumsum) on the same GPU
        from oracle import ref_torch_cpu as T

        def override(model, dn, rots, trans, intrins, post_rots, post_trans):
            calib = {"rots": rots, "trans": trans, "intrins": intrins, "post_rots": post_rots, "post_trans": post_trans}
            return T.liftsplat_forward(dn, model.frustum, calib, model.dx, model.bx, model.nx, dn.shape[1] - model.D)
    step = TrainStep(cfg, dev, splat_mode=mode, inverse_mode=inverse, splat_override=override, ddp=world > 1, local_rank=local,
                     channels_last=channels_last, amp=amp)
    batches = [make_train_batch(cfg, per_gpu, 10 * rank + i, dev) for i in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, ctx=contextlib.nullcontext):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with ctx():
            for i in range(n):
                step(batches[i % 2])
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)

    for i in range(warm):
        step(batches[i % 2])
    # eager trunk, ~1300 launches per step, host-bound on a busy box: two passes of `steps` steps each (alternating with the
    # no-all-reduce form under DDP), the faster pass of each form is reported
    elapsed, nosync = timed(steps), None
    if world > 1:
        nosync = timed(steps, step.net.no_sync)     # the same step without the gradient all-reduce
    elapsed = min(elapsed, timed(steps))
    if world > 1:
        nosync = min(nosync, timed(steps, step.net.no_sync))
    out = {"samples_per_s": round(world * per_gpu * steps / elapsed, 2), "ms_per_step": round(elapsed / steps * 1e3, 3),
           "per_gpu_batch": per_gpu, "global_batch": per_gpu * world, "steps": steps, "warmup": warm, "passes": 2}
    if amp:
        out["autocast"] = "bfloat16"
    if world > 1:
        out["ms_per_step_no_allreduce"] = round(nosync / steps * 1e3, 3)
        out["exposed_allreduce_ms"] = round((elapsed - nosync) / steps * 1e3, 3)
        n_par = sum(p.numel() * 4 for p in step.model.parameters() if p.requires_grad)
        out["allreduce_bytes_per_step"] = n_par
        try:
            log = step.net._get_ddp_logging_data()
            sizes = str(log.get("rebuilt_bucket_sizes") or log.get("bucket_sizes", ""))
            out["allreduce_buckets"] = len([s for s in sizes.split(",") if s.strip()]) or None
        except Exception:
            out["allreduce_buckets"] = None
        if not out.get("allreduce_buckets"):
            out["allreduce_buckets"] = int(-(-n_par // (25 * 1024 * 1024)))      # DDP default bucket cap 25 MiB
    del step, batches
    torch.cuda.empty_cache()
    return out


def run_train(args, cfg, rank, world, local):
    """`--metric train`: the training metric as the line's headline (the default run carries it in `train`)."""
    assert torch.cuda.is_available(), "bench.py --metric train needs a GPU"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    per_gpu = args.global_batch // world if args.global_batch else cfg.B
    steps, warm = min(args.steps, 200), max(3, min(args.warmup, 20))
    with ClockSampler(local) as clk:
        r = train_measure(cfg, dev, local, rank, world, per_gpu, steps, warm, args.mode, args.inverse,
                          channels_last=args.layout == "channels_last", splat=args.splat)
    if rank == 0:
        H, W = cfg.final_dim
        line = {"metric": "lss_train_samples_per_s", "value": r["samples_per_s"], "unit": "samples/s", "n_gpus": world,
                "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"LSS training step (fwd + BCE + bwd + clip 5.0 + Adam), {cfg.N} cams {H}x{W}, D={cfg.D}, "
                                       f"grid {cfg.nx[0]}x{cfg.nx[1]}x{cfg.nx[2]}, EfficientNet-B0-shaped trunk + ResNet-18 BEV encoder (PyTorch)",
                           "lift_splat": "liblss_b200 fused path" if args.splat == "ours" else "reference ATen op chain (QuickCumsum) on the GPU",
                           "bev_layout": args.layout, "per_gpu_batch": per_gpu, "global_batch": per_gpu * world, "parallelism": f"dp{world}"},
                "clocks": clk.summary(), "train": r, "gpu_launches": (4 * steps) if args.splat == "ours" else 0,
                "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def workload_name(cfg):
    fH, fW = cfg.fHW
    X, Y, Z = cfg.nx
    return (f"{cfg.name}: LSS lift-splat fwd+bwd, bsz={cfg.B}/GPU, {cfg.N} cams {cfg.final_dim[0]}x{cfg.final_dim[1]} "
            f"(feat {fH}x{fW}), D={cfg.D}, C={cfg.C}, grid {X}x{Y}x{Z}")


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

class BufferSet:
    """One complete set of device tensors for a step (inputs, plan workspace, outputs)."""

    def __init__(self, cfg, prob, seed, dev, channels_last, tile_cols, run):
        from lss_carla_b200 import ops
        b = make_batch(cfg, seed, "train")
        self.host = b
        self.rots, self.intrins, self.post_rots = (b[k].to(dev) for k in ("rots", "intrins", "post_rots"))
        self.post_trans, self.trans = b["post_trans"].to(dev).reshape(-1, 3), b["trans"].to(dev).reshape(-1, 3)
        self.dn = b["depthnet_out"].to(dev)
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        self.grad_bev = make_bev_grad(cfg, seed).to(dev).contiguous(memory_format=fmt)
        self.lift_out = (torch.empty((2, prob.B * prob.N, prob.D, prob.fH, prob.fW), dtype=torch.float32, device=dev),
                         torch.empty((prob.B * prob.N, prob.fH * prob.fW, prob.C), dtype=torch.float32, device=dev))
        self.bev_out = torch.empty(prob.bev_shape, dtype=torch.float32, device=dev).contiguous(memory_format=fmt)
        self.grad_out = torch.empty((prob.B * prob.N, prob.D + prob.C, prob.fH, prob.fW), dtype=torch.float32, device=dev)
        if run:
            self.plan = ops.RunPlan(prob, dev)
        else:
            self.plan = ops.Plan(prob, dev, tile_cols)
            self.rows = torch.empty((max(prob.n_voxels, self.plan.layout.n_rows_cap), prob.C), dtype=torch.float32, device=dev)
            self.vsum = torch.empty((self.plan.layout.n_rows_cap, prob.C), dtype=torch.float32, device=dev)
            self.side = torch.cuda.Stream(device=dev)      # lift operands are independent of the plan: second stream
        self.out = {}


class RunPlanPath:
    """channels_last, sorted: k_zero_flags || k_prologue (lift || run index) || k_fwd_columns (classify + gather + shared voxels,
    polling READY and the zero-fill progress) -> k_bwd_gather_px."""
    launches = 4
    kernels = ("k_zero_flags (bulk-copy zero-fill with progress counters), k_prologue<raw> (lift + run-index roles), "
               "k_fwd_columns<8> (classify + gather + shared voxels), k_bwd_gather_px<8>")
    stages = ("forward", "backward")

    def __init__(self, ops, prob, frustum):
        self.ops, self.prob, self.fr = ops, prob, frustum

    def plan_args(self, bs):
        return dict(frustum=self.fr, trans=bs.trans, post_trans=bs.post_trans, rots=bs.rots, intrins=bs.intrins, post_rots=bs.post_rots)

    def forward_op(self, bs):            # the whole forward of a step, plan build included (lss_liftsplat_forward): IN + G bytes
        bs.out["bev"], bs.out["pr"], bs.out["ct"] = self.ops.liftsplat_forward(self.prob, bs.plan, bs.dn, bs.lift_out, bs.bev_out,
                                                                                **self.plan_args(bs))

    def forward_kept_plan(self, bs):     # static calibration: the plan of the previous step is kept
        bs.out["bev"], bs.out["pr"], bs.out["ct"] = self.ops.liftsplat_forward(self.prob, bs.plan, bs.dn, bs.lift_out, bs.bev_out)

    def plan_only(self, bs):
        self.ops.build_runplan(self.prob, plan=bs.plan, **self.plan_args(bs))

    def zero_only(self, bs):
        self.ops.bev_zero(self.prob, bs.bev_out.device, out=bs.bev_out)

    def lift_only(self, bs):
        self.ops.lift_prepare(self.prob, bs.dn, out=bs.lift_out)

    def prologue_without_zero(self, bs):
        self.ops.liftsplat_prologue(self.prob, bs.dn, bs.lift_out, None, bs.plan, **self.plan_args(bs))

    def gather(self, bs):                # classify + gather alone, from the existing plan (the tensor is NOT cleared: timing only)
        self.ops.splat_fwd_cl(self.prob, bs.plan, bs.out["pr"], bs.out["ct"], out=bs.bev_out, precleared=True)

    def gather_with_zero(self, bs):      # the one-launch forward from an existing plan (zero CTAs inside the kernel)
        self.ops.splat_fwd_cl(self.prob, bs.plan, bs.out["pr"], bs.out["ct"], out=bs.bev_out)

    def backward(self, bs):
        bs.out["grad"] = self.ops.splat_bwd_cl(self.prob, bs.plan, bs.grad_bev, bs.out["pr"], bs.out["ct"], out=bs.grad_out)

    def step(self, bs, upto=2):
        self.forward_op(bs)
        if upto >= 2:
            self.backward(bs)

    def step_kept_plan(self, bs):        # static calibration (evaluation): the plan of the previous step is kept
        self.forward_kept_plan(bs)
        self.backward(bs)

    def forward_persistent(self, bs):    # the output tensor is kept between steps: only the rows the previous step wrote are cleared
        bs.out["bev"], bs.out["pr"], bs.out["ct"] = self.ops.liftsplat_forward(self.prob, bs.plan, bs.dn, bs.lift_out, bs.bev_out,
                                                                                 persistent=True, **self.plan_args(bs))

    def step_persistent(self, bs):
        self.forward_persistent(bs)
        self.backward(bs)


class TilePlanPath:
    """Reference memory format (NCHW) or the atomic / red modes: the tile-plan kernels of round 1."""
    stages = ("plan_build", "lift+forward", "backward")

    def __init__(self, ops, prob, frustum, mode, channels_last):
        self.ops, self.prob, self.fr, self.mode, self.cl = ops, prob, frustum, mode, channels_last
        self.launches = {"sorted": 8, "atomic": 6, "red": 7}[mode]
        self.kernels = ("k_voxel_index, k_plan_scatter, k_plan_sort, k_lift_prepare, k_fwd_gather + k_fwd_store_rows, "
                        "k_bwd_rows_compact + k_bwd_gather_px") if mode == "sorted" else "tile-plan kernels"

    def plan_only(self, bs):
        self.ops.build_plan_raw(self.prob, self.fr, bs.rots, bs.trans, bs.intrins, bs.post_rots, bs.post_trans,
                                sorted=(self.mode == "sorted"), plan=bs.plan)

    def lift_only(self, bs):
        bs.out["pr"], bs.out["ct"] = self.ops.lift_prepare(self.prob, bs.dn, out=bs.lift_out)

    def gather(self, bs):
        bs.out["bev"] = self.ops.splat_fwd(self.prob, bs.plan, bs.out["pr"], bs.out["ct"], self.mode, self.cl, voxel_sums=bs.vsum)

    def backward(self, bs):
        bs.out["grad"] = self.ops.splat_bwd(self.prob, bs.plan, bs.grad_bev, bs.out["pr"], bs.out["ct"], bs.rows)

    def forward_op(self, bs):
        self.lift_only(bs)
        self.gather(bs)

    def step(self, bs, upto=3):
        cur = torch.cuda.current_stream()
        bs.side.wait_stream(cur)                       # lift_prepare next to the plan build (fork / join)
        with torch.cuda.stream(bs.side):
            self.lift_only(bs)
        self.plan_only(bs)
        cur.wait_stream(bs.side)
        if upto >= 2:
            self.gather(bs)
        if upto >= 3:
            self.backward(bs)


def time_kernel(fn, sets, iters, stream):
    """Mean duration of one call of `fn(bs)`: one CUDA graph per buffer set (no host overhead in the timed region),
    replayed back to back over the rotating sets, bracketed by two events."""
    for bs in sets:
        fn(bs)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graphs = []
    for bs in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn(bs)
        graphs.append(g)
    torch.cuda.synchronize()
    for g in graphs:
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(iters):
        graphs[i % len(graphs)].replay()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="sorted", choices=["sorted", "atomic", "red"])
    ap.add_argument("--layout", default="channels_last", choices=["nchw", "channels_last"])
    ap.add_argument("--inverse", default="device", choices=["device", "reference"])
    ap.add_argument("--tile-cols", type=int, default=0)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the `train` object (training samples/s, DDP)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2000)
    ap.add_argument("--e2e-streams", type=int, default=4, help="e2e: steps (pinned buffer sets + device buffers) in flight")
    ap.add_argument("--train-steps", type=int, default=12)
    ap.add_argument("--metric", default="pool", choices=["pool", "train"],
                    help="pool: BEV-pool Mpoints/s fwd+bwd (headline, carries `train`); train: LSS training samples/s only")
    ap.add_argument("--splat", default="ours", choices=["ours", "aten"], help="--metric train: lift-splat implementation")
    ap.add_argument("--global-batch", type=int, default=0, help="--metric train: fixed global batch (strong scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = CONFIGS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank, world)
        return
    if args.metric == "train":
        run_train(args, cfg, rank, world, local)
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from lss_carla_b200 import api, ops
    from lss_carla_b200.dist import max_over_ranks
    from lss_carla_b200.tools import gen_dx_bx

    args.warmup = max(args.warmup, 3)
    channels_last = args.layout == "channels_last"
    dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    fH, fW = cfg.fHW
    prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, dx, bx, nx)
    ls = api.LiftSplat(cfg.grid_conf, cfg.data_aug_conf, C=cfg.C, splat_mode=args.mode, inverse_mode="device",
                       bev_channels_last=channels_last, device=dev, tile_cols=args.tile_cols)
    frustum = ls.frustum
    run = channels_last and args.mode == "sorted" and ops.runplan_supported(prob) and args.inverse == "device"
    path = RunPlanPath(ops, prob, frustum) if run else TilePlanPath(ops, prob, frustum, args.mode, channels_last)
    sets = [BufferSet(cfg, prob, 100 * rank + i, dev, channels_last, args.tile_cols, run) for i in range(args.sets)]
    stream = torch.cuda.current_stream()

    # ---- CUDA graphs of one step per buffer set
    use_graph = not args.no_graph
    graphs = []
    for bs in sets:
        path.step(bs)
    torch.cuda.synchronize()
    if use_graph:
        side = torch.cuda.Stream()
        for bs in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                path.step(bs)
            graphs.append(g)
        torch.cuda.synchronize()

    def step(i):
        if use_graph:
            graphs[i % len(graphs)].replay()
        else:
            path.step(sets[i % len(sets)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
    elapsed = max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    value = world * cfg.points * args.steps / elapsed / 1e6

    # ---- e2e through the public API: host buffers in, input gradient out, every step
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(10, min(args.e2e_steps, args.steps))
        pstreams = api.PipelineStreams(dev)
        psteps = []
        for i in range(max(len(sets), args.e2e_streams)):
            hb = api.pinned_step_buffers(cfg.B, cfg.N, cfg.D + cfg.C, fH, fW)
            src = sets[i].host if i < len(sets) else make_batch(cfg, 100 * rank + i, "train")
            for k in ("depthnet_out",) + CALIB_KEYS:
                hb[k].copy_(src[k].reshape(hb[k].shape))
            psteps.append(api.StepPipeline(ls, hb, sets[i % len(sets)].grad_bev, pstreams))
        h2d = psteps[0].host["in_block"].numel() * 4
        d2h = psteps[0].host["out_block"].numel() * 4
        for i in range(args.warmup):
            psteps[i % len(psteps)].run()
        barrier()
        e0.record()
        for st in pstreams.all():
            st.wait_stream(torch.cuda.current_stream())
        for i in range(e2e_steps):
            psteps[i % len(psteps)].run()
        for st in pstreams.all():
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        barrier()
        e2e_elapsed = max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
        e2e = {"value": round(world * cfg.points * e2e_steps / e2e_elapsed / 1e6, 1), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": round(e2e_elapsed / e2e_steps * 1e3, 4),
               "api": ("lss_carla_b200.api.StepPipeline.run: H2D of the step's pinned depthnet output + calibration block on a copy-in "
                       "stream, the step's kernels as one CUDA graph on the shared compute stream, D2H of the input gradient and a "
                       f"1024-float BEV probe on a copy-out stream; {len(psteps)} steps in flight.  The BEV tensor and the upstream "
                       "BEV gradient stay device-resident (their consumer / producer is BevEncode on the same GPU)")}
        del psteps

    # ---- training metric (all ranks take part: DDP)
    train = None
    if not args.no_train:
        try:
            tcfg = CONFIGS["cfg2"]
            train = {"what": "LSS training step (train_simbev.py:231-248: fwd + BCE pos_weight 2.13 + bwd + clip 5.0 + Adam) on synthetic "
                             "batches, EfficientNet-B0-shaped trunk + ResNet-18 BEV encoder in PyTorch (fp32, channels_last BEV), lift-splat = "
                             "liblss_b200; DDP over the ranks, NCCL gradient all-reduce, no collective inside the lift-splat",
                     "weak": train_measure(tcfg, dev, local, rank, world, 8, args.train_steps, 4, "sorted", "device"),
                     "strong": train_measure(tcfg, dev, local, rank, world, max(1, 64 // world), max(4, args.train_steps // 2), 2,
                                             "sorted", "device"),
                     # the same step under bfloat16 autocast (AMP); the lift-splat reads the bfloat16 depthnet output
                     "weak_amp_bf16": train_measure(tcfg, dev, local, rank, world, 8, args.train_steps, 4, "sorted", "device", amp=True)}
        except Exception as e:
            train = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-stage timings (rank 0): in-step attribution and each piece alone, over the rotating (L2-cold) sets
    kiters = 200
    for bs in sets:
        path.step(bs)
    torch.cuda.synchronize()
    X, Y, Z = cfg.nx
    v_hit = int((sets[0].out["bev"].reshape(cfg.B, Z, cfg.C, X, Y).abs().sum(2) > 0).sum())
    instep, prev = {}, 0.0
    for k, name in enumerate(path.stages, start=1):
        tk = time_kernel(lambda bs, k=k: path.step(bs, upto=k), sets, kiters, stream)
        instep[name] = tk - prev
        prev = tk
    alone = {"plan_build": time_kernel(path.plan_only, sets, kiters, stream),
             "lift_prepare": time_kernel(path.lift_only, sets, kiters, stream),
             "forward_op": time_kernel(path.forward_op, sets, kiters, stream),
             "gather": time_kernel(path.gather, sets, kiters, stream),
             "backward": time_kernel(path.backward, sets, kiters, stream)}
    if run:
        alone["zero_fill"] = time_kernel(path.zero_only, sets, kiters, stream)
        alone["prologue(lift+plan, no zero-fill)"] = time_kernel(path.prologue_without_zero, sets, kiters, stream)
        alone["forward_kept_plan"] = time_kernel(path.forward_kept_plan, sets, kiters, stream)
        alone["one_launch_forward_from_plan(zero+classify+gather)"] = time_kernel(path.gather_with_zero, sets, kiters, stream)
        alone["step_kept_plan"] = time_kernel(path.step_kept_plan, sets, kiters, stream)
        persist_err = None
        try:      # opt-in mode, reported next to the headline (which zero-fills the whole tensor every step)
            alone["forward_persistent_bev"] = time_kernel(path.forward_persistent, sets, kiters, stream)
            alone["step_persistent_bev"] = time_kernel(path.step_persistent, sets, kiters, stream)
        except Exception as e:
            alone.pop("forward_persistent_bev", None)
            persist_err = f"{type(e).__name__}: {e}"[:200]

    IN = 4 * cfg.B * cfg.N * (cfg.D + cfg.C) * fH * fW
    G = 4 * cfg.B * cfg.C * Z * X * Y
    fwd_bytes = IN + G                                    # forward (fused)
    bwd_bytes = 4 * cfg.C * v_hit + 2 * IN                # backward (fused)
    peak, peak_src = load_peaks()
    step_s = elapsed / args.steps
    traffic, traffic_detail = None, None
    try:      # dram__bytes_read.sum + dram__bytes_write.sum per kernel of this build, from the ncu --set full capture (scripts/gpu_profile.sh)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(f"{args.workload}_{args.mode}_{args.layout}")
        if isinstance(t, dict):
            traffic, traffic_detail = t.get("forward_op_total"), t
        else:
            traffic = t
    except Exception:
        pass
    fwd_s = alone["forward_op"]
    roof = {"bound": "hbm", "kernel": "forward of a step (plan build + lift + zero-fill + classify + gather): " + path.kernels.split(", k_bwd")[0],
            "achieved": round(fwd_bytes / fwd_s / 1e9, 1), "peak": peak, "unit": "GB/s", "frac": round(fwd_bytes / fwd_s / 1e9 / peak, 4),
            "traffic": traffic, "traffic_per_kernel": traffic_detail, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
            "algorithmic_bytes_per_launch": fwd_bytes, "algorithmic_bytes": "IN + G (SURVEY.md 8d, forward fused)",
            "kernel_us": round(fwd_s * 1e6, 2), "kernel_share_of_step": round(fwd_s / step_s, 3),
            "step_algorithmic_bytes": fwd_bytes + bwd_bytes, "step_us": round(step_s * 1e6, 2),
            "step_frac": round((fwd_bytes + bwd_bytes) / step_s / 1e9 / peak, 4),
            "backward_op": {"bytes": bwd_bytes, "us": round(alone["backward"] * 1e6, 2),
                            "frac": round(bwd_bytes / alone["backward"] / 1e9 / peak, 4)},
            "stage_us_in_step": {k: round(v * 1e6, 2) for k, v in instep.items()},
            "stage_us_alone_l2_cold": {k: round(v * 1e6, 2) for k, v in alone.items()}}
    if run:     # plans persist while the calibration repeats (api.LiftSplat(plan_cache=n), keyed by the host calibration bytes)
        roof["plan_reuse"] = {"step_us_kept_plan": round(alone["step_kept_plan"] * 1e6, 2), "step_us_cold_plan": round(step_s * 1e6, 2),
                              "hit_rate_static_calibration": 1.0, "hit_rate_training_stream": 0.0,
                              "note": "evaluation calibration is static (src/data_simbev.py:135-143): every batch after the first hits; the training "
                                      "loader draws crop_w per SAMPLE (data_simbev.py:119-133, 128 values), so a batch of 8 repeats with "
                                      "probability 128^-8 -- a per-batch plan never hits there, and since the plan build runs next to the lift "
                                      "and the zero-fill it costs only the difference above: the headline `value` is the cold-plan step"}
    if run and "step_persistent_bev" in alone:
        ps = alone["step_persistent_bev"]
        roof["persistent_bev"] = {"step_us": round(ps * 1e6, 2), "forward_us": round(alone["forward_persistent_bev"] * 1e6, 2),
                                  "mpoints_per_s": round(cfg.points / ps / 1e6, 1), "bytes_cleared": 4 * cfg.C * v_hit,
                                  "note": "NOT the headline: ops.liftsplat_forward(persistent=True) / lss_liftsplat_forward_persistent keep the "
                                          "output tensor between steps and zero only the rows the previous step wrote (named by the plan that "
                                          "is about to be overwritten) instead of the whole tensor -- same bits; for callers that own the "
                                          "BEV buffer.  `value` above zero-fills all G bytes every step, as torch.zeros in models.py:240 does"}
    elif run:
        roof["persistent_bev"] = {"unavailable": persist_err}
    if run:     # the one bandwidth-bound piece: the zero-fill role (G bytes) timed alone
        roof["zero_fill"] = {"bytes": G, "us": round(alone["zero_fill"] * 1e6, 2), "frac": round(G / alone["zero_fill"] / 1e9 / peak, 4)}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        n = 3
        v, ms, cores = cpu_reference_run(cfg, n, 1)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} full {cfg.name} fwd+bwd steps of oracle/ref_torch_cpu.py (reference ATen op chain), {ms:.0f} ms/step"}
    gref = None
    if not args.no_gpu_reference and world == 1:
        gref = gpu_reference_run(cfg, dev)

    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(step_s * 1e3, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "splat_mode": args.mode, "bev_layout": args.layout,
                       "plan": "run plan, rebuilt every step (cold)" if run else "tile plan, rebuilt every step (cold)",
                       "inverse": "device (closed-form 3x3 inverses inside the index kernel)", "cuda_graph": use_graph,
                       "kernels_per_step": path.kernels,
                       "l2": f"{args.sets} rotating buffer sets (~{(2 * G + 3 * IN) / 1e6:.0f} MB each) > 126 MB L2",
                       "points_per_step_per_gpu": cfg.points, "voxels_hit": v_hit},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": path.launches * args.steps,
            "roofline": roof, "cpu_baseline": cpu, "gpu_reference": gref, "train": train}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
