#!/usr/bin/env python
"""Benchmark of the lift-splat hot path (BASELINE.json: BEV-pool Mpoints/s, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]
                    [--mode sorted|atomic|red] [--layout nchw|channels_last] [--inverse device|reference]

One "step" = one pass of the path over one synthetic SimBEV-shaped batch:
    calibration matrices -> geometry/voxel ids/plan (sorted buckets) -> lift operands (softmax, ctx^T)
    -> splat forward (BEV written once) -> backward (gradient rows + gather to the depthnet-output gradient).
points/step = B*N*D*fH*fW (every frustum point, kept or not).

Numbers in the JSON line
    value     Mpoints/s, whole job (all ranks), inputs resident in HBM, steps replayed as CUDA graphs,
              4 rotating buffer sets (> 126 MB L2) so no step finds its tensors in L2.
    e2e       same metric through the public host-buffer API (`lss_carla_b200.api.StepPipeline.run`): per step the
              pinned depthnet output + calibration block copied in, plan + lift-splat fwd/bwd as one CUDA graph, the
              input gradient + a BEV probe copied out; `--e2e-mode graphs` / `--no-e2e-graph` select the older paths.
    roofline  dominant HBM-bound kernel (the forward's store kernel) alone: algorithmic bytes per launch / mean launch
              time, against MEASURED_PEAKS.json hbm_gbs; `kernel_in_forward` = the same kernel right behind its gather.
    cpu_baseline  oracle/ref_torch_cpu.py (the reference's ATen op chain) on the host cores, bounded sample.

`--impl reference` times that CPU port as the reference arm (the reference is pure Python/PyTorch and
`/root/reference` does not exist on the GPU box).  Under torchrun only rank 0 runs it.
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
L2_BYTES = 126e6


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
    calib = {k: b[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
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
    steps = max(1, min(args.steps, 20))          # bounded: ~0.5-1 s of CPU work per step at cfg2
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


def run_train(args, cfg, rank, world, local):
    """Second BASELINE metric: LSS training samples/s (train_simbev.py:231-248 step on synthetic batches),
    data-parallel over the ranks (DDP, NCCL gradient all-reduce; no collective inside the lift-splat)."""
    assert torch.cuda.is_available(), "bench.py --metric train needs a GPU"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from lss_carla_b200.dist import max_over_ranks
    from lss_carla_b200.harness import TrainStep, make_train_batch
    per_gpu = args.global_batch // world if args.global_batch else cfg.B
    override = None
    if args.splat == "aten":          # baseline arm: the reference's stock-ATen lift-splat (QuickCumsum) on the same GPU
        from oracle import ref_torch_cpu as T

        def override(model, dn, rots, trans, intrins, post_rots, post_trans):
            calib = {"rots": rots, "trans": trans, "intrins": intrins, "post_rots": post_rots, "post_trans": post_trans}
            return T.liftsplat_forward(dn, model.frustum, calib, model.dx, model.bx, model.nx, dn.shape[1] - model.D)
    step = TrainStep(cfg, dev, splat_mode=args.mode, inverse_mode=args.inverse, splat_override=override, ddp=world > 1,
                     local_rank=local)
    nsets = 2
    batches = [make_train_batch(cfg, per_gpu, 10 * rank + i, dev) for i in range(nsets)]
    steps, warm = min(args.steps, 200), max(3, min(args.warmup, 20))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warm):
        step(batches[i % nsets])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for i in range(steps):
            step(batches[i % nsets])
        e1.record()
        barrier()
    elapsed = max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    value = world * per_gpu * steps / elapsed
    # e2e: images + calibration from pinned host memory every step, loss read back every step
    host = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in batches]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    e2e_steps = max(5, steps // 4)
    probe = torch.empty(1, dtype=torch.float32).pin_memory()
    barrier()
    e0.record()
    for i in range(e2e_steps):
        b = {k: v.to(dev, non_blocking=True) for k, v in host[i % nsets].items()}
        probe.copy_(step(b).detach().reshape(1), non_blocking=True)
    e1.record()
    barrier()
    e2e_elapsed = max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    if rank == 0:
        H, W = cfg.final_dim
        line = {"metric": "lss_train_samples_per_s", "value": round(value, 2), "unit": "samples/s", "n_gpus": world,
                "steps": steps, "warmup": warm, "ms_per_step": round(elapsed / steps * 1e3, 3), "higher_is_better": True,
                "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"LSS training step (fwd + BCE + bwd + clip 5.0 + Adam), {cfg.N} cams {H}x{W}, D={cfg.D}, "
                                       f"grid {cfg.nx[0]}x{cfg.nx[1]}x{cfg.nx[2]}, EfficientNet-B0-shaped trunk + ResNet-18 BEV encoder (PyTorch)",
                           "lift_splat": "liblss_b200 fused path" if args.splat == "ours" else "reference ATen op chain (QuickCumsum) on the GPU",
                           "per_gpu_batch": per_gpu, "global_batch": per_gpu * world, "parallelism": f"dp{world}",
                           "l2": "activations of one step (> 1 GB) exceed the 126 MB L2"},
                "clocks": clk.summary(),
                "e2e": {"value": round(world * per_gpu * e2e_steps / e2e_elapsed, 2), "unit": "samples/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": e2e_steps},
                "gpu_launches": (LAUNCHES_PER_STEP[args.mode] * steps) if args.splat == "ours" else 0,
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

    def __init__(self, cfg, prob, seed, dev, channels_last, tile_cols):
        from lss_carla_b200 import ops
        b = make_batch(cfg, seed, "train")
        self.host = b
        self.rots, self.intrins, self.post_rots = (b[k].to(dev) for k in ("rots", "intrins", "post_rots"))
        self.post_trans, self.trans = b["post_trans"].to(dev).reshape(-1, 3), b["trans"].to(dev).reshape(-1, 3)
        self.dn = b["depthnet_out"].to(dev)
        fmt = torch.channels_last if channels_last else torch.contiguous_format
        self.grad_bev = make_bev_grad(cfg, seed).to(dev).contiguous(memory_format=fmt)
        self.plan = ops.Plan(prob, dev, tile_cols)
        self.rows = torch.empty((max(prob.n_voxels, self.plan.layout.n_rows_cap), prob.C), dtype=torch.float32, device=dev)
        self.vsum = torch.empty((self.plan.layout.n_rows_cap, prob.C), dtype=torch.float32, device=dev)
        self.lift_out = (torch.empty((2, prob.B * prob.N, prob.D, prob.fH, prob.fW), dtype=torch.float32, device=dev),
                         torch.empty((prob.B * prob.N, prob.fH * prob.fW, prob.C), dtype=torch.float32, device=dev))
        self.side = torch.cuda.Stream(device=dev)      # lift operands are independent of the plan: second stream
        self.bev_out = torch.empty(prob.bev_shape, dtype=torch.float32, device=dev, memory_format=fmt)
        self.grad_out = torch.empty((prob.B * prob.N, prob.D + prob.C, prob.fH, prob.fW), dtype=torch.float32, device=dev)
        self.out = {}


STAGES = ("calib", "plan_build", "lift_prepare", "splat_fwd", "splat_bwd")
NO_OVERLAP = bool(os.environ.get("LSS_BENCH_NO_OVERLAP"))
# calibration inverses inside k_voxel_index (lss_plan_build_raw) vs k_calib_matrices + programmatic dependent launch of
# k_voxel_index: with the single-wave voxel-index kernel the fused build is ahead (3467 vs 3426 Mpoints/s at cfg 2; it
# was 2953 vs 2989 with the two-wave kernel) and is the default -- the same entry point the e2e API path uses.
# LSS_BENCH_FUSED_CALIB=0 selects the separate kernel.
NO_FUSED_CALIB = os.environ.get("LSS_BENCH_FUSED_CALIB", "1") == "0"
# sample-range pipelining of gather/store on two streams: measured slower at cfg 2 (2 parts: 2404, 4 parts: 1868 vs 2876
# Mpoints/s unsplit) -- every kernel already fills the GPU, smaller launches only add tails -- so it stays off
PARTS = 1 if NO_OVERLAP else int(os.environ.get("LSS_BENCH_PARTS", "1"))


def one_step(ops, prob, frustum, bs, mode, channels_last, inverse, upto=len(STAGES)):
    """The whole path for one batch; every launch goes through the C ABI on the current stream.
    `upto` < 5 runs only the first stages (used to attribute in-step time to each stage)."""
    cur = torch.cuda.current_stream()
    overlap = upto >= 3 and not NO_OVERLAP
    if overlap:                                        # lift_prepare next to the plan build (fork / join)
        bs.side.wait_stream(cur)
        with torch.cuda.stream(bs.side):
            pr, ct = ops.lift_prepare(prob, bs.dn, out=bs.lift_out)
    if inverse == "device" and not NO_FUSED_CALIB:
        if upto < 2:                                   # (attribution only: the fused build has no calib launch)
            return
        ops.build_plan_raw(prob, frustum, bs.rots, bs.trans, bs.intrins, bs.post_rots, bs.post_trans,
                           sorted=(mode == "sorted"), plan=bs.plan)
    else:
        if inverse == "device":
            M1, M2 = ops.calib_matrices_device(bs.rots, bs.intrins, bs.post_rots)
        else:
            M1, M2 = ops.calib_matrices_reference(bs.rots, bs.intrins, bs.post_rots)
        bs.out.update({"M1": M1, "M2": M2})
        if upto < 2:
            return
        calib = (frustum, bs.post_trans, M1.reshape(-1, 3, 3), M2.reshape(-1, 3, 3), bs.trans)
        ops.build_plan(prob, calib=calib, sorted=(mode == "sorted"), plan=bs.plan)
    if upto < 3:
        return
    if overlap:
        cur.wait_stream(bs.side)
    else:
        pr, ct = ops.lift_prepare(prob, bs.dn, out=bs.lift_out)
    bs.out.update({"pr": pr, "ct": ct})
    if upto < 4:
        return
    if PARTS > 1 and mode == "sorted":      # sample ranges on two streams: stores / row gathers overlap the gathers
        bev = ops.splat_fwd_pipelined(prob, bs.plan, pr, ct, bs.side, PARTS, channels_last, voxel_sums=bs.vsum, out=bs.bev_out)
    else:
        bev = ops.splat_fwd(prob, bs.plan, pr, ct, mode, channels_last, voxel_sums=bs.vsum)
    bs.out["bev"] = bev
    if upto < 5:
        return
    if PARTS > 1 and mode == "sorted":
        bs.out["grad"] = ops.splat_bwd_pipelined(prob, bs.plan, bs.grad_bev, pr, ct, bs.side, PARTS, bs.rows, out=bs.grad_out)
    else:
        bs.out["grad"] = ops.splat_bwd(prob, bs.plan, bs.grad_bev, pr, ct, bs.rows)


# kernels of liblss_b200.so per step: calib, voxel+count, scatter, [sort], lift, forward (sorted: gather + store;
# atomic: one tile kernel; red: memset + one kernel), gradient rows, gather
LAUNCHES_PER_STEP = {"sorted": 9, "atomic": 7, "red": 8}     # minus 1 with the calibration fused into the plan build


def time_kernel(fn, sets, iters, stream):
    """Mean duration of one launch of `fn(bs)`: one CUDA graph per buffer set (no host overhead in the
    timed region), replayed back to back over the rotating sets, bracketed by two events."""
    for bs in sets:
        fn(bs)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graphs, keep = [], []
    for bs in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            keep.append(fn(bs))
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
    ap.add_argument("--layout", default="nchw", choices=["nchw", "channels_last"])
    ap.add_argument("--inverse", default="device", choices=["device", "reference"])
    ap.add_argument("--tile-cols", type=int, default=0)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2000)
    ap.add_argument("--e2e-inverse", default="device", choices=["device", "reference"])
    ap.add_argument("--e2e-streams", type=int, default=4, help="e2e: steps (pinned buffer sets + device buffers) in flight")
    ap.add_argument("--no-e2e-graph", dest="e2e_graph", action="store_false", help="e2e through eager LiftSplat.__call__ instead of StepGraph")
    ap.add_argument("--e2e-mode", default="pipeline", choices=["pipeline", "graphs"],
                    help="pipeline: api.StepPipeline (copy-in / compute / copy-out streams); graphs: api.StepGraph per stream")
    ap.add_argument("--metric", default="pool", choices=["pool", "train"],
                    help="pool: BEV-pool Mpoints/s fwd+bwd (headline); train: LSS training samples/s")
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
    ls = api.LiftSplat(cfg.grid_conf, cfg.data_aug_conf, C=cfg.C, splat_mode=args.mode, inverse_mode=args.e2e_inverse,
                       bev_channels_last=channels_last, device=dev, tile_cols=args.tile_cols)
    frustum = ls.frustum
    sets = [BufferSet(cfg, prob, 100 * rank + i, dev, channels_last, args.tile_cols) for i in range(args.sets)]
    stream = torch.cuda.current_stream()

    # ---- CUDA graphs of one step per buffer set (launch-bound sequence of 8 small kernels)
    use_graph = not args.no_graph and args.inverse == "device"
    graphs = []
    for bs in sets:
        one_step(ops, prob, frustum, bs, args.mode, channels_last, args.inverse)
    torch.cuda.synchronize()
    if use_graph:
        side = torch.cuda.Stream()
        for bs in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                one_step(ops, prob, frustum, bs, args.mode, channels_last, args.inverse)
            graphs.append(g)
        torch.cuda.synchronize()

    def step(i):
        if use_graph:
            graphs[i % len(graphs)].replay()
        else:
            one_step(ops, prob, frustum, sets[i % len(sets)], args.mode, channels_last, args.inverse)

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
    e2e_steps = max(10, min(args.e2e_steps, args.steps))
    pinned = []
    n_e2e = max(len(sets), args.e2e_streams) if (args.e2e_graph and args.e2e_inverse == "device") else len(sets)
    for i in range(n_e2e):        # pinned host batches of the e2e loop (their own seeds beyond the device buffer sets)
        hb = sets[i].host if i < len(sets) else make_batch(cfg, 100 * rank + i, "train")
        h = {k: hb[k].pin_memory() for k in ("depthnet_out", "rots", "trans", "intrins", "post_rots", "post_trans")}
        h["grad_out"] = torch.empty_like(h["depthnet_out"]).pin_memory()
        h["probe"] = torch.empty(1024, dtype=torch.float32).pin_memory()
        pinned.append(h)
    h2d = sum(pinned[0][k].numel() * 4 for k in ("depthnet_out", "rots", "trans", "intrins", "post_rots", "post_trans"))
    d2h = pinned[0]["grad_out"].numel() * 4 + pinned[0]["probe"].numel() * 4

    def e2e_step(i):
        h, bs = pinned[i % len(pinned)], sets[i % len(sets)]
        x = ls.upload(h["depthnet_out"]).requires_grad_(True)
        bev = ls(x, h["rots"], h["trans"], h["intrins"], h["post_rots"], h["post_trans"])
        bev.backward(bs.grad_bev)
        ls.download(x.grad, h["grad_out"])
        ls.download(bev.detach().reshape(-1)[:1024], h["probe"])

    e2e_api = (f"lss_carla_b200.api.LiftSplat.__call__ + autograd backward + LiftSplat.download; inverse_mode={args.e2e_inverse}; "
               "stream-ordered pinned H2D/D2H copies inside the timed region")
    if args.e2e_graph and args.e2e_inverse == "device" and args.e2e_mode == "pipeline":
        # three-stage pipeline (api.StepPipeline): copy-in stream, ONE compute stream replaying the step's kernel graph,
        # copy-out stream; `--e2e-streams` steps (pinned buffer sets + device buffers) in flight
        pstreams = api.PipelineStreams(dev)
        fH, fW = cfg.fHW
        psteps = []
        for i in range(len(pinned)):
            hb = api.pinned_step_buffers(cfg.B, cfg.N, cfg.D + cfg.C, fH, fW)
            for k in ("depthnet_out", "rots", "trans", "intrins", "post_rots", "post_trans"):
                hb[k].copy_(pinned[i][k].reshape(hb[k].shape))
            psteps.append(api.StepPipeline(ls, hb, sets[i % len(sets)].grad_bev, pstreams))
        d2h = psteps[0].host["grad_out"].numel() * 4 + psteps[0].host["probe"].numel() * 4
        e2e_api = ("lss_carla_b200.api.StepPipeline.run (H2D of the step's pinned depthnet output + calibration block on a "
                   "copy-in stream, plan + lift-splat fwd/bwd as one CUDA graph on the shared compute stream, D2H of the input "
                   f"gradient and a BEV probe on a copy-out stream; {len(psteps)} steps in flight)")

        def e2e_step(i):       # noqa: F811
            psteps[i % len(psteps)].run()

        def fork():
            for st in pstreams.all():
                st.wait_stream(torch.cuda.current_stream())

        def join():
            for st in pstreams.all():
                torch.cuda.current_stream().wait_stream(st)
    elif args.e2e_graph and args.e2e_inverse == "device":
        # the same step as one CUDA graph per pinned buffer set, replayed alternately on two streams: the copies of one
        # step overlap the kernels of its neighbour; every replay still moves that step's inputs H2D and results D2H
        streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, min(args.e2e_streams, len(pinned))))]
        sgraphs = [api.StepGraph(ls, pinned[i], sets[i % len(sets)].grad_bev, streams[i % len(streams)]) for i in range(len(pinned))]
        e2e_api = ("lss_carla_b200.api.StepGraph.replay (H2D of the step's pinned inputs + plan + lift-splat fwd/bwd + D2H of the "
                   f"input gradient and a BEV probe, one CUDA graph per buffer set, round-robin on {len(streams)} streams)")

        def e2e_step(i):       # noqa: F811
            sgraphs[i % len(sgraphs)].replay()

        def fork():
            for st in streams:
                st.wait_stream(torch.cuda.current_stream())

        def join():
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
    else:
        def fork():
            pass

        def join():
            pass
    for i in range(args.warmup):
        e2e_step(i)
    barrier()
    e0.record()
    fork()
    for i in range(e2e_steps):
        e2e_step(i)
    join()
    e1.record()
    barrier()
    e2e_elapsed = max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    e2e_value = world * cfg.points * e2e_steps / e2e_elapsed / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel timings (rank 0): each stage alone, back to back over the rotating sets
    kiters = 200
    for bs in sets:
        one_step(ops, prob, frustum, bs, args.mode, channels_last, args.inverse)
    torch.cuda.synchronize()
    layout_code = 1 if channels_last else 0
    # (a) in-step attribution: graphs of the first k stages, T_k - T_(k-1) = cost of stage k inside the step
    #     (its inputs are L2-hot exactly as in the real step); (b) each stage alone (inputs L2-cold)
    instep, prev = {}, 0.0
    # (the reference inverse mode calls LAPACK on the host inside the step: not capturable, attributed with device inverses)
    attr_inverse = "device"
    for k, name in enumerate(STAGES, start=1):
        tk = time_kernel(lambda bs, k=k: one_step(ops, prob, frustum, bs, args.mode, channels_last, attr_inverse, upto=k),
                         sets, kiters, stream)
        instep[name] = tk - prev
        prev = tk
    if args.inverse != "device":
        for bs in sets:                                # M1 / M2 for the stage timings below
            bs.out["M1"], bs.out["M2"] = ops.calib_matrices_device(bs.rots, bs.intrins, bs.post_rots)
    stages = {}
    stages["calib"] = time_kernel(lambda bs: ops.calib_matrices_device(bs.rots, bs.intrins, bs.post_rots), sets, kiters, stream)
    fused_calib = args.inverse == "device" and not NO_FUSED_CALIB
    if fused_calib:
        stages["plan_build"] = time_kernel(
            lambda bs: ops.build_plan_raw(prob, frustum, bs.rots, bs.trans, bs.intrins, bs.post_rots, bs.post_trans,
                                          sorted=(args.mode == "sorted"), plan=bs.plan), sets, kiters, stream)
    else:
        stages["plan_build"] = time_kernel(
            lambda bs: ops.build_plan(prob, calib=(frustum, bs.post_trans, bs.out["M1"].reshape(-1, 3, 3),
                                                   bs.out["M2"].reshape(-1, 3, 3), bs.trans),
                                      sorted=(args.mode == "sorted"), plan=bs.plan), sets, kiters, stream)
    stages["lift_prepare"] = time_kernel(lambda bs: ops.lift_prepare(prob, bs.dn), sets, kiters, stream)
    stages["splat_fwd"] = time_kernel(lambda bs: ops.splat_fwd(prob, bs.plan, bs.out["pr"], bs.out["ct"], args.mode, channels_last, voxel_sums=bs.vsum),
                                      sets, kiters, stream)
    if args.mode == "sorted":      # the two kernels of the deterministic forward, each alone
        stages["k_fwd_gather"] = time_kernel(lambda bs: ops.splat_fwd(prob, bs.plan, bs.out["pr"], bs.out["ct"], args.mode, channels_last,
                                                                      variant="group_gather", voxel_sums=bs.vsum, out=bs.out["bev"]), sets, kiters, stream)
        stages["k_fwd_store"] = time_kernel(lambda bs: ops.splat_fwd(prob, bs.plan, bs.out["pr"], bs.out["ct"], args.mode, channels_last,
                                                                     variant="group_store", voxel_sums=bs.vsum, out=bs.out["bev"]), sets, kiters, stream)
    stages["splat_bwd"] = time_kernel(lambda bs: ops.splat_bwd(prob, bs.plan, bs.grad_bev, bs.out["pr"], bs.out["ct"], bs.rows),
                                      sets, kiters, stream)

    X, Y, Z = cfg.nx
    IN = 4 * cfg.B * cfg.N * (cfg.D + cfg.C) * fH * fW
    G = 4 * cfg.B * cfg.C * Z * X * Y
    v_hit = int((sets[0].out["bev"].reshape(cfg.B, Z, cfg.C, X, Y).abs().sum(2) > 0).sum()) if not channels_last else \
        int((sets[0].out["bev"].abs().sum(1) > 0).sum())
    fwd_bytes = IN + G                                    # forward (fused)
    bwd_bytes = 4 * cfg.C * v_hit + 2 * IN                # backward (fused)
    peak, peak_src = load_peaks()
    step_s = elapsed / args.steps
    if args.mode == "sorted":
        # dominant HBM-bound kernel: k_fwd_store writes every BEV element once (G) and reads the compact voxel rows
        kname = "k_fwd_store_rows" if not channels_last else "k_fwd_store_tma"
        kbytes, ksec = G + 4 * cfg.C * v_hit, stages["k_fwd_store"]
    else:
        kname = {"atomic": "k_splat_fwd_tile (shared-memory atomics)", "red": "memset + k_splat_fwd_red"}[args.mode]
        kbytes, ksec = fwd_bytes, stages["splat_fwd"]
    achieved = kbytes / ksec / 1e9
    traffic = None
    try:      # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel, from the ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"{args.workload}_{args.mode}_{args.layout}")
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
            "algorithmic_bytes_per_launch": kbytes, "kernel_us": round(ksec * 1e6, 2),
            "kernel_share_of_step": round(ksec / step_s, 3),
            # informational: the same kernel right behind its gather (compact rows L2-resident, programmatic launch), by
            # difference of two direct measurements; `achieved` / `frac` above stay the conservative stand-alone figure
            "kernel_in_forward": ({"us": round((stages["splat_fwd"] - stages["k_fwd_gather"]) * 1e6, 2),
                                   "frac": round(kbytes / (stages["splat_fwd"] - stages["k_fwd_gather"]) / 1e9 / peak, 4),
                                   "how": "lss_splat_fwd alone minus k_fwd_gather alone"}
                                  if args.mode == "sorted" and "k_fwd_gather" in stages
                                  and stages["splat_fwd"] > stages["k_fwd_gather"] else None),
            "forward_op": {"what": "lss_splat_fwd (all its launches), IN + G bytes", "bytes": fwd_bytes,
                           "us": round(stages["splat_fwd"] * 1e6, 2), "frac": round(fwd_bytes / stages["splat_fwd"] / 1e9 / peak, 4)},
            "step_algorithmic_bytes": fwd_bytes + bwd_bytes,
            "step_frac": round((fwd_bytes + bwd_bytes) / step_s / 1e9 / peak, 4),
            "stage_us_in_step": {k: round(v * 1e6, 2) for k, v in instep.items()},
            "stage_us_alone_l2_cold": {k: round(v * 1e6, 2) for k, v in stages.items()}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        n = 3
        v, ms, cores = cpu_reference_run(cfg, n, 1)
        cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} full {cfg.name} fwd+bwd steps of oracle/ref_torch_cpu.py (reference ATen op chain), {ms:.0f} ms/step"}

    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(step_s * 1e3, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "splat_mode": args.mode, "bev_layout": args.layout,
                       "inverse": args.inverse + (" (fused into the plan build)" if (args.inverse == "device" and not NO_FUSED_CALIB) else ""),
                       "cuda_graph": use_graph,
                       "l2": f"{args.sets} rotating buffer sets (~{(2 * G + IN) * (1 if channels_last else 1.5) / 1e6:.0f} MB each) > 126 MB L2",
                       "points_per_step_per_gpu": cfg.points, "voxels_hit": v_hit},
            "clocks": clk.summary(),
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": round(e2e_elapsed / e2e_steps * 1e3, 4),
                    "api": e2e_api},
            "gpu_launches": (LAUNCHES_PER_STEP[args.mode] - (1 if (args.inverse == "device" and not NO_FUSED_CALIB) else 0)) * args.steps,
            "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
