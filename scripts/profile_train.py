#!/usr/bin/env python
"""torch.profiler summary of the training step (the second north_star metric): where the GPU time of a step goes -- camera trunk
and BEV encoder (cuDNN / cuBLAS / ATen kernels of PyTorch), the lift-splat (k_* kernels of liblss_b200), the NCCL all-reduce --
and how busy the GPU is (GPU kernel time / wall time of the step: an eager fp32 PyTorch trunk is launch-bound at small batches).

    python scripts/profile_train.py [--per-gpu 8] [--steps 6]                       # one GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/profile_train.py   # DDP
Writes gpurun_out/train_profile_n<N>.json (rank 0); summarised in profiles/r02_train_profile.md."""
import argparse
import collections
import json
import os
import re
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lss_carla_b200.harness import TrainStep, make_train_batch  # noqa: E402
from lss_carla_b200.synthetic import CONFIGS  # noqa: E402

CATS = [("lift-splat (liblss_b200)", r"^(void )?k_"), ("NCCL", r"nccl"), ("convolution / GEMM (cuDNN, cuBLAS)", r"cudnn|conv|gemm|cutlass|xmma|sm\d+_|implicit|wgrad|dgrad|winograd"),
        ("batch norm", r"batch_norm|bn_fw|bn_bw|batchnorm"), ("optimizer / clip (multi-tensor)", r"multi_tensor|adam|lpnorm|norm_kernel"),
        ("pooling / upsample / cat / copy", r"pool|upsample|CatArray|copy|memcpy|Memcpy|memset|Memset|fill"), ("other elementwise / reduce", r".")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=8)
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS["cfg2"]
    step = TrainStep(cfg, dev, ddp=world > 1, local_rank=local)
    batches = [make_train_batch(cfg, args.per_gpu, 10 * rank + i, dev) for i in range(2)]
    for i in range(4):
        step(batches[i % 2])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(batches[i % 2])
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) / args.steps * 1e3
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            step(batches[i % 2])
        torch.cuda.synchronize()
    per_kernel = collections.Counter()
    launches = collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA and ev.device_time_total > 0:
            per_kernel[ev.name] += ev.device_time_total
            launches[ev.name] += 1
    cats = collections.OrderedDict((c, [0.0, 0]) for c, _ in CATS)
    for name, us in per_kernel.items():
        for c, pat in CATS:
            if re.search(pat, name):
                cats[c][0] += us
                cats[c][1] += launches[name]
                break
    gpu_ms = sum(per_kernel.values()) / args.steps / 1e3
    out = {"n_gpus": world, "per_gpu_batch": args.per_gpu, "steps": args.steps, "wall_ms_per_step": round(wall_ms, 2),
           "samples_per_s": round(world * args.per_gpu / wall_ms * 1e3, 1), "gpu_kernel_ms_per_step": round(gpu_ms, 2),
           "gpu_busy_fraction": round(gpu_ms / wall_ms, 3), "kernel_launches_per_step": sum(launches.values()) // args.steps,
           "categories_ms_per_step": {c: {"ms": round(v[0] / args.steps / 1e3, 3), "launches": v[1] // args.steps} for c, v in cats.items()},
           "top_kernels_ms_per_step": [{"name": n[:90], "ms": round(us / args.steps / 1e3, 3), "launches": launches[n] // args.steps}
                                       for n, us in per_kernel.most_common(14)]}
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"train_profile_n{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
