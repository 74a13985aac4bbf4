"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: the lift-splat path shards by batch with no
exchange, so (1) the shards tile the batch, (2) pooling a shard equals the corresponding slab of pooling the
whole batch (checked with the oracle -- the CUDA path is not involved here), (3) timing is the max over ranks
and throughput the sum of the ranks' units over that time."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lss_carla_b200 import dist as D
from lss_carla_b200.synthetic import CONFIGS, make_batch
from oracle import lss_oracle as O


def test_shard_ranges_tile_the_batch():
    for B in (1, 2, 7, 8, 64):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        D.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = CONFIGS["tiny"]                                  # B = 2: one sample per rank
        full = make_batch(cfg, 0, "train")
        mine = D.shard_batch(full, rank, world)
        lo, hi = D.shard_range(cfg.B, rank, world)
        assert mine["trans"].shape[0] == hi - lo and mine["depthnet_out"].shape[0] == (hi - lo) * cfg.N
        dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
        fr = O.create_frustum(cfg.final_dim, list(cfg.dbound))

        def pool(b):
            calib = {k: b[k].numpy() for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
            M1, M2 = O.calib_matrices_torch(calib["rots"], calib["intrins"], calib["post_rots"])
            bev, _ = O.liftsplat_forward(b["depthnet_out"].numpy(), fr, calib, dx, bx, nx, cfg.C, M1=M1, M2=M2)
            return bev

        part = pool(mine)
        # gather the shards on every rank and compare with pooling the whole batch locally: no exchange is needed
        parts = [torch.zeros_like(torch.from_numpy(part)) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(part))
        whole = pool(full)
        ok_bev = bool(np.array_equal(torch.cat(parts).numpy(), whole))
        # timing: max over ranks; throughput: all units / slowest time
        t = D.max_over_ranks(0.5 + rank)
        rate = D.whole_job_rate(100.0 * (rank + 1), 0.5 + rank)
        q.put((rank, ok_bev, t, rate))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_sharded_pooling_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_bev, t, rate in res:
        assert ok_bev, f"rank {rank}: concatenated shard BEVs differ from the whole-batch BEV"
        assert t == 1.5                                        # slowest rank
        assert abs(rate - 300.0 / 1.5) < 1e-9                  # (100 + 200) units / 1.5 s


def test_single_process_helpers_do_not_need_a_group():
    assert D.max_over_ranks(0.25) == 0.25
    assert D.whole_job_rate(10.0, 0.5) == 20.0
