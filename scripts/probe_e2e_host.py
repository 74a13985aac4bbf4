"""Is the StepGraph e2e loop host-bound?  Times N replays (round-robin over S graphs/streams) on the host clock
without synchronising, then the device completion time.  usage: python scripts/probe_e2e_host.py [streams] [steps]"""
import sys, time, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lss_carla_b200 import api
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad

S = int(sys.argv[1]) if len(sys.argv) > 1 else 6
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
cfg = CONFIGS["cfg2"]
dev = torch.device("cuda:0")
ls = api.LiftSplat(cfg.grid_conf, cfg.data_aug_conf, C=cfg.C, splat_mode="sorted", inverse_mode="device", device=dev)
gb = make_bev_grad(cfg, 0).to(dev)
graphs = []
for i in range(S):
    hb = make_batch(cfg, i, "train")
    h = {k: hb[k].pin_memory() for k in ("depthnet_out", "rots", "trans", "intrins", "post_rots", "post_trans")}
    h["grad_out"] = torch.empty_like(h["depthnet_out"]).pin_memory()
    h["probe"] = torch.empty(1024, dtype=torch.float32).pin_memory()
    graphs.append(api.StepGraph(ls, h, gb, torch.cuda.Stream(device=dev)))
for i in range(50):
    graphs[i % S].replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(N):
    graphs[i % S].replay()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"streams {S}: host issue {1e6 * (t1 - t0) / N:.1f} us/replay, total {1e6 * (t2 - t0) / N:.1f} us/step "
      f"-> {cfg.points / ((t2 - t0) / N) / 1e6:.0f} Mpoints/s")
