import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lss_carla_b200 import api
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad
dev = torch.device("cuda:0")
cfg = CONFIGS[os.environ.get("CFG", "cfg2")]
ls = api.LiftSplat(cfg.grid_conf, cfg.data_aug_conf, C=cfg.C, inverse_mode="device", device=dev)
def host_set(seed):
    b = make_batch(cfg, seed, "train")
    h = {k: b[k].pin_memory() for k in ("depthnet_out", "rots", "trans", "intrins", "post_rots", "post_trans")}
    h["grad_out"] = torch.empty_like(h["depthnet_out"]).pin_memory()
    h["probe"] = torch.empty(1024).pin_memory()
    return h
gb = make_bev_grad(cfg, 0).to(dev)
n = int(os.environ.get("NG", "1"))
streams = [torch.cuda.Stream() for _ in range(2)]
hs = [host_set(i) for i in range(n)]
gbs = [gb.clone() if os.environ.get("CLONE_GB") else gb for _ in range(n)]
gs = [api.StepGraph(ls, hs[i], gbs[i], streams[0 if os.environ.get("ONE_STREAM") else i % 2]) for i in range(n)]
print("captured", n); torch.cuda.synchronize()
for it in range(int(os.environ.get("IT", "4"))):
    for g in gs:
        g.replay()
    if os.environ.get("SYNC_EACH"): torch.cuda.synchronize()
torch.cuda.synchronize()
# compare with eager
x = hs[0]["depthnet_out"].to(dev).requires_grad_(True)
bev = ls(x, *[hs[0][k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans")])
bev.backward(gb); torch.cuda.synchronize()
print("grad equal:", torch.equal(x.grad.cpu(), hs[0]["grad_out"]), " probe equal:", torch.equal(bev.detach().reshape(-1)[:1024].cpu(), hs[0]["probe"]))
