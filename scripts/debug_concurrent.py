"""Debug aid: two CUDA graphs of the path replayed concurrently on two streams; STAGE limits how far the step goes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lss_carla_b200 import api, ops, models
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad
dev = torch.device("cuda:0")
cfg = CONFIGS["cfg2"]
STAGE = int(os.environ.get("STAGE", "4"))
ls = api.LiftSplat(cfg.grid_conf, cfg.data_aug_conf, C=cfg.C, inverse_mode="device", device=dev)
gb = make_bev_grad(cfg, 0).to(dev)
fH, fW = cfg.fHW
prob = models._problem_for(ls, cfg.B, cfg.N, fH, fW, cfg.C)
class G:
    def __init__(self, seed, stream):
        b = make_batch(cfg, seed, "train")
        self.t = {k: v.to(dev) for k, v in b.items()}
        self.ws = ops.Plan(prob, dev)
        self.vs = torch.empty((self.ws.layout.n_rows_cap, cfg.C), device=dev)
        self.rows = torch.empty((prob.n_voxels, cfg.C), device=dev)
        self.stream = stream
        self.keep = None
        def step():
            t = self.t
            plan = ops.build_plan_raw(prob, ls.frustum, t["rots"], t["trans"], t["intrins"], t["post_rots"], t["post_trans"], plan=self.ws)
            if STAGE < 2: return None
            pr, ct = ops.lift_prepare(prob, t["depthnet_out"])
            if STAGE < 3: return (pr, ct)
            bev = ops.splat_fwd(prob, plan, pr, ct, "sorted", voxel_sums=self.vs)
            if STAGE < 4: return (pr, ct, bev)
            g = ops.splat_bwd(prob, plan, gb, pr, ct, self.rows)
            return (pr, ct, bev, g)
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            for _ in range(2): step()
        stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=stream):
            self.keep = step()
        stream.synchronize()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
gs = [G(i, streams[i]) for i in range(2)]
for it in range(20):
    for g in gs:
        with torch.cuda.stream(g.stream): g.graph.replay()
torch.cuda.synchronize()
print("STAGE", STAGE, "ok")
