import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lss_carla_b200 import ops
from lss_carla_b200.synthetic import CONFIGS, make_batch
from lss_carla_b200.tools import gen_dx_bx
for name in ("cfg2", "cfg4"):
    cfg = CONFIGS[name]; dev = torch.device("cuda:0")
    dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    fH, fW = cfg.fHW
    prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, dx, bx, nx)
    ds = torch.arange(*cfg.dbound, dtype=torch.float)
    fr = torch.empty(ds.shape[0], fH, fW, 3)
    fr[..., 0] = torch.linspace(0, cfg.final_dim[1] - 1, fW).view(1, 1, fW)
    fr[..., 1] = torch.linspace(0, cfg.final_dim[0] - 1, fH).view(1, fH, 1)
    fr[..., 2] = ds.view(-1, 1, 1)
    fr = fr.to(dev)
    for i in range(2):
        b = make_batch(cfg, i, "train")
        cal = {k: b[k].to(dev) for k in ("rots", "trans", "intrins", "post_rots", "post_trans")}
        rp = ops.build_runplan(prob, fr, trans=cal["trans"].reshape(-1, 3), post_trans=cal["post_trans"].reshape(-1, 3), rots=cal["rots"], intrins=cal["intrins"], post_rots=cal["post_rots"])
        torch.cuda.synchronize()
        q = rp._view(rp.layout.off_qcount, 32 * 16, torch.int64).view(32, 16)[:, 0].cpu()
        cnt = (q & 0xFFFFFFFF).tolist()
        print(name, i, "sum", sum(cnt), "max", max(cnt), "min", min(cnt), "slots", max(cnt) * 32, cnt)
