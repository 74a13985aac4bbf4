"""Profiling aid (not product): per-CTA phase timestamps of k_splat_fwd_tile via %globaltimer."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
NT = int(os.environ.get("NT", "1600"))
tb = torch.zeros(NT * 6 * 8, dtype=torch.int64, device=dev)
os.environ["LSS_FWD_TBUF"] = str(tb.data_ptr())
from lss_carla_b200 import ops
from lss_carla_b200.synthetic import CONFIGS, make_batch
from lss_carla_b200.tools import gen_dx_bx
from lss_carla_b200.api import LiftSplat
cfg = CONFIGS["cfg2"]
dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, *cfg.fHW, cfg.C, dx, bx, nx)
ls = LiftSplat(cfg.grid_conf, cfg.data_aug_conf, device=dev)
b = make_batch(cfg, 0)
M1, M2 = ops.calib_matrices_device(b["rots"].to(dev), b["intrins"].to(dev), b["post_rots"].to(dev))
calib = (ls.frustum, b["post_trans"].to(dev).reshape(-1, 3), M1.reshape(-1, 3, 3), M2.reshape(-1, 3, 3), b["trans"].to(dev).reshape(-1, 3))
plan = ops.build_plan(prob, calib=calib, sorted=True, tile_cols=int(os.environ.get("TC", "0")))
assert plan.layout.n_tiles == NT, plan.layout.n_tiles
junk = torch.empty(64 * 1024 * 1024, device=dev)
for it in range(3):
    junk.fill_(it)            # flush L2 with dirty lines like a previous step would
    pr, ct = ops.lift_prepare(prob, b["depthnet_out"].to(dev))
    tb.zero_(); torch.cuda.synchronize()
    bev = ops.splat_fwd(prob, plan, pr, ct, "sorted", False)
    torch.cuda.synchronize()
t = tb.cpu().numpy()[: NT * 6].reshape(NT, 6).astype(np.float64)
t0 = t[:, 0].min()
start, zf, seg, end, n = (t[:, 0] - t0) / 1e3, (t[:, 1] - t[:, 0]) / 1e3, (t[:, 2] - t[:, 1]) / 1e3, (t[:, 4] - t[:, 3]) / 1e3, t[:, 5]
ne = n > 0
print("kernel span us", (t[:, 4].max() - t0) / 1e3, " ctas", len(t), "non-empty", int(ne.sum()))
print("start time  us: p50 %.2f p90 %.2f max %.2f" % (np.percentile(start, 50), np.percentile(start, 90), start.max()))
print("zero-fill   us: mean %.2f p90 %.2f max %.2f" % (zf[ne].mean(), np.percentile(zf[ne], 90), zf[ne].max()))
print("segments    us: mean %.2f p50 %.2f p90 %.2f max %.2f" % (seg[ne].mean(), np.percentile(seg[ne], 50), np.percentile(seg[ne], 90), seg[ne].max()))
print("store       us: mean %.2f p90 %.2f max %.2f" % (end[ne].mean(), np.percentile(end[ne], 90), end[ne].max()))
for lo, hi in [(1, 64), (64, 160), (160, 320), (320, 640), (640, 4000)]:
    m = (n >= lo) & (n < hi)
    if m.any():
        print(f"  n in [{lo},{hi}): ctas {int(m.sum()):4d}  segments-phase mean {seg[m].mean():.2f} us  max {seg[m].max():.2f}")
order = np.argsort(t[:, 0])
print("cta start times, every 100th (us):", np.round(start[order][::100], 2))
