"""Profiling aid (not product): per-CTA phase timestamps of the GROUP forward kernels via %globaltimer."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
from lss_carla_b200 import ops, _lib
from lss_carla_b200.synthetic import CONFIGS, make_batch
from lss_carla_b200.tools import gen_dx_bx
from lss_carla_b200.api import LiftSplat
cfg = CONFIGS[os.environ.get("CFG", "cfg2")]
dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
prob = ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, *cfg.fHW, cfg.C, dx, bx, nx)
ls = LiftSplat(cfg.grid_conf, cfg.data_aug_conf, device=dev)
b = make_batch(cfg, 0)
M1, M2 = ops.calib_matrices_device(b["rots"].to(dev), b["intrins"].to(dev), b["post_rots"].to(dev))
calib = (ls.frustum, b["post_trans"].to(dev).reshape(-1, 3), M1.reshape(-1, 3, 3), M2.reshape(-1, 3, 3), b["trans"].to(dev).reshape(-1, 3))
plan = ops.build_plan(prob, calib=calib, sorted=True, tile_cols=int(os.environ.get("TC", "0")))
NT, NK = plan.layout.n_tiles, cfg.B * cfg.N * cfg.fHW[1]
tb = torch.zeros(NT * 8, dtype=torch.int64, device=dev)
NM = 2 * 148
tg = torch.zeros((NK + NM) * 8, dtype=torch.int64, device=dev)
_lib.check(_lib.lib().lss_debug_set_timeline(C.c_void_p(tb.data_ptr()), C.c_void_p(tg.data_ptr())))
vs = torch.empty((plan.layout.n_rows_cap, prob.C), device=dev)
junk = torch.empty(64 * 1024 * 1024, device=dev)
cl = os.environ.get("CL", "0") == "1"
flush = os.environ.get("FLUSH", "1") == "1"
for it in range(3):
    if flush:
        junk.fill_(it)            # flush L2 with dirty lines like a previous step would
    pr, ct = ops.lift_prepare(prob, b["depthnet_out"].to(dev))
    tb.zero_(); tg.zero_(); torch.cuda.synchronize()
    bev = ops.splat_fwd(prob, plan, pr, ct, "sorted", cl, voxel_sums=vs)
    torch.cuda.synchronize()
us = lambda a: a / 1e3
def stat(name, a):
    print("%-26s mean %6.2f  p50 %6.2f  p90 %6.2f  max %6.2f" % (name, a.mean(), np.percentile(a, 50), np.percentile(a, 90), a.max()))
gall = tg.cpu().numpy().reshape(NK + NM, 8).astype(np.float64)
g, gm = gall[:NK], gall[NK:]
print('mixed-queue CTAs: records', int(plan._view(plan.layout.off_counters, 2, torch.int32)[1]), ' start us', us(gm[:, 0].min() - g[:, 0].min()), ' end mean/max us', us(gm[:, 3].mean() - g[:, 0].min()), us(gm[:, 3].max() - g[:, 0].min()))
t = tb.cpu().numpy().reshape(NT, 8).astype(np.float64)
g0 = g[:, 0].min()
ne = g[:, 7] > 0
print("== gather: span us %.2f  ctas %d  non-empty %d  records/cta mean %.1f max %d" % (us(g[ne, 3].max() - g0), NK, int(ne.sum()), g[ne, 7].mean(), g[:, 7].max()))
stat("start time", us(g[:, 0] - g0))
stat("key_count load", us(g[ne, 1] - g[ne, 0]))
stat("group 0 first voxel done", us(g[ne, 2] - g[ne, 1]))
stat("all warps done", us(g[ne, 3] - g[ne, 1]))
stat("CTA lifetime", us(g[ne, 3] - g[ne, 0]))
life = us(g[:, 3] - g[:, 0])
o = np.argsort(-life)
fW = cfg.fHW[1]
print("slowest gather CTAs (life us, start, n_rec, b, n, w, smid):", [(round(life[i], 1), round(us(g[i, 0] - g0), 1), int(g[i, 7]), i // (cfg.N * fW), (i // fW) % cfg.N, i % fW, int(g[i, 4])) for i in o[:12]])
print("fastest:", [(round(life[i], 1), int(g[i, 7]), i // (cfg.N * fW), (i // fW) % cfg.N, i % fW, int(g[i, 4])) for i in o[-8:]])
sm = g[:, 4].astype(int)
cnt = np.bincount(sm, minlength=148)
print("CTAs per SM min/max", cnt.min(), cnt.max(), " mean life by SM-load:", {int(c): round(float(life[np.isin(sm, np.where(cnt == c)[0])].mean()), 2) for c in np.unique(cnt)})
print("gap gather end -> store start us %.2f" % us(t[:, 0].min() - g[ne, 3].max()))
t0 = t[:, 0].min()
nseg = t[:, 7]; ne = nseg > 0
print("== store: span us %.2f  ctas %d  non-empty %d" % (us(t[:, 3].max() - t0), NT, int(ne.sum())))
stat("start time (all)", us(t[:, 0] - t0))
stat("meta loads", us(t[ne, 1] - t[ne, 0]))
stat("rows + map", us(t[ne, 2] - t[ne, 1]))
stat("store", us(t[ne, 3] - t[ne, 2]))
stat("CTA lifetime", us(t[ne, 3] - t[ne, 0]))
order = np.argsort(t[:, 0])
print("cta start times, every 100th (us):", np.round(us(t[order, 0] - t0)[::100], 2))
print("cta end times,   every 100th (us):", np.round(us(t[order, 3] - t0)[::100], 2))
