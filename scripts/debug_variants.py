"""Debug aid (not product): compare every forward variant with the sequential oracle on a golden case."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from lss_carla_b200 import ops
from lss_carla_b200.synthetic import CONFIGS
from oracle import lss_oracle as O
from test_cuda_parity import problem_of, calib_of
dev = torch.device("cuda:0")
for case in sys.argv[1:] or ["tiny_train_s0"]:
    g = load_golden(case); cfg = CONFIGS[str(g["cfg"])]
    prob = problem_of(cfg, g)
    for tc in (0, 24):
        plan = ops.build_plan(prob, calib=calib_of(g), sorted=True, tile_cols=tc)
        pr, ct = ops.lift_prepare(prob, torch.from_numpy(g["depthnet_out"]).to(dev))
        vox = plan.vox.cpu().numpy().astype(np.int64)
        want = O.splat_from_prob(pr.cpu().numpy(), ct.cpu().numpy(), vox, cfg.B, cfg.C, g["nx"])
        for variant in ("warp", "group"):
            for cl in (False, True):
                bev = ops.splat_fwd(prob, plan, pr, ct, "sorted", cl, variant=variant).cpu().numpy()
                bad = bev != want
                print(case, "tc", tc, variant, "cl" if cl else "nchw", "mismatch", int(bad.sum()), "of", bad.size,
                      "maxdiff", float(np.abs(bev - want).max()), "first", np.argwhere(bad)[:3].tolist())
