"""Digest of one lift-splat forward + backward (cfg1, sorted mode, NCHW) -- run under a tuning knob (environment variable)
by tests/test_cuda_parity.py::test_tuning_knobs_keep_the_bits: every kernel variant must reproduce the default's bits."""
import hashlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lss_carla_b200 import api
from lss_carla_b200.synthetic import CONFIGS, make_batch, make_bev_grad

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg1"]
dev = torch.device("cuda:0")
ls = api.LiftSplat(cfg.grid_conf, cfg.data_aug_conf, C=cfg.C, inverse_mode="device", device=dev)
b = make_batch(cfg, 7, "train")
x = b["depthnet_out"].to(dev).requires_grad_(True)
bev = ls(x, *[b[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans")])
bev.backward(make_bev_grad(cfg, 7).to(dev))
torch.cuda.synchronize()
h = lambda t: hashlib.sha256(t.detach().cpu().numpy().tobytes()).hexdigest()[:16]
print("DIGEST", h(bev), h(x.grad))
