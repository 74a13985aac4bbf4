"""Training-step harness for the second BASELINE metric (LSS training samples/s, 1/2/4/8 GPUs).

Mirrors the step structure of the reference's training loop (train_simbev.py:231-248): forward, BCE-with-logits
with pos_weight 2.13 (tools.py:222-229), backward, gradient clipping at 5.0, Adam (lr 1e-3, weight decay 1e-7,
train_simbev.py:29-53, :192), on a synthetic SimBEV-shaped batch resident on the device.  The camera trunk and the
BEV encoder are the PyTorch stand-ins of `trunk.py` (they stay in PyTorch per north_star); the lift-splat between
them is the CUDA path -- or, for the baseline arm, whatever `splat_override` the caller supplies."""
from __future__ import annotations

import types

import torch
from torch import nn

from . import models
from .synthetic import make_calibration


def make_train_batch(cfg, B, seed, device):
    """imgs N(0,1) f32[B,N,3,H,W] (ImageNet-normalised range, tools.py:167-171), calibration of the synthetic rig with
    the script's default augmentation, targets Bernoulli(0.03) f32[B,1,X,Y] (README.md:231)."""
    g = torch.Generator().manual_seed(5000 + seed)
    H, W = cfg.final_dim
    X, Y, _ = cfg.nx
    c = type(cfg)(**{**cfg.__dict__, "B": B}) if B != cfg.B else cfg
    cal = make_calibration(c, seed, "train")
    batch = {k: v.to(device) for k, v in cal.items()}
    batch["imgs"] = torch.randn(B, cfg.N, 3, H, W, generator=g).to(device)
    batch["binimgs"] = (torch.rand(B, 1, X, Y, generator=g) < 0.03).float().to(device)
    return batch


class TrainStep:
    """`channels_last`: BEV emitted in channels_last strides and BevEncode converted to that memory format (the faster
    layout end to end on B200, see models.install); `splat_override(model, depthnet_out, *calibration) -> BEV` replaces the
    lift-splat by something else -- the baseline arm passes the reference's stock ATen op chain here (measurement only: the
    hook lives in this harness, not in the product's model code); `amp`: the forward and the loss under bfloat16 autocast
    (the lift-splat then reads the bfloat16 depthnet output directly, `lss_lift_prepare_bf16`, accumulates in float32 and hands
    BevEncode a float32 BEV; parameters, gradients and Adam stay float32 -- no loss scaling needed with bfloat16)."""

    def __init__(self, cfg, device, splat_mode="sorted", inverse_mode="device", splat_override=None, ddp=False,
                 local_rank=0, seed=0, channels_last=True, amp=False):
        torch.manual_seed(seed)                       # same initial weights on every rank (DDP broadcasts anyway)
        self.model = models.LiftSplatShoot(cfg.grid_conf, cfg.data_aug_conf, outC=1, splat_mode=splat_mode,
                                           inverse_mode=inverse_mode, bev_channels_last=channels_last).to(device)
        if channels_last:
            self.model.bevencode.to(memory_format=torch.channels_last)
        if splat_override is not None:
            def get_voxels(m, x, rots, trans, intrins, post_rots, post_trans):
                B, N, _, imH, imW = x.shape
                ce = m.camencode
                dn = ce.depthnet(ce.dropout(ce.get_eff_depth(x.view(B * N, x.shape[2], imH, imW))))
                return splat_override(m, dn, rots, trans, intrins, post_rots, post_trans)
            self.model.get_voxels = types.MethodType(get_voxels, self.model)
        self.net = self.model
        if ddp:
            self.net = nn.parallel.DistributedDataParallel(self.model, device_ids=[local_rank], gradient_as_bucket_view=True)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=1e-3, weight_decay=1e-7)
        self.loss_fn = nn.BCEWithLogitsLoss(pos_weight=torch.tensor([2.13], device=device))
        self.net.train()
        self.amp = bool(amp)

    def __call__(self, batch):
        self.opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            preds = self.net(batch["imgs"], batch["rots"], batch["trans"], batch["intrins"], batch["post_rots"], batch["post_trans"])
            loss = self.loss_fn(preds.float(), batch["binimgs"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.net.parameters(), 5.0)
        self.opt.step()
        return loss
