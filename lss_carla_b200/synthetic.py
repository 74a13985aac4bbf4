"""Seeded synthetic SimBEV-shaped inputs for the lift-splat path.

Everything is generated on the CPU with a seeded ``torch.Generator`` so the same arrays can be fed
to the CUDA path, to the oracle and (in the build container) to the real reference.

Shapes follow what the reference's loader hands to the model (data_simbev.py:147-218, batch tuple
at data_simbev.py:307) and what ``CamEncode.depthnet`` produces (models.py:47,56):

* ``rots`` f32[B,N,3,3], ``trans`` f32[B,N,3]      camera -> ego
* ``intrins`` f32[B,N,3,3]                          pinhole K for the raw H x W image
* ``post_rots`` f32[B,N,3,3], ``post_trans`` f32[B,N,3]   image augmentation homography
  (tools.py:120-144 ``img_transform``; one augmentation per sample shared by its cameras,
  data_simbev.py:166-168)
* ``depthnet_out`` f32[B*N, D+C, fH, fW]            logits (first D channels) and context (last C)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

# ----------------------------------------------------------------------------------------------
# Workload configurations named by BASELINE.json `configs`
# ----------------------------------------------------------------------------------------------


@dataclass(frozen=True)
class LiftSplatConfig:
    name: str
    B: int
    N: int = 6
    raw_hw: tuple = (224, 480)          # H, W of the raw camera image (train_simbev.py:29-30)
    final_dim: tuple = (128, 352)       # network input H, W (train_simbev.py:32)
    xbound: tuple = (-50.0, 50.0, 0.5)
    ybound: tuple = (-50.0, 50.0, 0.5)
    zbound: tuple = (-10.0, 10.0, 20.0)
    dbound: tuple = (4.0, 45.0, 1.0)
    C: int = 64                         # camC, models.py:148
    downsample: int = 16                # models.py:147

    @property
    def grid_conf(self):
        return {"xbound": list(self.xbound), "ybound": list(self.ybound),
                "zbound": list(self.zbound), "dbound": list(self.dbound)}

    @property
    def data_aug_conf(self):
        return {"final_dim": tuple(self.final_dim), "H": self.raw_hw[0], "W": self.raw_hw[1],
                "resize_lim": (1.0, 1.0), "bot_pct_lim": (0.0, 0.0), "rot_lim": (0.0, 0.0),
                "rand_flip": False, "Ncams": self.N}

    @property
    def fHW(self):
        return self.final_dim[0] // self.downsample, self.final_dim[1] // self.downsample

    @property
    def D(self):
        lo, hi, st = self.dbound
        return int(math.ceil((hi - lo) / st))   # == len(torch.arange(lo, hi, st)) for these bounds

    @property
    def nx(self):
        return tuple(int((b[1] - b[0]) / b[2]) for b in (self.xbound, self.ybound, self.zbound))

    @property
    def points(self):
        fH, fW = self.fHW
        return self.B * self.N * self.D * fH * fW


CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "cfg1": LiftSplatConfig("cfg1", B=1),
    # configs[1]: the configuration the metric is quoted on (bsz 8, 6 cams 128x352, D=41, 200x200x1)
    "cfg2": LiftSplatConfig("cfg2", B=8),
    # configs[3]: multi-height voxels stress test
    "cfg4": LiftSplatConfig("cfg4", B=4, raw_hw=(448, 960), final_dim=(256, 704),
                            zbound=(-10.0, 10.0, 2.5), dbound=(1.0, 60.0, 0.5)),
    # small shapes for fast tests / fixtures
    "tiny": LiftSplatConfig("tiny", B=2, N=3, raw_hw=(112, 240), final_dim=(64, 176),
                            xbound=(-20.0, 20.0, 1.0), ybound=(-20.0, 20.0, 1.0),
                            zbound=(-4.0, 4.0, 2.0), dbound=(2.0, 18.0, 1.0), C=64),
    "tiny_c32": LiftSplatConfig("tiny_c32", B=1, N=2, raw_hw=(112, 240), final_dim=(64, 176),
                                xbound=(-16.0, 16.0, 0.5), ybound=(-12.0, 12.0, 0.5),
                                zbound=(-10.0, 10.0, 20.0), dbound=(2.0, 12.0, 0.5), C=32),
}


# ----------------------------------------------------------------------------------------------
# Calibration rig
# ----------------------------------------------------------------------------------------------

_RING_YAW_DEG = (55.0, 0.0, -55.0, 110.0, 180.0, -110.0)
# camera axes (x right, y down, z forward) expressed in ego axes (x forward, y left, z up)
_CAM_TO_EGO_AXES = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]], dtype=np.float64)


def _rz(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=np.float64)


def make_calibration(cfg: LiftSplatConfig, seed: int = 0, aug: str = "train"):
    """Return dict of CPU float32 tensors: rots, trans, intrins, post_rots, post_trans.

    aug = "train": script defaults (train_simbev.py:31-35): no resize / rotation / flip, random
                   horizontal crop -> post_rots = I, post_trans = (-crop_w, -(H - fH), 0).
    aug = "eval":  centre resize+crop (data_simbev.py:135-142): post_rots = diag(s, s, 1).
    aug = "full":  random resize, rotation and flip through the same algebra as tools.py:120-144,
                   so post_rots is a general 2x2 block (stress case for the 3x3 inverse).
    """
    g = torch.Generator().manual_seed(1000 + seed)
    B, N = cfg.B, cfg.N
    H, W = cfg.raw_hw
    fH_img, fW_img = cfg.final_dim

    def randn(*s):
        return torch.randn(*s, generator=g, dtype=torch.float64).numpy()

    def rand(*s):
        return torch.rand(*s, generator=g, dtype=torch.float64).numpy()

    rots = np.zeros((B, N, 3, 3), np.float64)
    trans = np.zeros((B, N, 3), np.float64)
    intr = np.zeros((B, N, 3, 3), np.float64)
    post_rots = np.zeros((B, N, 3, 3), np.float64)
    post_trans = np.zeros((B, N, 3), np.float64)

    f = W / (2.0 * math.tan(math.radians(90.0) / 2.0))
    K = np.array([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]])

    for b in range(B):
        # ---- one augmentation per sample (data_simbev.py:166-168)
        if aug == "train":
            resize, rotate, flip = 1.0, 0.0, False
            newW, newH = int(W * resize), int(H * resize)
            crop_h = int((1 - 0.0) * newH) - fH_img
            crop_w = int(rand(1)[0] * max(0, newW - fW_img))
        elif aug == "eval":
            resize, rotate, flip = max(fH_img / H, fW_img / W), 0.0, False
            newW, newH = int(W * resize), int(H * resize)
            crop_h = int((1 - 0.0) * newH) - fH_img
            crop_w = int(max(0, newW - fW_img) / 2)
        elif aug == "full":
            resize = float(0.75 + 0.25 * rand(1)[0]) * max(fH_img / H, fW_img / W) * 1.3
            rotate = float(-5.4 + 10.8 * rand(1)[0])
            flip = bool(rand(1)[0] < 0.5)
            newW, newH = int(W * resize), int(H * resize)
            crop_h = int((1 - 0.22 * rand(1)[0]) * newH) - fH_img
            crop_w = int(rand(1)[0] * max(0, newW - fW_img))
        else:
            raise ValueError(aug)
        crop = (crop_w, crop_h, crop_w + fW_img, crop_h + fH_img)
        # post-homography algebra of tools.py:131-142, in float32 like the loader
        pr = torch.eye(2) * resize
        pt = torch.zeros(2) - torch.tensor([float(crop[0]), float(crop[1])])
        if flip:
            A = torch.tensor([[-1.0, 0.0], [0.0, 1.0]])
            bb = torch.tensor([float(crop[2] - crop[0]), 0.0])
            pr = A.matmul(pr)
            pt = A.matmul(pt) + bb
        h = rotate / 180.0 * math.pi
        A = torch.tensor([[math.cos(h), math.sin(h)], [-math.sin(h), math.cos(h)]], dtype=torch.float32)
        bb = torch.tensor([float(crop[2] - crop[0]), float(crop[3] - crop[1])]) / 2
        bb = A.matmul(-bb) + bb
        pr = A.matmul(pr)
        pt = A.matmul(pt) + bb
        pr3 = np.eye(3)
        pr3[:2, :2] = pr.numpy()
        pt3 = np.zeros(3)
        pt3[:2] = pt.numpy()
        for n in range(N):
            yaw = _RING_YAW_DEG[n % len(_RING_YAW_DEG)] + float(randn(1)[0])
            rots[b, n] = _rz(yaw) @ _CAM_TO_EGO_AXES
            a = math.radians(yaw)
            trans[b, n] = np.array([1.5 * math.cos(a), 0.5 * math.sin(a), 1.6]) + 0.05 * randn(3)
            intr[b, n] = K
            post_rots[b, n] = pr3
            post_trans[b, n] = pt3

    t = lambda a: torch.from_numpy(a.astype(np.float32))
    return {"rots": t(rots), "trans": t(trans), "intrins": t(intr),
            "post_rots": t(post_rots), "post_trans": t(post_trans)}


def make_depthnet_out(cfg: LiftSplatConfig, seed: int = 0):
    """N(0,1) stand-in for the depthnet 1x1-conv output, f32[B*N, D+C, fH, fW] (models.py:56)."""
    g = torch.Generator().manual_seed(2000 + seed)
    fH, fW = cfg.fHW
    return torch.randn(cfg.B * cfg.N, cfg.D + cfg.C, fH, fW, generator=g, dtype=torch.float32)


def make_bev_grad(cfg: LiftSplatConfig, seed: int = 0):
    """N(0,1) upstream gradient of the BEV tensor, f32[B, Z*C, X, Y] (models.py:240-244)."""
    g = torch.Generator().manual_seed(3000 + seed)
    X, Y, Z = cfg.nx
    return torch.randn(cfg.B, Z * cfg.C, X, Y, generator=g, dtype=torch.float32)


def make_batch(cfg: LiftSplatConfig, seed: int = 0, aug: str = "train"):
    d = make_calibration(cfg, seed, aug)
    d["depthnet_out"] = make_depthnet_out(cfg, seed)
    return d
