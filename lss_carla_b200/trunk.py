"""Plain-PyTorch stand-ins for the parts of LSS that stay in PyTorch (north_star): camera trunk and BEV encoder.

NOT part of the hot path and not a product of this repository: the reference builds its camera encoder on the
third-party `efficientnet_pytorch` package (src/models.py:9, :43), which cannot be installed offline, so the
training-throughput harness (`train_bench.py`) and the stand-alone `models.LiftSplatShoot` need something of
the same shape and cost to drive the lift-splat path with.  `EffNetB0Features` is an EfficientNet-B0 feature
extractor written from the published architecture table (MBConv stages, squeeze-excite, swish, stochastic depth),
random-init (BASELINE config 1 says random-init); it returns the /16 (112 ch) and /32 (320 ch) maps the
reference fuses (src/models.py:63-84).  `CamEncode` / `BevEncode` keep the attribute and method names the
drop-in touches (`get_eff_depth`, `dropout`, `depthnet`, `get_depth_dist`, `get_depth_feat`).
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn import functional as F

# (expand ratio, kernel, stride, out channels, repeats) of EfficientNet-B0
_B0_STAGES = ((1, 3, 1, 16, 1), (6, 3, 2, 24, 2), (6, 5, 2, 40, 2), (6, 3, 2, 80, 3),
              (6, 5, 1, 112, 3), (6, 5, 2, 192, 4), (6, 3, 1, 320, 1))


class _SamePadConv(nn.Conv2d):
    """Convolution with TensorFlow-style 'same' padding (asymmetric when needed)."""

    def forward(self, x):
        ih, iw = x.shape[-2:]
        kh, kw = self.kernel_size
        sh, sw = self.stride
        ph = max((math.ceil(ih / sh) - 1) * sh + kh - ih, 0)
        pw = max((math.ceil(iw / sw) - 1) * sw + kw - iw, 0)
        if ph or pw:
            x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        return F.conv2d(x, self.weight, self.bias, self.stride, 0, self.dilation, self.groups)


class _MBConv(nn.Module):
    def __init__(self, cin, cout, expand, k, stride, se_ratio=0.25):
        super().__init__()
        mid = cin * expand
        self.use_res = stride == 1 and cin == cout
        layers = []
        if expand != 1:
            layers += [_SamePadConv(cin, mid, 1, bias=False), nn.BatchNorm2d(mid, momentum=0.01, eps=1e-3), nn.SiLU()]
        layers += [_SamePadConv(mid, mid, k, stride=stride, groups=mid, bias=False),
                   nn.BatchNorm2d(mid, momentum=0.01, eps=1e-3), nn.SiLU()]
        self.pre = nn.Sequential(*layers)
        sq = max(1, int(cin * se_ratio))
        self.se_reduce, self.se_expand = nn.Conv2d(mid, sq, 1), nn.Conv2d(sq, mid, 1)
        self.project = nn.Sequential(_SamePadConv(mid, cout, 1, bias=False), nn.BatchNorm2d(cout, momentum=0.01, eps=1e-3))

    def forward(self, x, drop_connect_rate=0.0):
        y = self.pre(x)
        y = y * torch.sigmoid(self.se_expand(F.silu(self.se_reduce(y.mean((2, 3), keepdim=True)))))
        y = self.project(y)
        if self.use_res:
            if self.training and drop_connect_rate:
                keep = 1.0 - drop_connect_rate
                mask = torch.floor(keep + torch.rand(y.shape[0], 1, 1, 1, dtype=y.dtype, device=y.device))
                y = y / keep * mask
            y = y + x
        return y


class EffNetB0Features(nn.Module):
    """EfficientNet-B0 up to the last MBConv stage; forward returns (stride-16 map, stride-32 map)."""

    def __init__(self, drop_connect_rate=0.2):
        super().__init__()
        self.drop_connect_rate = drop_connect_rate
        self.stem = nn.Sequential(_SamePadConv(3, 32, 3, stride=2, bias=False), nn.BatchNorm2d(32, momentum=0.01, eps=1e-3), nn.SiLU())
        blocks, cin = [], 32
        for expand, k, stride, cout, reps in _B0_STAGES:
            for r in range(reps):
                blocks.append(_MBConv(cin, cout, expand, k, stride if r == 0 else 1))
                cin = cout
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        x = self.stem(x)
        taps, prev = [], x
        n = len(self.blocks)
        for i, blk in enumerate(self.blocks):
            x = blk(x, self.drop_connect_rate * i / n)
            if prev.shape[2] > x.shape[2]:
                taps.append(prev)                      # last map of every resolution
            prev = x
        taps.append(x)
        return taps[-2], taps[-1]                      # /16 (112 ch), /32 (320 ch)


class _Fuse(nn.Module):
    """Upsample the coarse map, concatenate with the fine one, two 3x3 conv-BN-ReLU."""

    def __init__(self, cin, cout, scale):
        super().__init__()
        self.scale = scale
        self.conv = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                                  nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, coarse, fine):
        up = F.interpolate(coarse, scale_factor=self.scale, mode="bilinear", align_corners=True)
        return self.conv(torch.cat([fine, up], dim=1))


class CamEncode(nn.Module):
    """Image -> (D + C)-channel map at 1/16 resolution; same method names as the reference class."""

    def __init__(self, D, C, downsample=16):
        super().__init__()
        self.D, self.C = D, C
        self.trunk = EffNetB0Features()
        self.up1 = _Fuse(320 + 112, 512, 2)
        self.dropout = nn.Dropout(0.2)
        self.depthnet = nn.Conv2d(512, D + C, kernel_size=1)

    def get_depth_dist(self, x, eps=1e-20):
        return x.softmax(dim=1)

    def get_eff_depth(self, x):
        s16, s32 = self.trunk(x)
        return self.up1(s32, s16)

    def get_depth_feat(self, x):
        """Materialising lift (API compatibility only; the fused path never calls it)."""
        x = self.depthnet(self.dropout(self.get_eff_depth(x)))
        depth = self.get_depth_dist(x[:, :self.D])
        return depth, depth.unsqueeze(1) * x[:, self.D:self.D + self.C].unsqueeze(2)

    def forward(self, x):
        return self.get_depth_feat(x)[1]


class BevEncode(nn.Module):
    """ResNet-18 stages 1-3 over the BEV grid + two upsampling heads -> outC logits."""

    def __init__(self, inC, outC):
        super().__init__()
        from torchvision.models.resnet import resnet18
        r = resnet18(weights=None, zero_init_residual=True)
        self.conv1 = nn.Conv2d(inC, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1, self.relu = r.bn1, r.relu
        self.layer1, self.layer2, self.layer3 = r.layer1, r.layer2, r.layer3
        self.up1 = _Fuse(64 + 256, 256, 4)
        self.dropout = nn.Dropout2d(0.1)
        self.up2 = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
                                 nn.Conv2d(256, 128, 3, padding=1, bias=False), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
                                 nn.Conv2d(128, outC, 1))

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        x1 = self.layer1(x)
        x = self.layer3(self.layer2(x1))
        return self.up2(self.dropout(self.up1(x, x1)))
