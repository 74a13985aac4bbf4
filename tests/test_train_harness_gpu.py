"""GPU tests of the training-step harness behind north_star's second metric (lss_carla_b200/harness.py: forward, BCE with
pos_weight 2.13, backward, clip 5.0, Adam -- train_simbev.py:231-248) with the lift-splat of liblss_b200 in the middle, in float32
and under bfloat16 autocast (the lift-splat then reads the bfloat16 depthnet output, `lss_lift_prepare_bf16`).
Nothing here reads /root/reference."""
import math

import pytest
import torch

from lss_carla_b200.harness import TrainStep, make_train_batch
from lss_carla_b200.synthetic import CONFIGS

pytestmark = pytest.mark.gpu


def _steps(amp, n, seen=None):
    dev = torch.device("cuda:0")
    cfg = CONFIGS["cfg1"]
    step = TrainStep(cfg, dev, seed=3, amp=amp)
    if seen is not None:
        step.model.camencode.depthnet.register_forward_hook(lambda m, i, o: seen.append(o.dtype))
    batch = make_train_batch(cfg, 2, 0, dev)
    losses = [step(batch).item() for _ in range(n)]
    return step, losses


def test_train_step_float32_learns_one_batch():
    step, losses = _steps(False, 8)
    assert all(math.isfinite(v) for v in losses)
    assert min(losses[1:]) < losses[0], losses
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in step.model.parameters() if p.requires_grad)


def test_train_step_bf16_autocast_feeds_the_lift_splat_bf16_and_tracks_float32():
    seen = []
    step, amp = _steps(True, 8, seen)
    _, ref = _steps(False, 1)
    assert seen and all(d == torch.bfloat16 for d in seen)              # the bf16 entry point is the one that ran
    assert all(math.isfinite(v) for v in amp)
    # first step, same weights and batch: bfloat16 rounding of the trunk's activations only (8 mantissa bits, mean-reduced loss)
    assert abs(amp[0] - ref[0]) <= 3e-2 * abs(ref[0]), (amp[0], ref[0])
    assert min(amp[1:]) < amp[0], amp
    for p in step.model.parameters():                                    # parameters and their gradients stay float32
        if p.requires_grad:
            assert p.dtype == torch.float32 and p.grad.dtype == torch.float32 and torch.isfinite(p.grad).all()
