"""CPU checks of the measurement contract and of the PyTorch-side harness pieces (no GPU, no CUDA kernels)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the reference's ATen op chain on the host cores) prints ONE JSON line with the
    keys the driver reads; under torchrun only rank 0 prints."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "cfg1"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "bev_pool_mpoints_per_s_fwd_bwd" and d["unit"] == "Mpoints/s"
    assert d["value"] > 0 and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a non-zero rank stays silent and exits 0
    env.update(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1", "--workload", "cfg1"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_trunk_stand_ins_have_the_reference_shapes():
    """The PyTorch stand-ins that drive the path in the training harness: (D+C) channels at 1/16 resolution in, one
    logit map of the BEV size out; parameter count of the order of the reference's (12.6 M used parameters)."""
    from lss_carla_b200.trunk import BevEncode, CamEncode
    ce = CamEncode(41, 64).eval()
    with torch.no_grad():
        dn = ce.depthnet(ce.get_eff_depth(torch.randn(1, 3, 128, 352)))
    assert tuple(dn.shape) == (1, 105, 8, 22)
    be = BevEncode(64, 1).eval()
    with torch.no_grad():
        assert tuple(be(torch.randn(1, 64, 200, 200)).shape) == (1, 1, 200, 200)
    n = sum(p.numel() for p in ce.parameters()) + sum(p.numel() for p in be.parameters())
    assert 11e6 < n < 14e6


def test_model_constants_and_state_dict_keys_match_the_reference_contract():
    """dx / bx / nx (int64) / frustum are no-grad Parameters with the reference's names, dtypes and values (fixtures)."""
    import numpy as np
    from conftest import load_golden
    from lss_carla_b200 import models
    from lss_carla_b200.synthetic import CONFIGS
    cfg = CONFIGS["cfg1"]
    g = load_golden("cfg1_train_s0")
    m = models.LiftSplatShoot(cfg.grid_conf, cfg.data_aug_conf, outC=1, camencode=torch.nn.Identity(), bevencode=torch.nn.Identity())
    sd = m.state_dict()
    for k in ("dx", "bx", "nx", "frustum"):
        assert k in sd and not getattr(m, k).requires_grad
    assert sd["nx"].dtype == torch.int64 and sd["dx"].dtype == torch.float32
    assert np.array_equal(sd["dx"].numpy(), g["dx"]) and np.array_equal(sd["bx"].numpy(), g["bx"]) and np.array_equal(sd["nx"].numpy(), g["nx"])
    assert np.array_equal(sd["frustum"].numpy(), g["frustum"])
    assert m.D == 41 and m.downsample == 16 and m.camC == 64 and m.use_quickcumsum is True


def test_calibration_key_is_a_function_of_the_host_bytes():
    """api.calibration_key (plan cache of api.LiftSplat): equal for equal calibration, different for any changed float, None
    for device tensors (no hashing round trip)."""
    import torch
    from lss_carla_b200 import api
    from lss_carla_b200.synthetic import CONFIGS, make_batch
    names = ("rots", "trans", "intrins", "post_rots", "post_trans")
    b = make_batch(CONFIGS["tiny"], 0, "train")
    k = api.calibration_key(*[b[n] for n in names])
    assert isinstance(k, bytes) and k == api.calibration_key(*[b[n].clone() for n in names])
    assert k != api.calibration_key(*[make_batch(CONFIGS["tiny"], 1, "train")[n] for n in names])
    c = {n: b[n].clone() for n in names}
    c["post_trans"][0, 0, 0] += 1.0                      # one crop offset changed
    assert k != api.calibration_key(*[c[n] for n in names])
    assert api.calibration_key(*[b[n].to("meta") if False else b[n] for n in names]) == k
