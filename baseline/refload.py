"""Load the REAL reference classes (`src/models.py`, `src/tools.py`) for measurement and parity work on the GPU box.

MEASUREMENT / TEST INFRASTRUCTURE ONLY -- never imported by `lss_carla_b200`.

`/root/reference` exists only in the build container.  `scripts/install_reference.py` (called by
`__graft_entry__.build()`) stages a verbatim, git-ignored copy of the path's source files under `baseline/_ref/`,
which travels to the GPU box with the snapshot.  This loader imports that copy under its own package name `src`
with empty stand-ins for the third-party packages the reference imports at module top but never touches on the
lift-splat path (efficientnet_pytorch, pyquaternion, matplotlib, nuscenes -- SURVEY.md Appendix A).
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.environ.get("LSS_REFERENCE_ROOT", ""), os.path.join(HERE, "_ref"), "/root/reference")


def reference_root():
    for r in CANDIDATES:
        if r and os.path.isfile(os.path.join(r, "src", "models.py")):
            return r
    return None


def reference_available():
    return reference_root() is not None


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference(trunk_cls=None):
    """(models, tools) modules of the reference.  `trunk_cls`: class to stand in for efficientnet_pytorch.EfficientNet
    (needs a `from_pretrained(name)` classmethod and the attributes CamEncode.get_eff_depth touches, models.py:63-84);
    default: an empty module, enough for everything except running the camera trunk."""
    from torch import nn
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources not found (baseline/_ref missing: run scripts/install_reference.py in the build container)")

    class _Trunk(nn.Module):
        @classmethod
        def from_pretrained(cls, name):
            return cls()

    eff = _stub("efficientnet_pytorch")
    eff.EfficientNet = trunk_cls or getattr(eff, "EfficientNet", None) or _Trunk
    _stub("pyquaternion", Quaternion=object)
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("nuscenes")
    _stub("nuscenes.utils")
    _stub("nuscenes.utils.data_classes", LidarPointCloud=object)
    _stub("nuscenes.utils.geometry_utils", transform_matrix=None)
    _stub("nuscenes.map_expansion")
    _stub("nuscenes.map_expansion.map_api", NuScenesMap=object)
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("src.models"), importlib.import_module("src.tools")


def build_liftsplat_model(models, cfg, device=None):
    """A reference `LiftSplatShoot` whose camera trunk is the identity: `get_cam_feats` then accepts a depthnet-shaped
    tensor [B, N, D+C, fH, fW] and every line of geometry / lift / splat that runs is the reference's own
    (models.py:170-254).  Same construction as tests/golden/make_golden.py."""
    import torch
    m = models.LiftSplatShoot(cfg.grid_conf, cfg.data_aug_conf, outC=1)
    m.camencode.get_eff_depth = lambda x: x
    m.camencode.dropout = torch.nn.Identity()
    m.camencode.depthnet = torch.nn.Identity()
    m.downsample = 1
    m.camC = cfg.C
    m.camencode.C = cfg.C
    return m.to(device) if device is not None else m
