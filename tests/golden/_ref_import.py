"""Import the REAL reference (`/root/reference/src/models.py`, `tools.py`) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by `make_golden.py` (fixture generation) and by the
`reference_available`-gated tests.  `/root/reference` does not exist on the GPU box, so nothing
in the `-m gpu` tests, `smoke()` or `bench.py` imports this module.

The reference's import chain needs packages that are not installed offline
(efficientnet_pytorch, pyquaternion, matplotlib, nuscenes).  We register
empty stand-in modules for them; none of them is touched by the lift-splat path
(models.py:157-254, tools.py:174-219).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("LSS_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "src", "models.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Return (models, tools) modules of the reference, importing with third-party stubs."""
    import torch
    from torch import nn

    class _Trunk(nn.Module):  # stand-in for efficientnet_pytorch.EfficientNet (never executed)
        @classmethod
        def from_pretrained(cls, name):
            return cls()

    _stub("efficientnet_pytorch", EfficientNet=_Trunk)
    _stub("pyquaternion", Quaternion=object)
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    _stub("nuscenes")
    _stub("nuscenes.utils")
    _stub("nuscenes.utils.data_classes", LidarPointCloud=object)
    _stub("nuscenes.utils.geometry_utils", transform_matrix=None)
    _stub("nuscenes.map_expansion")
    _stub("nuscenes.map_expansion.map_api", NuScenesMap=object)
    try:
        import tqdm  # noqa: F401
    except Exception:  # pragma: no cover
        _stub("tqdm", tqdm=lambda x, **k: x)

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # the reference package is called `src`; import it under that name
    import importlib
    models = importlib.import_module("src.models")
    tools = importlib.import_module("src.tools")
    return models, tools


def reference_geometry_cpu(model, rots, trans, intrins, post_rots, post_trans):
    """Run the reference's OWN `get_geometry` (models.py:170-190) on a CPU-only box.

    The method hard-codes `.cpu()` / `.cuda()` hops around the two 3x3 inverses (models.py:180,186).
    On a box without a GPU we make `Tensor.cuda` the identity for the duration of the call, so the
    code that executes is the unmodified reference method."""
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        return model.get_geometry(rots, trans, intrins, post_rots, post_trans)
    finally:
        torch.Tensor.cuda = orig
