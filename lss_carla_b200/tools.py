"""Host-side mirror of the reference's `src/tools.py` symbols that sit on the lift-splat path.

    gen_dx_bx       src/tools.py:174-179   (construction-time constants; plain torch like the reference)
    cumsum_trick    src/tools.py:182-190   -> CUDA (lss_quickcumsum_*)
    QuickCumsum     src/tools.py:193-219   -> CUDA (lss_quickcumsum_*)
"""
import torch

from .ops import QuickCumsum, cumsum_trick  # noqa: F401  (re-exported under the reference's names)


def gen_dx_bx(xbound, ybound, zbound):
    """Voxel size, first-bin centre and bin count of the BEV grid: float32[3], float32[3], int64[3].
    The bin count truncates (hi - lo) / step toward zero, as `torch.LongTensor(float)` does."""
    rows = (xbound, ybound, zbound)
    dx = torch.tensor([float(r[2]) for r in rows], dtype=torch.float32)
    bx = torch.tensor([float(r[0]) + float(r[2]) / 2.0 for r in rows], dtype=torch.float32)
    nx = torch.tensor([int((float(r[1]) - float(r[0])) / float(r[2])) for r in rows], dtype=torch.int64)
    return dx, bx, nx
