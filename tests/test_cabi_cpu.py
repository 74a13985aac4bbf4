"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/lss_b200.h
declares, host-only entry points work, and the Python layer refuses to run without CUDA (no fallback)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from lss_carla_b200 import _lib, ops
from lss_carla_b200.synthetic import CONFIGS
from lss_carla_b200.tools import gen_dx_bx
from oracle import lss_oracle as O


@pytest.fixture(scope="module")
def L():
    if not os.path.isfile(_lib.SO_PATH):
        _lib.build_library()
    return _lib.lib()


def test_every_header_symbol_is_exported_and_bound(L):
    names = _lib.header_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/lss_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert L.lss_version() == 100
    assert b"workspace" in L.lss_status_string(-5)


def test_struct_sizes_match_header():
    # lss_problem: 9 int32 + 6 float = 60 bytes; lss_plan_layout: 3 int32 (+pad) + int64 + 15 size_t + int64
    assert C.sizeof(_lib.LssProblem) == 60
    assert C.sizeof(_lib.LssPlanLayout) == 16 + 8 + 15 * 8 + 8
    # field order of the ctypes mirror == field order of the C struct
    import re
    hdr = open(_lib.HEADER).read()
    body = hdr[hdr.index("typedef struct lss_plan_layout {"):hdr.index("} lss_plan_layout;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = re.findall(r"(?:int32_t|int64_t|size_t)\s+([a-z_0-9]+)\s*;", body)
    assert c_fields == [f[0] for f in _lib.LssPlanLayout._fields_]
    body = hdr[hdr.index("typedef struct lss_runplan_layout {"):hdr.index("} lss_runplan_layout;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = re.findall(r"(?:int32_t|int64_t|size_t)\s+([a-z_0-9]+)\s*;", body)
    assert c_fields == [f[0] for f in _lib.LssRunplanLayout._fields_]
    assert C.sizeof(_lib.LssRunplanLayout) == 8 * len(c_fields)


def test_runplan_layout_and_options_host_only(L):
    cfg = CONFIGS["cfg2"]
    p = _problem(cfg)
    lay = _lib.LssRunplanLayout()
    assert L.lss_runplan_layout_init(C.byref(p.c), C.byref(lay)) == 0
    assert lay.n_points == cfg.points and lay.n_runs == cfg.points // cfg.fHW[0] and lay.n_voxels == 8 * 200 * 200
    offs = [getattr(lay, f[0]) for f in _lib.LssRunplanLayout._fields_ if f[0].startswith("off_")]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and offs[-1] < lay.bytes
    assert lay.bytes - lay.off_head >= 8 * lay.n_voxels and lay.off_zero_done > lay.off_counters   # counters, progress, heads: the tail lss_runplan_reset clears
    assert ops.runplan_supported(p) and ops.runplan_raw_supported(p)
    tiny_cam = ops.Problem.from_grid(8, 6, 2, 4, 2, 64, *gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound))   # an index CTA would span 16 cameras
    assert ops.runplan_supported(tiny_cam) and not ops.runplan_raw_supported(tiny_cam)
    odd = ops.Problem.from_grid(1, 1, 8, 40, 4, 64, *gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound))   # a run must fit a warp
    assert L.lss_runplan_layout_init(C.byref(odd.c), C.byref(lay)) == -3 and not ops.runplan_supported(odd)
    c48 = ops.Problem.from_grid(1, 1, 8, 8, 4, 48, *gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound))
    assert not ops.runplan_supported(c48)
    # process-wide options: the library's only global state
    assert L.lss_get_option(0) == 1
    assert L.lss_set_option(0, 0) == 0 and L.lss_get_option(0) == 0
    assert L.lss_set_option(0, 1) == 0 and L.lss_set_option(7, 1) == -1
    null = C.c_void_p(0)
    assert L.lss_liftsplat_fwd_cl(C.byref(p.c), None, null, null, null, null, 0, null) == -5
    assert L.lss_liftsplat_fwd_cl(C.byref(p.c), C.byref(lay), C.c_void_p(256), C.c_void_p(256), C.c_void_p(256), C.c_void_p(256), 3, null) == -1
    assert L.lss_bev_zero(C.byref(p.c), null, 0, 1, null) == -1


def _problem(cfg):
    dx, bx, nx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    fH, fW = cfg.fHW
    return ops.Problem.from_grid(cfg.B, cfg.N, cfg.D, fH, fW, cfg.C, dx, bx, nx)


def test_problem_lo_matches_oracle():
    for name in ("cfg1", "cfg4", "tiny"):
        cfg = CONFIGS[name]
        p = _problem(cfg)
        dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
        lo = bx - dx / np.float32(2)
        assert np.array_equal(np.array(p.lo, np.float32), lo)
        assert np.array_equal(np.array(p.dx, np.float32), dx)
        assert tuple(p.nx) == tuple(int(v) for v in nx)
        tdx, tbx, tnx = gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
        assert tnx.dtype == torch.int64 and tdx.dtype == torch.float32
        assert np.array_equal(tbx.numpy(), bx)


def test_plan_layout_host_only(L):
    cfg = CONFIGS["cfg2"]
    p = _problem(cfg)
    lay = _lib.LssPlanLayout()
    assert L.lss_plan_layout_init(C.byref(p.c), 0, C.byref(lay)) == 0
    assert lay.tile_cols == 200 and lay.tiles_per_row == 1 and lay.n_tiles == 8 * 200
    assert lay.n_points == cfg.points == 346368
    assert lay.bytes >= 2 * 4 * cfg.points
    assert lay.n_rows_cap == cfg.points      # compact rows are indexed by bucket slot: below the number of kept points
    offs = [getattr(lay, f[0]) for f in _lib.LssPlanLayout._fields_ if f[0].startswith("off_")]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and offs[-1] < lay.bytes
    assert L.lss_plan_layout_init(C.byref(p.c), 56, C.byref(lay)) == 0
    assert lay.tiles_per_row == 4 and lay.n_tiles == 8 * 200 * 4
    assert L.lss_plan_layout_init(C.byref(p.c), 13, C.byref(lay)) == -3      # not a multiple of 8
    lim = _lib.LssLimits()
    L.lss_get_limits(C.byref(lim))
    assert lim.max_points_per_sample == 1 << 20 and lim.max_depth_bins == 256


def test_bad_arguments_return_status_codes(L):
    cfg = CONFIGS["tiny"]
    p = _problem(cfg)
    null = C.c_void_p(0)
    assert L.lss_geometry(C.byref(p.c), null, null, null, null, null, null, null) == -1
    assert L.lss_lift_prepare(C.byref(p.c), null, null, null, null, null) == -1
    bad = _lib.LssProblem()
    lay = _lib.LssPlanLayout()
    assert L.lss_plan_layout_init(C.byref(bad), 0, C.byref(lay)) == -1       # zero dims
    big = _problem(cfg)
    big.c.D = 1 << 20
    assert L.lss_plan_layout_init(C.byref(big.c), 0, C.byref(lay)) == -3     # > 2^20 points per sample
    assert L.lss_splat_fwd(C.byref(p.c), None, null, null, null, null, null, null, 0, 0, 0, 0, 0, 0, null) == -5
    assert L.lss_quickcumsum_scratch_elems(5000) >= 5000 + 5


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of computing with ATen."""
    cfg = CONFIGS["tiny"]
    p = _problem(cfg)
    fH, fW = cfg.fHW
    x = torch.zeros(cfg.B * cfg.N, cfg.D + cfg.C, fH, fW)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.lift_prepare(p, x)
    with pytest.raises(RuntimeError):
        ops.QuickCumsum.apply(torch.zeros(4, 8), torch.zeros(4, 4, dtype=torch.long), torch.zeros(4, dtype=torch.long))
    import lss_carla_b200.ops as ops_src
    src = open(ops_src.__file__).read()
    assert "oracle" not in src.replace("# oracle", ""), "product code must not import the oracle"


def test_pipe_stage_rejects_bad_arguments_without_touching_cuda():
    """lss_pipe_stage validates its arguments before any CUDA call (more than 4 copies, copies without arrays)."""
    import ctypes as C
    from lss_carla_b200 import _lib
    L = _lib.lib()
    assert L.lss_pipe_stage(None, None, None, 5, None, None, None, None) != 0
    assert L.lss_pipe_stage(None, None, None, 1, None, None, None, None) != 0
    assert L.lss_pipe_stage(None, None, None, -1, None, None, None, None) != 0
    assert L.lss_pipe_event_synchronize(None) != 0
    assert L.lss_pipe_event_destroy(None) == 0
