"""ctypes binding of liblss_b200.so (the C ABI declared in include/lss_b200.h).

The library is built in-tree by ``build_library()`` (``nvcc -gencode arch=compute_100a,code=sm_100a``).
There is no CPU fallback: ``lib()`` raises if the shared object is missing, and every compute call
raises ``RuntimeError`` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SO_PATH = os.path.join(_PKG, "liblss_b200.so")
CSRC = os.path.join(_PKG, "csrc")
SOURCES = {"plan.cu": [], "runplan.cu": [], "lift.cu": [], "splat.cu": [], "ops.cu": []}      # source -> extra nvcc flags
HEADER = os.path.join(_ROOT, "include", "lss_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


class LssProblem(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("D", C.c_int32), ("fH", C.c_int32), ("fW", C.c_int32),
                ("C", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
                ("dx", C.c_float * 3), ("lo", C.c_float * 3)]


class LssPlanLayout(C.Structure):
    _fields_ = [("tile_cols", C.c_int32), ("tiles_per_row", C.c_int32), ("n_tiles", C.c_int32),
                ("n_points", C.c_int64), ("off_vox", C.c_size_t), ("off_entries", C.c_size_t),
                ("off_tile_start", C.c_size_t), ("off_segs", C.c_size_t), ("off_tile_nseg", C.c_size_t), ("off_tile_row0", C.c_size_t), ("off_seg_recs", C.c_size_t), ("off_key_count", C.c_size_t), ("off_mixed_recs", C.c_size_t), ("off_prow", C.c_size_t), ("off_counters", C.c_size_t),
                ("n_rows_cap", C.c_int64),
                ("off_tile_count", C.c_size_t), ("off_cursor", C.c_size_t),
                ("off_sync", C.c_size_t), ("bytes", C.c_size_t)]


class LssRunplanLayout(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("n_runs", C.c_int64), ("n_voxels", C.c_int64),
                ("off_prow", C.c_size_t), ("off_sub", C.c_size_t), ("off_sub2", C.c_size_t), ("off_pool", C.c_size_t),
                ("n_rec_cap", C.c_int64), ("off_recs", C.c_size_t), ("off_longs", C.c_size_t), ("off_counters", C.c_size_t), ("off_qcount", C.c_size_t), ("off_zero_done", C.c_size_t), ("off_ready", C.c_size_t),
                ("off_head", C.c_size_t), ("bytes", C.c_size_t)]


class LssLimits(C.Structure):
    _fields_ = [("max_points_per_sample", C.c_int32), ("max_tile_cols", C.c_int32),
                ("max_depth_bins", C.c_int32), ("max_channels", C.c_int32)]


LAYOUT_NCHW, LAYOUT_CHANNELS_LAST = 0, 1
SPLAT_SORTED, SPLAT_SMEM_ATOMIC, SPLAT_RED_GLOBAL = 0, 1, 2
ZERO_ORDERED, ZERO_PRECLEARED = 0, 1
SPLAT_MODES = {"sorted": SPLAT_SORTED, "atomic": SPLAT_SMEM_ATOMIC, "red": SPLAT_RED_GLOBAL}
VARIANTS = {"auto": 0, "warp": 1, "group": 2, "group_gather": 3, "group_store": 4}

_P = C.c_void_p
_PP = C.POINTER(LssProblem)
_PL = C.POINTER(LssPlanLayout)
_PR = C.POINTER(LssRunplanLayout)

# name -> (restype, argtypes); must list every symbol include/lss_b200.h declares
SIGNATURES = {
    "lss_version": (C.c_int, []),
    "lss_status_string": (C.c_char_p, [C.c_int]),
    "lss_get_limits": (None, [C.POINTER(LssLimits)]),
    "lss_plan_layout_init": (C.c_int, [_PP, C.c_int, _PL]),
    "lss_plan_reset": (C.c_int, [_PL, _P, _P]),
    "lss_calib_matrices": (C.c_int, [C.c_int32, _P, _P, _P, _P, _P, _P]),
    "lss_geometry": (C.c_int, [_PP, _P, _P, _P, _P, _P, _P, _P]),
    "lss_voxel_index": (C.c_int, [_PP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "lss_plan_build": (C.c_int, [_PP, _PL, _P, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "lss_plan_build_raw": (C.c_int, [_PP, _PL, _P, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "lss_plan_reference_order": (C.c_int, [_PP, _PL, _P, _P, _P, _P, _P]),
    "lss_lift_prepare": (C.c_int, [_PP, _P, _P, _P, _P, _P]),
    "lss_lift_prepare_bf16": (C.c_int, [_PP, _P, _P, _P, _P, _P]),
    "lss_set_option": (C.c_int, [C.c_int, C.c_int]),
    "lss_get_option": (C.c_int, [C.c_int]),
    "lss_pipe_event_create": (C.c_void_p, []),
    "lss_pipe_event_destroy": (C.c_int, [_P]),
    "lss_pipe_event_synchronize": (C.c_int, [_P]),
    "lss_pipe_stage": (C.c_int, [_P, _P, _P, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_size_t), _P]),
    "lss_pipe_step": (C.c_int, [_P] * 9 + [C.c_size_t, _P, _P, C.c_size_t]),
    "lss_bev_clear": (C.c_int, [_PP, _P, _P]),
    "lss_runplan_layout_init": (C.c_int, [_PP, _PR]),
    "lss_runplan_reset": (C.c_int, [_PR, _P, _P]),
    "lss_runplan_raw_supported": (C.c_int, [_PP]),
    "lss_runplan_build": (C.c_int, [_PP, _PR, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "lss_bev_zero": (C.c_int, [_PP, _P, C.c_int, C.c_int, _P]),
    "lss_liftsplat_prologue": (C.c_int, [_PP, _PR] + [_P] * 15),
    "lss_liftsplat_fwd_cl": (C.c_int, [_PP, _PR, _P, _P, _P, _P, C.c_int, _P]),
    "lss_liftsplat_forward": (C.c_int, [_PP, _PR] + [_P] * 15),
    "lss_liftsplat_forward_persistent": (C.c_int, [_PP, _PR] + [_P] * 15),
    "lss_liftsplat_bwd_cl": (C.c_int, [_PP, _PR, _P, _P, _P, _P, _P, _P]),
    "lss_splat_fwd": (C.c_int, [_PP, _PL, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "lss_splat_bwd": (C.c_int, [_PP, _PL, _P, _P, C.c_int, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "lss_voxel_pooling_fwd": (C.c_int, [_PP, _PL, _P, _P, C.POINTER(C.c_int64), _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "lss_voxel_pooling_bwd": (C.c_int, [_PP, _PL, _P, _P, C.c_int, _P, _P, _P]),
    "lss_quickcumsum_scratch_elems": (C.c_size_t, [C.c_int64]),
    "lss_quickcumsum_runs": (C.c_int, [C.c_int64, _P, _P, _P, _P, _P]),
    "lss_quickcumsum_fwd": (C.c_int, [C.c_int64, C.c_int32, _P, C.c_int64, _P, _P, C.c_int32, _P, _P, _P]),
    "lss_quickcumsum_bwd": (C.c_int, [C.c_int64, C.c_int32, _P, _P, _P, _P]),
}

_lib = None


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_library(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """Compile csrc/*.cu into lss_carla_b200/liblss_b200.so for sm_100a (cross-compiles without a GPU): one object per source
    (in parallel, under build/obj/), then one link.  `out` / `defines`: a second build next to it, e.g. the bounds-checking
    one of scripts/run_with_asserts.py."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in ("common.cuh", "geom.cuh", "lift.cuh")] + [HEADER, os.path.abspath(__file__)]
    target = out or SO_PATH
    if not force and os.path.isfile(target) and all(os.path.getmtime(target) >= os.path.getmtime(d) for d in deps):
        return target
    objdir = os.path.join(_ROOT, "build", "obj", os.path.splitext(os.path.basename(target))[0])
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for name, extra in SOURCES.items():
        obj = os.path.join(objdir, os.path.splitext(name)[0] + ".o")
        cmd = ([nvcc_path()] + NVCC_FLAGS + extra + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
               + ["-c", os.path.join(CSRC, name), "-o", obj])
        jobs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for cmd, obj, proc in jobs:
        log, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + log)
        if verbose:
            print(log)
        objs.append(obj)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] + objs + ["-o", target]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if target == SO_PATH:
        global _lib
        _lib = None
    return target


def lib():
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.isfile(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  lss_carla_b200 has no CPU or PyTorch fallback for the lift-splat path.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().lss_status_string(status).decode()
        raise RuntimeError(f"liblss_b200: {what} failed with status {status}: {msg}")


def header_symbols():
    """Names of the functions declared in include/lss_b200.h (used by the export test)."""
    import re
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lss_[a-z0-9_]+)\s*\(", txt)))
