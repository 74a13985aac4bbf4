#!/usr/bin/env python
"""Run scripts/sanitize_case.py (+ the run-plan parity tests' long-voxel case) against a library built with
-DLSS_DEVICE_ASSERTS: bounds, list-integrity and capacity checks inside the run-plan kernels.

compute-sanitizer is CLOSED on this GPU pool ("runs under it have left GPUs needing a reset"), so this build is the
memcheck substitute of SURVEY.md section 5; racecheck's role is played by the bit-exact comparison of the deterministic
mode against the sequential oracle and by run-to-run identity.  Summary: profiles/r02_sanitizer.md."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lss_carla_b200 import _lib  # noqa: E402

so = _lib.build_library(out=os.path.join(ROOT, "lss_carla_b200", "liblss_b200_asserts.so"), defines=("LSS_DEVICE_ASSERTS",))
_lib.SO_PATH = so
_lib._lib = None
print("library:", so, flush=True)
runpy.run_path(os.path.join(ROOT, "scripts", "sanitize_case.py"), run_name="__main__")
import pytest  # noqa: E402
sys.exit(pytest.main(["-x", "-q", os.path.join(ROOT, "tests", "test_runplan_gpu.py"), "-k", "tiny or cfg1 or long_voxels or graph_replay"]))
