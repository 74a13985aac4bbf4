#!/usr/bin/env python
"""Stage the UNMODIFIED reference sources of the lift-splat path under baseline/_ref/ (git-ignored, travels to the GPU box).

    python scripts/install_reference.py            # needs /root/reference (build container only)

The reference has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference`
has nothing to build; the equivalent for a pure-Python tree is a verbatim copy of its `src/` package.  Nothing under
baseline/_ref is tracked by git and nothing in `lss_carla_b200/` imports it: it is loaded (with the third-party stubs of
`lss_carla_b200.refload`) only by the `reference_gpu`-gated tests, by `scripts/gpu_reference_probe.py` and by
`bench.py`'s `gpu_reference` leg, which time / compare the reference's own classes on the B200.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LSS_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("src/__init__.py", "src/models.py", "src/tools.py", "src/explore.py")


def install(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "src", "models.py")):
        if verbose:
            print(f"{SRC} not present: nothing staged")
        return False
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Verbatim copy of the reference's src/ files for the lift-splat path (git-ignored).\n"
                "Made by scripts/install_reference.py; never imported by lss_carla_b200.\n")
    if verbose:
        print(f"staged {len(FILES)} reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
