"""Puts the reference's UNCHANGED Python wrapper (pure Python: /root/reference/src/power_grid_model) under baseline/_ref/ --
git-ignored, not part of the product, travels to the GPU box with the snapshot -- so that tests/test_reference_wrapper.py and
tests/test_gpu_reference_wrapper.py can drive libpgm_b200.so through it (the wrapper resolves the whole PGM_* C API by name when
it loads its core library).  The reference's C++ core is NOT installed: it cannot be built in this image."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", "src", "power_grid_model")
DST = os.path.join(ROOT, "baseline", "_ref", "power_grid_model")

if not os.path.isdir(SRC):
    sys.exit(f"{SRC} not found")
shutil.rmtree(DST, ignore_errors=True)
shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.license"))
print("installed", DST)
