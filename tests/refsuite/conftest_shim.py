"""pytest plugin used when running the REFERENCE's own test files against the shim (tests/test_reference_suite.py):
installs robogym_b200.mujoco_py_shim as `mujoco_py` with the fp64 oracle engine (CPU tier)."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "stubs"), os.environ.get("ROBOGYM_REFERENCE", "/root/reference")):
    if p not in sys.path:
        sys.path.insert(0, p)

import robogym_b200.mujoco_py_shim as shim  # noqa: E402

shim.install()
if os.environ.get("RG_SHIM_ENGINE", "oracle") == "oracle":
    from oracle_engine import OracleEngine  # noqa: E402

    shim.set_engine_factory(OracleEngine)
