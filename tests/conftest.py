import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def locked_blob():
    with open(os.path.join(ROOT, "robogym_b200", "assets", "dactyl_locked.rgm"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def locked_names():
    import json

    with open(os.path.join(ROOT, "robogym_b200", "assets", "dactyl_locked.names.json")) as f:
        return json.load(f)
