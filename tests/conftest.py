import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes
        from opencloth_b200 import _abi
        lib = _abi.load()
        return b"devices=0" not in lib.oc_version()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; if someone runs the whole suite on a CPU box, skip them
    # loudly instead of failing on OC_ERR_NO_DEVICE.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device: GPU tests run with -m gpu on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checkers (oracle, emulator) and the product library if they are missing."""
    import helpers
    helpers.ensure_built()
