import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def has_gpu():
    try:
        from rapiddoc_b200 import _lib
        return _lib.load().rdb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a visible sm_100 device: on a CPU box they are skipped (not failed) even without `-m "not gpu"`."""
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (gpu-marked test)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
