import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "sub-gc_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests skip (instead of failing with 'no NVIDIA driver') on a machine without CUDA or without the built library."""
    import torch
    from subgc import _lib
    reason = None
    if not torch.cuda.is_available():
        reason = "needs a CUDA device"
    elif not os.path.isfile(_lib.LIB_PATH):
        reason = "libsubgc_b200.so has not been built (python -m subgc.build)"
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
