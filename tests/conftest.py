import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device AND the built library; without either they are skipped, not failed (the
    product itself still fails loudly: Engine raises TpcError)."""
    import torch
    lib = os.path.join(ROOT, "transformer-pointer-critic_b200", "libtpc.so")
    if torch.cuda.is_available() and os.path.exists(lib):
        return
    why = "no CUDA device" if not torch.cuda.is_available() else "libtpc.so not built"
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
