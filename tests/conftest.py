import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_sessionstart(session):
    """The host logic calls native host routines of libvlb200 (DDPO diff): make sure the in-tree library exists
    (nvcc cross-compiles here; on the GPU box the prebuilt .so travelled with the snapshot)."""
    if not os.path.exists(os.path.join(ROOT, "vlrlhf_b200", "libvlb200.so")):
        import __graft_entry__ as g
        g.build()


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
