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


class mocked_ops:
    """Swap `vlrlhf_b200.ops` for tests/mock_ops.py (CPU stand-in for the C ABI) and re-import the modules that bind it at import
    time; restores the real modules afterwards.  Both `sys.modules` and the package attribute are swapped: `from . import ops`
    resolves the attribute first once the real module has been imported by an earlier test file."""

    def __init__(self, *reimport):
        self.names = ("vlrlhf_b200.ops",) + tuple(reimport)

    def __enter__(self):
        import importlib
        import vlrlhf_b200
        from tests import mock_ops
        self.pkg = vlrlhf_b200
        self.saved = {k: sys.modules.get(k) for k in self.names}
        self.saved_attr = {k: getattr(vlrlhf_b200, k.split(".")[-1], None) for k in self.names}
        sys.modules["vlrlhf_b200.ops"] = mock_ops
        vlrlhf_b200.ops = mock_ops
        for k in self.names[1:]:
            sys.modules.pop(k, None)
            if hasattr(vlrlhf_b200, k.split(".")[-1]):
                delattr(vlrlhf_b200, k.split(".")[-1])
        self.modules = {k.split(".")[-1]: importlib.import_module(k) for k in self.names[1:]}
        return self

    def __exit__(self, *exc):
        for k in self.names:
            attr = k.split(".")[-1]
            if self.saved[k] is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = self.saved[k]
            if self.saved_attr[k] is None:
                if hasattr(self.pkg, attr):
                    delattr(self.pkg, attr)
            else:
                setattr(self.pkg, attr, self.saved_attr[k])
        return False
