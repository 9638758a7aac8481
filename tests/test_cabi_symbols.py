"""CPU: the C-ABI library loads and exports every symbol include/vlb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vlb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vlb200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vlb200.h but not exported"
    # the ctypes binding covers exactly the declared ABI
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.vlb200_abi_version() == 1


def test_missing_extension_fails_loudly(monkeypatch):
    import vlrlhf_b200  # noqa: F401
    from vlrlhf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libvlb200.so")
    with pytest.raises(ImportError):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "vlrlhf_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):   # (the reference tree is cited in comments; no Python file may open it either)
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "/root/reference" in src:
                    offenders.append(os.path.relpath(os.path.join(dirpath, f), root))
    assert not offenders, offenders
