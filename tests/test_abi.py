"""The C-ABI library loads without a GPU and exports every symbol include/hipgp_b200.h declares; the ctypes
prototypes in hipgp_b200/_lib.py cover exactly that set.  (No compute calls here.)"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    hdr = open(os.path.join(ROOT, "include", "hipgp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(hipgp_[a-z0-9_]+)\s*\(", hdr)))


def test_header_matches_ctypes_prototypes():
    from hipgp_b200 import _lib as L
    assert header_symbols() == sorted(L.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from hipgp_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    L.declare(lib)
    assert lib.hipgp_version() >= 100


def test_product_refuses_cpu():
    import torch
    from hipgp_b200.plan import Plan
    from hipgp_b200 import kernels as hk
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Plan([8, 8], torch.float32, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hk.SqExp()(torch.zeros(2, 2), torch.zeros(3, 2), (1.0, 1.0))


def test_product_never_imports_oracle():
    """Nothing under hipgp_b200/ may reference oracle/ or the emulation library."""
    pkg = os.path.join(ROOT, "hipgp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f
                assert "emu_build" not in src and "libhipgp_emu" not in src, f
