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


def _lib():
    from hipgp_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return L, L.declare(ctypes.CDLL(L.LIB_PATH))


def _create(lib, dims, dtype=0):
    h = ctypes.c_void_p()
    arr = (ctypes.c_int64 * len(dims))(*dims)
    rc = lib.hipgp_plan_create(len(dims), arr, dtype, 0, ctypes.byref(h))
    return rc, h


def test_plan_create_validates_arguments_without_a_gpu():
    """hipgp_plan_create / _sizes / _embedding are host-only: argument errors come back as a negative status with a
    message (never an abort), and a null handle is refused by every query."""
    L, lib = _lib()
    for dims, dtype, what in (([0, 4], 0, "extents"), ([4, -1], 0, "extents"), ([4, 4], 7, "dtype"), ([3, 3, 3, 3], 0, "at most 3")):
        rc, h = _create(lib, dims, dtype)
        assert rc != 0 and what in lib.hipgp_last_error().decode(), (dims, lib.hipgp_last_error())
    h = ctypes.c_void_p()
    assert lib.hipgp_plan_create(0, (ctypes.c_int64 * 1)(4), 0, 0, ctypes.byref(h)) != 0
    assert lib.hipgp_plan_create(1, None, 0, 0, ctypes.byref(h)) != 0
    M = ctypes.c_int64(); E = ctypes.c_int64()
    assert lib.hipgp_plan_sizes(None, ctypes.byref(M), ctypes.byref(E)) != 0
    assert "null plan" in lib.hipgp_last_error().decode()
    assert lib.hipgp_plan_embedding(None, None, None) != 0
    assert lib.hipgp_plan_launch_count(None, None) != 0
    assert lib.hipgp_plan_set_slab(None, 0, 1) != 0
    assert lib.hipgp_plan_destroy(None) == 0               # destroying nothing is not an error


@pytest.mark.parametrize("dims", [(2,), (100,), (1000, 1000), (300, 300), (128, 128, 64), (1, 50, 1), (17, 1, 40), (512, 512, 512)])
def test_embedding_lengths(dims):
    """M = prod m, M' = prod (2m-2) as in the reference (toeplitz_tensor.py:17-40, unit axes dropped); the embedding
    lengths the transforms run at are 5-smooth, >= 2m-1 (narrow: K, C^-1) and >= (2m-2)+m-1 (wide: R^T, R)."""
    L, lib = _lib()
    rc, h = _create(lib, list(dims))
    assert rc == 0, lib.hipgp_last_error()
    M = ctypes.c_int64(); E = ctypes.c_int64()
    assert lib.hipgp_plan_sizes(h, ctypes.byref(M), ctypes.byref(E)) == 0
    act = [m for m in dims if m > 1]
    expM = 1; expE = 1
    for m in act:
        expM *= m; expE *= 2 * m - 2
    assert (M.value, E.value) == (expM, expE)
    n = len(dims)
    Ln = (ctypes.c_int64 * n)(); Lw = (ctypes.c_int64 * n)()
    assert lib.hipgp_plan_embedding(h, Ln, Lw) == 0
    for d, m in enumerate(dims):
        if m == 1:
            assert (Ln[d], Lw[d]) == (1, 1)
            continue
        for Lv, need in ((Ln[d], 2 * m - 1), (Lw[d], 3 * m - 3)):
            assert Lv >= need and Lv < 2 * need + 8, (Lv, need)
            r = Lv
            for p in (2, 3, 5):
                while r % p == 0:
                    r //= p
            assert r == 1, Lv
    assert lib.hipgp_plan_destroy(h) == 0


def test_new_entry_points_validate_before_touching_the_device():
    """hipgp_toeplitz_quadform / hipgp_block_* reject a null plan and an index that does not tile M' without any CUDA call."""
    L, lib = _lib()
    buf = (ctypes.c_float * 8)()
    idx = (ctypes.c_int64 * 8)(*range(8))
    assert lib.hipgp_toeplitz_quadform(None, buf, buf, 1, 1.0, buf, None) != 0
    assert "null plan" in lib.hipgp_last_error().decode()
    assert lib.hipgp_block_lam(0, buf, buf, idx, 1, 8, 2, 3, 1.0, 1.0, buf, None) != 0          # 2 x 3 != 8
    assert "num_blocks * block_size" in lib.hipgp_last_error().decode()
    assert lib.hipgp_block_diag_multiply(0, buf, buf, None, 1, 8, 2, 4, buf, None) != 0
    assert "null block index" in lib.hipgp_last_error().decode()
    assert lib.hipgp_block_lam(5, buf, buf, idx, 1, 8, 2, 4, 1.0, 1.0, buf, None) != 0 and "dtype" in lib.hipgp_last_error().decode()
    assert lib.hipgp_block_diag_multiply(0, buf, buf, idx, 0, 8, 2, 4, buf, None) == 0          # empty batch: no-op


def test_numa_binding_helper_is_harmless_without_a_gpu():
    """hostmem.bind_to_gpu_numa_node is an optimisation: it parses sysfs cpu lists and must never raise or change the
    affinity when the device's node is unknown (no GPU here)."""
    import os
    from hipgp_b200 import hostmem
    assert hostmem._parse_cpulist("0-3,8,10-11") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    info = hostmem.bind_to_gpu_numa_node(0)
    assert info["bound"] is False
    assert os.sched_getaffinity(0) == before
