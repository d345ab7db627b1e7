"""Error behaviour and degenerate shapes of the C ABI on the GPU: a call with bad arguments returns a negative status
and a message and leaves the plan usable; empty batches are no-ops; unit grid axes are dropped exactly as the
reference's ToeplitzTensor does (toeplitz_tensor.py:17-40: only axes of extent > 1 are embedded)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _col(m, ell=0.2):
    x = np.linspace(0, 1, m)
    c = np.exp(-0.5 * (x - x[0]) ** 2 / ell ** 2)
    c[0] += 1e-2
    return c


def test_calls_before_the_spectrum_exists_fail_cleanly():
    from hipgp_b200.plan import Plan
    from hipgp_b200 import _lib as L
    plan = Plan([40, 30], torch.float64, DEV)
    v = torch.zeros(2, 1200, device=DEV, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="set_first_row"):
        plan.matvec(L.MV_K, v)
    with pytest.raises(RuntimeError, match="set_first_row"):
        plan.pcg(v)
    col = np.outer(_col(40), _col(30)).reshape(-1)
    plan.set_first_row(torch.from_numpy(col).to(DEV))
    out = plan.matvec(L.MV_K, v + 1.0)                     # the failed calls left the plan usable
    assert torch.isfinite(out).all()


def test_bad_arguments_return_a_status_not_an_abort():
    from hipgp_b200.plan import Plan
    from hipgp_b200 import _lib as L
    lib = L.load()
    plan = Plan([64], torch.float32, DEV).set_first_row(torch.from_numpy(_col(64)).to(DEV, torch.float32))
    v = torch.ones(3, 64, device=DEV, dtype=torch.float32)
    o = torch.full_like(v, 7.0)
    h = plan.handle if hasattr(plan, "handle") else plan._h
    assert lib.hipgp_matvec(h, 9, v.data_ptr(), o.data_ptr(), 3, None) != 0 and b"mode" in lib.hipgp_last_error()
    assert lib.hipgp_matvec(h, L.MV_K, v.data_ptr(), o.data_ptr(), -1, None) != 0
    assert lib.hipgp_matvec(h, L.MV_K, None, o.data_ptr(), 3, None) != 0 and b"null" in lib.hipgp_last_error()
    assert lib.hipgp_matvec(None, L.MV_K, v.data_ptr(), o.data_ptr(), 3, None) != 0
    assert lib.hipgp_matvec(h, L.MV_K, v.data_ptr(), o.data_ptr(), 0, None) == 0      # empty batch: no-op
    torch.cuda.synchronize()
    assert (o == 7.0).all()
    it = ctypes.c_int(-5)
    assert lib.hipgp_pcg(h, v.data_ptr(), o.data_ptr(), 0, 5, 1e-8, 1, ctypes.byref(it), None, None, L.ITER_CB(), None, None) == 0
    assert it.value == 0 and (o == 7.0).all()              # empty solve: no-op, zero iterations
    assert lib.hipgp_pcg(h, v.data_ptr(), o.data_ptr(), -2, 5, 1e-8, 1, None, None, None, L.ITER_CB(), None, None) != 0
    assert lib.hipgp_pcg_begin(h, v.data_ptr(), o.data_ptr(), 0, 1e-8, 1, None) != 0
    assert lib.hipgp_plan_spectrum(h, 11, o.data_ptr(), None) != 0
    assert lib.hipgp_plan_set_slab(h, 0, 2) != 0 and b"3-D" in lib.hipgp_last_error()
    assert lib.hipgp_matvec(h, L.MV_K, v.data_ptr(), o.data_ptr(), 3, None) == 0      # still usable
    torch.cuda.synchronize()
    assert torch.isfinite(o).all() and not (o == 7.0).any()


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 1e-6)])
def test_unit_axes_are_dropped(dtype, tol):
    """[1, 50, 1] and [17, 1, 40] behave as [50] and [17, 40]."""
    from hipgp_b200.plan import Plan
    from hipgp_b200 import _lib as L
    rng = np.random.default_rng(3)
    for full, act in (([1, 50, 1], [50]), ([17, 1, 40], [17, 40])):
        col = _col(act[0]) if len(act) == 1 else np.outer(_col(act[0]), _col(act[1])).reshape(-1)
        c = torch.from_numpy(col).to(DEV, dtype)
        pa = Plan(full, dtype, DEV).set_first_row(c)
        pb = Plan(act, dtype, DEV).set_first_row(c)
        assert pa.M == pb.M and pa.Mprime == pb.Mprime
        v = torch.from_numpy(rng.standard_normal((2, pa.M))).to(DEV, dtype)
        for mode in (L.MV_K, L.MV_CINV, L.MV_RT):
            a = pa.matvec(mode, v); b = pb.matvec(mode, v)
            assert torch.equal(a, b)
        xa = pa.pcg(v, maxiter=15, tol=1e-9); xb = pb.pcg(v, maxiter=15, tol=1e-9)
        assert torch.equal(xa, xb)


def test_single_point_and_two_point_grids():
    """m = 2 is the smallest grid the reference's embedding (N = 2m - 2 = 2) admits; K is the 2x2 Toeplitz matrix."""
    from hipgp_b200.plan import Plan
    from hipgp_b200 import _lib as L
    col = torch.tensor([2.0, 0.5], device=DEV, dtype=torch.float64)
    plan = Plan([2], torch.float64, DEV).set_first_row(col)
    v = torch.tensor([[1.0, 0.0], [0.0, 1.0], [3.0, -2.0]], device=DEV, dtype=torch.float64)
    K = torch.tensor([[2.0, 0.5], [0.5, 2.0]], device=DEV, dtype=torch.float64)
    assert torch.allclose(plan.matvec(L.MV_K, v), v @ K, rtol=0, atol=1e-13)
    x = plan.pcg(v, maxiter=5, tol=1e-12)
    assert torch.allclose(x @ K, v, rtol=0, atol=1e-10)
    kn = plan.matvec(L.MV_RT, v)
    assert kn.shape == (3, 2) and torch.allclose((kn * kn).sum(1), (v * (v @ K)).sum(1), atol=1e-12)
