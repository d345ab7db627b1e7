"""Drop-in for `ziggy/misc/cg.py` : `conj_grad` (column layout, b is M x L) and `conj_grad2` (row layout, b is bsz x M).

When `A_mul` / `precond` are the matvec methods of one of our structured operators the whole solve runs as the fused,
device-resident PCG of libhipgp_b200.so (hipgp_pcg).  Arbitrary closures still work: the matvecs are whatever the
closure does, and the vector updates / dot products use the library's fused vector kernels (hipgp_vec_*).
Iteration semantics follow the reference exactly (cg.py:58-78): stop test after the x/r update, callback after it.
"""
import ctypes as C

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.float64: L.F64}


def _structured_owner(A_mul, precond):
    """Returns (plan, use_precond) if A_mul is `_matmul_by_K` of a structured operator and precond is None or the
    same operator's `_matmul_by_Cinv`; else None."""
    owner = getattr(A_mul, "__self__", None)
    if owner is None or getattr(A_mul, "__name__", "") != "_matmul_by_K" or not hasattr(owner, "_plan"):
        return None
    if precond is None:
        return owner._plan, False
    if getattr(precond, "__self__", None) is owner and getattr(precond, "__name__", "") == "_matmul_by_Cinv":
        return owner._plan, True
    return None


def _generic_cg(A_mul, b, precond, maxiter, tol, callback, to_rows, from_rows, reduce=None):
    """b in the caller's layout; internally vectors are (bsz, M) contiguous rows."""
    if not b.is_cuda:
        raise RuntimeError("hipgp_b200.cg: tensors must be CUDA tensors (no CPU fallback)")
    lib = L.load()
    dt = _DT[b.dtype]
    dev = b.device
    st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p_ = lambda t: C.c_void_p(t.data_ptr())
    rows = lambda t: to_rows(t).contiguous()
    apply_A = lambda v: rows(A_mul(from_rows(v)))
    apply_P = (lambda v: rows(precond(from_rows(v)))) if precond is not None else (lambda v: v)
    x = torch.zeros_like(rows(b))
    B, M = x.shape
    r = rows(b) - apply_A(x)
    z = apply_P(r)
    p = z.clone()
    rs = torch.empty(B, dtype=torch.float64, device=dev); pAp = torch.empty_like(rs)
    rr = torch.empty_like(rs); zr = torch.empty_like(rs)
    red = reduce if reduce is not None else (lambda t: t)      # sums the per-rank partial dot products (grid-sharded vectors)
    with torch.cuda.device(dev):
        L.check(lib, lib.hipgp_vec_dot(dt, p_(r), p_(z), p_(rs), B, M, st()))
        red(rs)
        for n in range(maxiter):
            Ap = apply_A(p)
            L.check(lib, lib.hipgp_vec_dot(dt, p_(p), p_(Ap), p_(pAp), B, M, st()))
            red(pAp)
            L.check(lib, lib.hipgp_vec_xr_update(dt, p_(x), p_(r), p_(p), p_(Ap), p_(rs), p_(pAp), p_(rr), B, M, st()))
            red(rr)
            if bool(torch.all(torch.sqrt(rr) < tol)):
                break
            z = apply_P(r)
            L.check(lib, lib.hipgp_vec_dot(dt, p_(z), p_(r), p_(zr), B, M, st()))
            red(zr)
            L.check(lib, lib.hipgp_vec_p_update(dt, p_(p), p_(z), p_(zr), p_(rs), B, M, st()))
            rs, zr = zr, rs
            if callback is not None:
                callback(n, from_rows(x.clone()))
    return from_rows(x)


def conj_grad(A_mul, b, precond=None, maxiter=20, tol=1e-10, callback=None):
    """A^{-1} b by (P)CG; A is M x M, b is M x L (cg.py:5-41)."""
    s = _structured_owner(A_mul, precond)
    if s is not None:
        plan, use_p = s
        cb = (lambda n, x: callback(n, x.t())) if callback is not None else None
        return plan.pcg(b.t(), maxiter=maxiter, tol=tol, precond=use_p, callback=cb).t()
    return _generic_cg(A_mul, b, precond, maxiter, tol, callback, lambda t: t.t(), lambda t: t.t())


def conj_grad2(A_mul, b, precond=None, maxiter=20, tol=1e-10, callback=None, reduce=None):
    """A^{-1} b by (P)CG; b is (bsz, M) (cg.py:44-80).  `reduce` (not in the reference) sums device scalars over ranks when
    the vectors are sharded along the grid (slab decomposition)."""
    s = _structured_owner(A_mul, precond) if reduce is None else None
    if s is not None:
        plan, use_p = s
        return plan.pcg(b, maxiter=maxiter, tol=tol, precond=use_p, callback=callback)
    return _generic_cg(A_mul, b, precond, maxiter, tol, callback, lambda t: t, lambda t: t, reduce=reduce)
