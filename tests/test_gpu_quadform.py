"""GPU parity of the Toeplitz-column quadratic form and InvMatmul.backward (learn_kernel=True path):
  * hipgp_toeplitz_quadform against the reference's sym_toeplitz_derivative_quadratic_form on golden vectors
    (gpt_toeplitz.py:169-209, flattened 1-/2-/3-D grid vectors);
  * the column and right-hand-side gradients of InvMatmul.backward against autograd through the reference's own
    ToeplitzTensor.inv_matmul (_inv_matmul.py:28-64);
  * medium / full BASELINE sizes (specialised fp32 lane layout, every embedding family) against a numpy fp64 evaluation
    of the flattened linear correlation (host-side checker only).
Tolerances: 1e-5 relative in fp32, 1e-10 in fp64."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DT = {"f32": torch.float32, "f64": torch.float64}
TOL = {"f32": 1e-5, "f64": 1e-10}


def relerr(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    a = a.astype(np.float64); b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def flat_quadform(u, v):
    """sum_j (c_j[i] + c_j[-i]) for i >= 1, sum_j u_j.v_j at i = 0, c_j the flattened linear cross-correlation (fp64)."""
    S, M = u.shape
    n = 1
    while n < 2 * M:
        n *= 2
    U = np.fft.rfft(u, n, axis=1); V = np.fft.rfft(v, n, axis=1)
    c = np.fft.irfft((U * np.conj(V)).sum(0), n)
    out = c[:M].copy()
    out[1:] += c[:-M:-1][:M - 1]
    return out


def unit_plan(dims, dtype):
    from hipgp_b200.plan import Plan
    M = int(np.prod(dims))
    col = torch.zeros(M, device=DEV, dtype=dtype); col[0] = 1.0     # the form does not depend on the Toeplitz column
    return Plan(list(dims), dtype, DEV).set_first_row(col)


@pytest.mark.parametrize("dname", ["f32", "f64"])
def test_quadform_vs_reference_golden(dname, golden_dir):
    g = np.load(os.path.join(golden_dir, "quadform_%s.npz" % dname))
    for tag in ("1d", "2d", "3d", "2d_odd"):
        dims = [int(x) for x in g[tag + "_dims"]]
        plan = unit_plan(dims, DT[dname])
        u = torch.from_numpy(g[tag + "_u"]).to(DEV); v = torch.from_numpy(g[tag + "_v"]).to(DEV)
        assert relerr(plan.toeplitz_quadform(u, v), g[tag + "_quad"]) < TOL[dname], tag
        assert relerr(plan.toeplitz_quadform(v, u, scale=-0.5), -0.5 * g[tag + "_quad"]) < TOL[dname], tag   # symmetric, scaled


@pytest.mark.parametrize("dname", ["f32", "f64"])
def test_inv_matmul_backward_vs_reference_autograd(dname, golden_dir):
    from hipgp_b200.toeplitz_tensor import ToeplitzTensor
    from hipgp_b200 import kernels as hk
    g = np.load(os.path.join(golden_dir, "quadform_%s.npz" % dname))
    t = np.load(os.path.join(golden_dir, "toeplitz_2d_17x40_matern32_%s.npz" % dname), allow_pickle=True)
    dtype = DT[dname]
    xgrids = [torch.linspace(lo, hi, int(m), dtype=dtype, device=DEV) for lo, hi, m in t["grids"]]
    kern = hk.Matern(nu=1.5, dtype=dtype)
    tt = ToeplitzTensor(xgrids, lambda x, y: kern.forward(x, y, params=(float(t["sig2"]), float(t["ell"]))),
                        batch_shape=None, jitter_val=float(t["jitter"]))
    tt.column.requires_grad_(True)
    Rt = torch.from_numpy(g["bw_R"]).to(DEV).requires_grad_(True)
    wts = torch.from_numpy(g["bw_w"]).to(DEV)
    sol = tt.inv_matmul(Rt, do_precond=True, maxiter=int(g["bw_maxiter"]), tol=float(g["bw_tol"]))
    (sol * wts).sum().backward()
    lim = 1e-9 if dname == "f64" else 2e-3                     # solves: the reference's fp32 PCG noise, as for `_solve`
    assert relerr(sol, g["bw_solves"]) < lim
    assert relerr(Rt.grad, g["bw_right_grad"]) < lim
    assert tt.column.grad.shape == tt.column.shape
    assert relerr(tt.column.grad, g["bw_column_grad"]) < lim
    # the quadratic form itself, on the REFERENCE's solves, to the tight tolerance
    Ls = torch.from_numpy(g["bw_right_grad"]).to(DEV); Rs = torch.from_numpy(g["bw_solves"]).to(DEV)
    assert relerr(tt._plan.toeplitz_quadform(Ls, Rs, scale=-1.0), g["bw_column_grad"]) < 20 * TOL[dname]


@pytest.mark.parametrize("dims,S", [((1000,), 3), ((300, 300), 5), ((96, 200), 2), ((48, 40, 36), 3), ((128, 128, 64), 2),
                                    ((20, 257, 33), 1), ((1000, 1000), 18)],
                         ids=lambda c: "x".join(map(str, c)) if isinstance(c, tuple) else str(c))
def test_quadform_medium_and_full_sizes(dims, S):
    M = int(np.prod(dims))
    rng = np.random.default_rng(2)
    u = rng.standard_normal((S, M)); v = rng.standard_normal((S, M))
    want = flat_quadform(u, v)
    for dname in ("f64", "f32"):
        plan = unit_plan(dims, DT[dname])
        got = plan.toeplitz_quadform(torch.from_numpy(u).to(DEV, DT[dname]), torch.from_numpy(v).to(DEV, DT[dname]))
        assert relerr(got, want) < TOL[dname], (dname, relerr(got, want))
        del plan


def test_quadform_random_small_shapes():
    """carry-pattern gather on many small grids (extents down to 2, odd / prime extents, embeddings with L = 2m - 1 where a
    skipped out-of-range lag would alias onto a valid one), fp64 against the flattened correlation."""
    rng = np.random.default_rng(7)
    shapes = [(2,), (3,), (2, 2), (2, 3), (3, 2), (2, 2, 2), (5, 2, 3), (2, 7, 2), (3, 3, 3), (13, 2), (2, 13)]
    for _ in range(14):
        D = int(rng.integers(1, 4))
        shapes.append(tuple(int(rng.integers(2, 24)) for _ in range(D)))
    for dims in shapes:
        M = int(np.prod(dims))
        S = int(rng.integers(1, 4))
        u = rng.standard_normal((S, M)); v = rng.standard_normal((S, M))
        plan = unit_plan(dims, torch.float64)
        got = plan.toeplitz_quadform(torch.from_numpy(u).to(DEV), torch.from_numpy(v).to(DEV))
        assert relerr(got, flat_quadform(u, v)) < 1e-11, dims
