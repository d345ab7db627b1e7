"""dev: where the BlockToeplitzGP step spends its time (300x300 grid, batch 200, 13x13 blocks)"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hipgp_b200 import kernels as hk
from hipgp_b200.hipgp import BlockToeplitzGP
from hipgp_b200.plan import meanfield_rowstats, meanfield_colstats
DEV = "cuda:0"; dt = torch.float32; B = 200

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

xgrids = [torch.linspace(0, 1, 300, dtype=dt), torch.linspace(0, 1, 300, dtype=dt)]
mod = BlockToeplitzGP(hk.Matern(nu=1.5, dtype=dt), xgrids, num_obs=10 ** 6, block_sizes=[13, 13], ell_init=0.02, dtype=dt).cuda_params(0)
x = torch.rand(B, 2, device=DEV, dtype=dt)
with torch.no_grad():
    Knm, Kd = mod._make_grams(x)
    kn = mod.compute_kn(Knm, maxiter_cg=20)
    qm, qS = mod.standard_variational_params()
    w = torch.rand(B, device=DEV, dtype=dt)
    print("grams+compute_kn", timeit(lambda: mod.compute_kn(mod._make_grams(x)[0], maxiter_cg=20)))
    print("standard_variational_params", timeit(mod.standard_variational_params))
    print("  torch.inverse", timeit(lambda: torch.inverse(-2 * mod.global_theta2.data)))
    print("row_stats", timeit(lambda: mod._row_stats(kn, qm, qS)))
    print("batch_stats", timeit(lambda: mod._batch_stats(kn, w, w)))
    lam = mod._batch_stats(kn, w, w)[1]
    print("natural_gradient", timeit(lambda: mod._natural_gradient(qm, lam, qm, 2.0)))
    print("kl", timeit(lambda: mod.get_kl_to_prior(qm, qS)))
    A = -2 * mod.global_theta2.data
    eye = torch.eye(A.shape[-1], device=DEV, dtype=dt).expand_as(A).contiguous()
    print("  cholesky", timeit(lambda: torch.linalg.cholesky(A)))
    Lc = torch.linalg.cholesky(A)
    print("  cholesky_solve(I)", timeit(lambda: torch.cholesky_solve(eye, Lc)))
    print("  cholesky_inverse", timeit(lambda: torch.cholesky_inverse(Lc)))
    print("  linalg.inv", timeit(lambda: torch.linalg.inv(A)))
    print("  solve_triangular+bmm", timeit(lambda: (lambda Li: Li.transpose(-1, -2) @ Li)(torch.linalg.solve_triangular(Lc, eye, upper=False))))
