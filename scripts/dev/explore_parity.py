"""Exploratory numbers behind the bounds of tests/test_gpu_bench_parity.py (developer tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from hipgp_b200.plan import Plan
from hipgp_b200 import _lib as L
from oracle import ziggy_oracle as zo
DEV = "cuda:0"
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
m = 1000
torch.manual_seed(42)
v64 = torch.randn(2, m * m, dtype=torch.float64)
res = {}
for dt in (torch.float32, torch.float64):
    g1 = torch.linspace(0, 4, m, dtype=dt); g2 = torch.linspace(-2, 2, m, dtype=dt)
    t0 = time.time()
    ora = zo.OracleToeplitz([g1, g2], lambda x, y: zo.matern(x, y, 1.0, 0.01, 2.5), jitter_val=1e-3)
    v = v64.to(dt)
    ncb = [0]
    xo = ora.solve(v, do_precond=True, maxiter=20, tol=1e-8, callback=lambda n, x: ncb.__setitem__(0, ncb[0] + 1))
    print(dt, "oracle solve s", time.time() - t0, "callbacks", ncb[0], flush=True)
    plan = Plan([m, m], dt, DEV).set_first_row(ora.column.to(DEV))
    nd = [0]
    x, info = plan.pcg(v.to(DEV), maxiter=20, tol=1e-8, callback=lambda n, xx: nd.__setitem__(0, nd[0] + 1), return_info=True)
    res[dt] = (xo, x.cpu(), ora, plan, info)
    print(dt, "dev callbacks", nd[0], "iters", info["iters"], "x dev vs oracle", rel(x.cpu(), xo), "resid dev", info["resid"], flush=True)
    r_or = (v - ora.matmul_K(xo)).norm(dim=1); r_dev = (v - ora.matmul_K(x.cpu())).norm(dim=1)
    print(dt, "true residual oracle", r_or.tolist(), "true residual dev", r_dev.tolist(), flush=True)
    for nm, f, mode in (("K", ora.matmul_K, L.MV_K), ("Cinv", ora.matmul_Cinv, L.MV_CINV), ("RT", ora.matmul_RT, L.MV_RT)):
        res[(dt, nm)] = (f(v), plan.matvec(mode, v.to(DEV)).cpu())
x32o, x32d = res[torch.float32][0], res[torch.float32][1]; x64o, x64d = res[torch.float64][0], res[torch.float64][1]
print("fp32 oracle vs fp64 oracle", rel(x32o, x64o), "fp32 dev vs fp64 oracle", rel(x32d, x64o), "fp64 dev vs fp64 oracle", rel(x64d, x64o))
for nm in ("K", "Cinv", "RT"):
    o32, d32 = res[(torch.float32, nm)]; o64, d64 = res[(torch.float64, nm)]
    print(nm, "f32: dev vs oracle32", rel(d32, o32), "| oracle32 vs truth64", rel(o32, o64), "| dev32 vs truth64", rel(d32, o64), "| f64 dev vs oracle", rel(d64, o64))
