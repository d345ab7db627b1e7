"""Single-GPU 512^3 K matvec / PCG(20) (undecomposed plan) -- the denominator of the slab scaling efficiency."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hipgp_b200.plan import Plan
from hipgp_b200 import _lib as L, kernels as hk
dev = torch.device("cuda:0"); dtype = torch.float32
dims = (512, 512, 512)
xg = [torch.linspace(0, 1, m, dtype=dtype, device=dev) for m in dims]
step = float(xg[0][1] - xg[0][0])
plan = Plan(list(dims), dtype, dev).set_first_row(hk.first_row(xg, hk.Matern(nu=2.5, dtype=dtype), (1.0, 2.5 * step), jitter=1e-3))
v = torch.randn(1, 512 ** 3, dtype=dtype, device=dev)
def timed(fn, n):
    for _ in range(2): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timed(lambda: plan.matvec(L.MV_K, v), 10)
pcg = timed(lambda: plan.pcg(v, maxiter=20, tol=1e-8), 2)
print(json.dumps({"bench": "matvec_512cubed_1gpu", "matvec_ms": ms, "pcg20_s": pcg / 1e3, "device_MB": plan.device_bytes() / 1e6 if hasattr(plan, "device_bytes") else None}))
