"""Developer timing: steady-state cost of hipgp_plan_set_first_row (spectrum set-up) at cfg2 / cfg3 / cfg4."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hipgp_b200.plan import Plan
from hipgp_b200 import kernels as hk
dev = torch.device("cuda:0")
for dims in ((1000, 1000), (300, 300), (128, 128, 64)):
    for dtype in (torch.float32, torch.float64):
        xg = [torch.linspace(0, 1, m, dtype=dtype, device=dev) for m in dims]
        col = hk.first_row(xg, hk.Matern(nu=2.5, dtype=dtype), (1.0, 2.5 / dims[0]), jitter=1e-3)
        plan = Plan(list(dims), dtype, dev)
        for _ in range(2): plan.set_first_row(col)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): plan.set_first_row(col)
        torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 5
        t1 = time.perf_counter()
        for _ in range(5): hk.first_row(xg, hk.Matern(nu=2.5, dtype=dtype), (1.0, 2.5 / dims[0]), jitter=1e-3)
        torch.cuda.synchronize(); tr = (time.perf_counter() - t1) / 5
        print(dims, str(dtype)[6:], "set_first_row %.2f ms   first_row %.3f ms" % (1e3 * t, 1e3 * tr), flush=True)
