"""Developer timing: 3-D K matvec / PCG on cfg4-shaped (128,128,64) and 256^3 grids; per-kernel-class split."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hipgp_b200.plan import Plan
from hipgp_b200 import _lib as L, kernels as hk
if os.environ.get('HIPGP_DEV_LIB'): L.LIB_PATH = os.environ['HIPGP_DEV_LIB']
dev = torch.device("cuda:0"); dtype = torch.float32
for dims, B in (((128, 128, 64), 200), ((128, 128, 64), 16), ((256, 256, 256), 1), ((300, 300), 200)):
    xg = [torch.linspace(0, 1, m, dtype=dtype, device=dev) for m in dims]
    plan = Plan(list(dims), dtype, dev).set_first_row(hk.first_row(xg, hk.Matern(nu=2.5, dtype=dtype), (1.0, 2.5 / dims[0]), jitter=1e-3))
    M = 1
    for m in dims: M *= m
    v = torch.randn(B, M, dtype=dtype, device=dev)
    for mode, nm in ((L.MV_K, "K"), (L.MV_RT, "RT")):
        for _ in range(2): plan.matvec(mode, v)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5): plan.matvec(mode, v)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        plan.profile(True); plan.profile_read(True)
        for _ in range(3): plan.matvec(mode, v)
        pr = plan.profile_read(True); plan.profile(False)
        print(dims, "B=%d" % B, nm, "%.3f ms" % ms, "emb", plan.embedding(), " ".join("%s=%.1fus(x%d)" % (k, 1e3 * a / max(n, 1), n // 3) for k, (a, n) in pr.items() if n), flush=True)
    del plan, v
    torch.cuda.empty_cache()
