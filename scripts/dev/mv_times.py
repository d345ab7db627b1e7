"""Developer timing: K matvec at cfg2 (fp32/fp64, B=1/16) with CUDA events; per-kernel-class split."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hipgp_b200.plan import Plan
from hipgp_b200 import _lib as L, kernels as hk
if os.environ.get('HIPGP_DEV_LIB'): L.LIB_PATH = os.environ['HIPGP_DEV_LIB']
dev = torch.device("cuda:0")
dts = [torch.float32, torch.float64] if "f64" in sys.argv else [torch.float32]
for dtype in dts:
    m = 1000
    g1 = torch.linspace(0, 4, m, dtype=dtype, device=dev); g2 = torch.linspace(-2, 2, m, dtype=dtype, device=dev)
    plan = Plan([m, m], dtype, dev).set_first_row(hk.first_row([g1, g2], hk.Matern(nu=2.5, dtype=dtype), (1.0, 0.01), jitter=1e-3))
    for B in (1, 16):
        v = torch.randn(B, m * m, dtype=dtype, device=dev)
        for _ in range(3): plan.matvec(L.MV_K, v)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): plan.matvec(L.MV_K, v)
        e1.record(); torch.cuda.synchronize()
        tot_ms = e0.elapsed_time(e1) / 20
        plan.profile(True); plan.profile_read(True)
        for _ in range(10): plan.matvec(L.MV_K, v)
        pr = plan.profile_read(True); plan.profile(False)
        print(str(dtype)[6:], "B=%d" % B, "matvec=%.1fus" % (1e3 * tot_ms), " ".join("%s=%.1fus" % (k, 1e3 * a / max(n, 1)) for k, (a, n) in pr.items() if n), flush=True)
    if "rt" in sys.argv:
        v = torch.randn(16, m * m, dtype=dtype, device=dev)
        for mode, nm in ((L.MV_RT, "RT"), (L.MV_R, "R")):
            w = v if mode == L.MV_RT else plan.matvec(L.MV_RT, v)
            for _ in range(2): plan.matvec(mode, w)
            plan.profile(True); plan.profile_read(True)
            for _ in range(5): plan.matvec(mode, w)
            pr = plan.profile_read(True); plan.profile(False)
            print(str(dtype)[6:], nm, "B=16", " ".join("%s=%.1fus" % (k, 1e3 * a / max(n, 1)) for k, (a, n) in pr.items() if n), "emb", plan.embedding(), flush=True)
    if "pcg" in sys.argv:
        v = torch.randn(16, m * m, dtype=dtype, device=dev)
        for _ in range(2): plan.pcg(v, maxiter=20, tol=1e-8)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3): plan.pcg(v, maxiter=20, tol=1e-8)
        e1.record(); torch.cuda.synchronize()
        plan.profile(True); plan.profile_read(True)
        plan.pcg(v, maxiter=20, tol=1e-8)
        pr = plan.profile_read(True); plan.profile(False)
        print(str(dtype)[6:], "PCG(20) B=16: %.2f ms" % (e0.elapsed_time(e1) / 3), " ".join("%s=%.1fus" % (k, 1e3 * a / max(n, 1)) for k, (a, n) in pr.items() if n), flush=True)
