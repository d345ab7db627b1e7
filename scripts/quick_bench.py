"""Quick timing of the structured matvecs / PCG at the 10^6-point grid (dev tool; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hipgp_b200.plan import Plan
from hipgp_b200 import _lib as L

def matern52_row(m1, m2, ell, dtype, jitter=1e-3):
    g1 = torch.linspace(0, 4, m1, dtype=torch.float64); g2 = torch.linspace(-2, 2, m2, dtype=torch.float64)
    r = torch.sqrt((g1[:, None] - g1[0]) ** 2 + (g2[None, :] - g2[0]) ** 2)
    dp = np.sqrt(5) * r / ell
    k = (1 + dp + 5. / 3. * r * r / ell ** 2) * torch.exp(-dp)
    k[0, 0] += jitter
    return k.reshape(-1).to(dtype).cuda()

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for dtype in (torch.float32, torch.float64):
    w = 4 if dtype == torch.float32 else 8
    plan = Plan([m, m], dtype, "cuda:0")
    t0 = time.time(); plan.set_first_row(matern52_row(m, m, 0.01 * 1000 / m, dtype)); torch.cuda.synchronize()
    print(dtype, "setup s", round(time.time() - t0, 3), "embedding", plan.embedding(), "clamped", plan.num_clamped)
    for B in (1, 16):
        torch.manual_seed(42)
        v = torch.randn(B, m * m, dtype=dtype, device="cuda")
        for mode, name in ((L.MV_K, "K"), (L.MV_CINV, "Cinv"), (L.MV_RT, "RT")):
            ms = timeit(lambda: plan.matvec(mode, v))
            Ln = plan.embedding()[0]
            alg = w * (2 * m * m * B + Ln[0] * (Ln[1] // 2 + 1))
            print("  B=%d %s: %.3f ms  alg GB/s %.1f" % (B, name, ms, alg / ms / 1e6))
        ms = timeit(lambda: plan.pcg(v, maxiter=20, tol=1e-8), n=5, warm=2)
        x, info = plan.pcg(v, maxiter=20, tol=1e-8, return_info=True)
        print("  B=%d PCG(20): %.3f ms  (%.3f ms/iter) iters %d resid %s" % (B, ms, ms / 20, info["iters"], info["resid"][:2]))
        # cuFFT comparison point (same algebra via torch.fft)
        Lf = plan.embedding()[0]
        S = torch.randn(Lf[0], Lf[1] // 2 + 1, dtype=dtype, device="cuda")
        def cufft():
            F = torch.fft.rfft2(v.view(B, m, m), s=(Lf[0], Lf[1]))
            return torch.fft.irfft2(F * S, s=(Lf[0], Lf[1]))[:, :m, :m]
        print("  B=%d cuFFT-based matvec: %.3f ms" % (B, timeit(cufft)))
    del plan
