"""Per-kernel-class event timings (rows_fwd / cols / rows_inv) for a K matvec at cfg2 -- dev tool."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hipgp_b200.plan import Plan
from hipgp_b200 import _lib as L, kernels as hk
dev = torch.device("cuda:0")
for dtype in (torch.float32, torch.float64):
    m = 1000
    g1 = torch.linspace(0, 4, m, dtype=dtype, device=dev); g2 = torch.linspace(-2, 2, m, dtype=dtype, device=dev)
    plan = Plan([m, m], dtype, dev).set_first_row(hk.first_row([g1, g2], hk.Matern(nu=2.5, dtype=dtype), (1.0, 0.01), jitter=1e-3))
    for B in (1, 16):
        v = torch.randn(B, m * m, dtype=dtype, device=dev)
        for _ in range(3): plan.matvec(L.MV_K, v)
        plan.profile(True); plan.profile_read(True)
        for _ in range(10): plan.matvec(L.MV_K, v)
        pr = plan.profile_read(True); plan.profile(False)
        tot = sum(v_[0] for v_ in pr.values()) / 10
        print(str(dtype)[6:], "B=%d" % B, " ".join("%s=%.1fus" % (k, 1e3 * a / max(n, 1)) for k, (a, n) in pr.items() if n), "sum=%.1fus" % (1e3 * tot))
