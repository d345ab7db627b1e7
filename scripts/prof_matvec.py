"""Small driver for ncu: a few K matvecs + one PCG iteration pair at cfg2 (fp32 by default)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hipgp_b200.plan import Plan
from hipgp_b200 import _lib as L, kernels as hk
dtype = torch.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else torch.float32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
m = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
dev = torch.device("cuda:0")
g1 = torch.linspace(0, 4, m, dtype=dtype, device=dev); g2 = torch.linspace(-2, 2, m, dtype=dtype, device=dev)
plan = Plan([m, m], dtype, dev).set_first_row(hk.first_row([g1, g2], hk.Matern(nu=2.5, dtype=dtype), (1.0, 0.01 * 1000 / m), jitter=1e-3))
torch.manual_seed(42)
v = torch.randn(B, m * m, dtype=dtype, device=dev)
for _ in range(3):
    plan.matvec(L.MV_K, v)
x = plan.pcg(v, maxiter=2, tol=1e-8)
torch.cuda.synchronize()
print("ok", float(x.abs().sum()))
