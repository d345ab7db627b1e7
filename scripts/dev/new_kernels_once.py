"""dev: one call each of hipgp_toeplitz_quadform (cfg2, 16 pairs, fp32) and the block-family kernels (cfg3 shape), for an
ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hipgp_b200.plan import Plan, block_lam, block_diag_multiply
from hipgp_b200.util import define_block_chunks
DEV = "cuda:0"; dt = torch.float32
M = 10 ** 6
col = torch.zeros(M, device=DEV, dtype=dt); col[0] = 1
plan = Plan([1000, 1000], dt, DEV).set_first_row(col)
u = torch.randn(16, M, device=DEV, dtype=dt); v = torch.randn(16, M, device=DEV, dtype=dt)
torch.cuda.synchronize()
for _ in range(2):
    out = plan.toeplitz_quadform(u, v)
idx, _, _ = define_block_chunks([torch.arange(598), torch.arange(598)], [13, 13])
idx = idx.to(DEV)
kn = torch.randn(200, 598 * 598, device=DEV, dtype=dt); w = torch.rand(200, device=DEV, dtype=dt)
S = torch.randn(idx.shape[0], 169, 169, device=DEV, dtype=dt)
for _ in range(2):
    lam = block_lam(kn, w, idx, 2.0, 1.0)
    sv = block_diag_multiply(S, kn, idx)
torch.cuda.synchronize()
print("ok", float(out[0]), float(lam[0, 0, 0]), float(sv[0, 0]))
