"""Where does a config-3 mean-field step spend its time when a rank holds only 25 of the 200 observations (the 8-GPU share)?
GPU-busy time (sum of the pass kernels' CUDA-event durations) against the wall clock of the step (developer tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from hipgp_b200 import hipgp as hh, kernels as hk
dev = torch.device("cuda:0"); dtype = torch.float32
xg = [torch.linspace(-5.7, 1.8, 300, dtype=dtype), torch.linspace(50, 55.5, 300, dtype=dtype)]
mod = hh.MeanFieldToeplitzGP(hk.Matern(nu=1.5, dtype=dtype), xg, num_obs=2_000_000, sig2_init=1.0, ell_init=0.05, dtype=dtype, jitter_val=1e-3).cuda_params(0)
rs = np.random.RandomState(0)
for bsz in (200, 25):
    X = torch.from_numpy(np.stack([rs.uniform(-5.7, 1.8, bsz), rs.uniform(50, 55.5, bsz)], 1)).to(dtype).to(dev)
    Y = torch.randn(bsz, 1, dtype=dtype, device=dev); NS = torch.full((bsz, 1), 0.3, dtype=dtype, device=dev)
    for _ in range(3): mod.elbo_and_grad(X, Y, NS, maxiter_cg=20)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): mod.elbo_and_grad(X, Y, NS, maxiter_cg=20)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 20
    plan = mod.make_Kmm()._plan
    plan.profile(True); plan.profile_read(True)
    for _ in range(5): mod.elbo_and_grad(X, Y, NS, maxiter_cg=20)
    pr = plan.profile_read(True); plan.profile(False)
    busy = sum(v[0] for v in pr.values()) / 5
    nl = sum(v[1] for v in pr.values()) / 5
    # parts
    t1 = time.perf_counter()
    for _ in range(20): Knm, Knn = mod._make_grams(X)
    torch.cuda.synchronize(); t_grams = (time.perf_counter() - t1) / 20
    t1 = time.perf_counter()
    for _ in range(20): kn = mod.compute_kn(Knm, maxiter_cg=20)
    torch.cuda.synchronize(); t_kn = (time.perf_counter() - t1) / 20
    print("bsz %3d: step wall %.3f ms | pass kernels busy %.3f ms in %d launches | make_grams %.3f ms | compute_kn %.3f ms" % (bsz, wall * 1e3, busy, nl, t_grams * 1e3, t_kn * 1e3), flush=True)
