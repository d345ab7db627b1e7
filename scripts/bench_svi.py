"""BASELINE config 3/4 shaped SVI benchmark: mean-field natural-gradient minibatch steps with the observations of each
minibatch sharded over the ranks (one packed all-reduce per step).  Run under torchrun for N > 1:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_svi.py [cfg3|cfg4] [steps] [mean-field|block]
Prints one JSON line from rank 0.  cfg3: 300x300 grid, Matern-3/2, point observations (UK-housing shape);
cfg4: 128x128x64 grid, SqExp analytic line integrals from the origin (dust-map shape)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from hipgp_b200 import hipgp as hh, kernels as hk

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
family = sys.argv[3] if len(sys.argv) > 3 else "mean-field"
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
dtype = torch.float32
bsz = 200
rs = np.random.RandomState(42)
if cfg == "cfg3":
    xgrids = [torch.linspace(-5.7, 1.8, 300, dtype=dtype), torch.linspace(50, 55.5, 300, dtype=dtype)]
    kern = hk.Matern(nu=1.5, dtype=dtype); ell = 0.05; sig2 = 1.0; integ = False
    nobs = 2_000_000
    X = np.stack([rs.uniform(-5.7, 1.8, steps * bsz), rs.uniform(50, 55.5, steps * bsz)], 1)
    Y = rs.randn(steps * bsz, 1); NS = np.full((steps * bsz, 1), 0.3)
else:
    xgrids = [torch.linspace(-.25, .25, 128, dtype=dtype), torch.linspace(-.25, .25, 128, dtype=dtype), torch.linspace(-.05, .05, 64, dtype=dtype)]
    kern = hk.SqExp(dtype=dtype); ell = 0.01; sig2 = 0.1; integ = True
    nobs = 5_000_000
    X = np.stack([rs.uniform(-.25, .25, steps * bsz), rs.uniform(-.25, .25, steps * bsz), rs.uniform(-.05, .05, steps * bsz)], 1)
    Y = rs.uniform(5e-4, 3.2e-2, (steps * bsz, 1)); NS = rs.uniform(0.0025, 0.0075, (steps * bsz, 1))
    # the doubly-integrated diagonal table is host quadrature at ctor time in the reference; a fixed table keeps this bench
    # about the hot path
    tab = np.stack([np.linspace(0, 5, 50), np.zeros(50), np.linspace(1, 0.1, 50)])
    kern._diag_interp = hk.KernelDoublyDiagInterpolator(kern, table=tab)
if family == "block":      # neighbouring chunks of the whitened grid: 13x13 of 598x598 (cfg3), 2x2x2 of 254x254x126 (cfg4)
    blocks = [13, 13] if cfg == "cfg3" else [2, 2, 2]
    mod = hh.BlockToeplitzGP(kern, xgrids, num_obs=nobs, block_sizes=blocks, sig2_init=sig2, ell_init=ell, dtype=dtype,
                             jitter_val=1e-3).cuda_params(local)
else:
    mod = hh.MeanFieldToeplitzGP(kern, xgrids, num_obs=nobs, sig2_init=sig2, ell_init=ell, dtype=dtype, jitter_val=1e-3).cuda_params(local)
opt = torch.optim.SGD([mod.global_theta1, mod.global_theta2], lr=1e-4)   # random synthetic targets: keep the iteration tame
Xh = torch.from_numpy(X).to(dtype).pin_memory(); Yh = torch.from_numpy(Y).to(dtype).pin_memory(); Nh = torch.from_numpy(NS).to(dtype).pin_memory()

def step(i):
    sl = slice(i * bsz, (i + 1) * bsz)
    xb = Xh[sl].to(dev, non_blocking=True); yb = Yh[sl].to(dev, non_blocking=True); nb = Nh[sl].to(dev, non_blocking=True)
    elbo = mod.elbo_and_grad(xb, yb, nb, maxiter_cg=20, integrated_obs=integ)
    opt.step()
    return elbo

for i in range(3):
    step(i)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    el = step(i)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = ms.item() / steps / 1e3
    print(json.dumps({"bench": "svi_minibatch_step", "config": cfg, "family": family, "n_gpus": world, "batch_size": bsz, "maxiter_cg": 20, "dtype": "f32",
                      "s_per_step": t, "obs_per_s": bsz / t, "epoch_s_extrapolated": nobs / bsz * t, "elbo_last": float(el),
                      "M": mod.M, "Mprime": mod.Mprime, "embedding": list(mod.make_Kmm()._plan.embedding()[0])}))
if world > 1:
    dist.destroy_process_group()
