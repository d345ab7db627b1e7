"""Run under torchrun with N >= 2 GPUs (optional argument: block): the sharded natural-gradient step (packed all-reduce) must give every rank the same
gradients as the unsharded step computed locally on rank 0's GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from hipgp_b200 import hipgp as hh, kernels as hk
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
dtype = torch.float64
xg = [torch.linspace(-5.7, 1.8, 60, dtype=dtype), torch.linspace(50, 55.5, 44, dtype=dtype)]
torch.manual_seed(1)
family = sys.argv[1] if len(sys.argv) > 1 else "mean-field"
if family == "block":      # M' = 118 x 86 -> 59 x 43-point blocks
    mod = hh.BlockToeplitzGP(hk.Matern(nu=1.5, dtype=dtype), xg, num_obs=10000, block_sizes=[59, 43], sig2_init=1.0, ell_init=0.4,
                             dtype=dtype).cuda_params(local)
else:
    mod = hh.MeanFieldToeplitzGP(hk.Matern(nu=1.5, dtype=dtype), xg, num_obs=10000, sig2_init=1.0, ell_init=0.4, dtype=dtype).cuda_params(local)
rs = np.random.RandomState(0)
x = torch.tensor(np.stack([rs.uniform(-5.7, 1.8, 37), rs.uniform(50, 55.5, 37)], 1), dtype=dtype, device=dev)
y = torch.tensor(rs.randn(37, 1), dtype=dtype, device=dev); nb = torch.full((37, 1), 0.3, dtype=dtype, device=dev)
e_sh = mod.elbo_and_grad(x, y, nb, maxiter_cg=20, shard=True)
g1s, g2s = mod.global_theta1.grad.clone(), mod.global_theta2.grad.clone()
e_full = mod.elbo_and_grad(x, y, nb, maxiter_cg=20, shard=False)
err = max(float((g1s - mod.global_theta1.grad).abs().max() / mod.global_theta1.grad.abs().max()),
          float((g2s - mod.global_theta2.grad).abs().max() / mod.global_theta2.grad.abs().max()), abs(float(e_sh - e_full)))
t = torch.tensor([err], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("%s family: sharded vs unsharded max rel diff over ranks: %.3e (world %d)" % (family, t.item(), world))
    assert t.item() < 1e-9
dist.destroy_process_group()
