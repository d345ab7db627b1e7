"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

What shards: right-hand sides / observations.  Every observation's k_n = R^T K_uu^-1 K_un is an
independent solve and every GPU holds a replica of the (small) spectrum, so a minibatch is split contiguously across
ranks, preserving the reference's shuffle=False order inside a step.  The only data-path collective is ONE all-reduce per
step of the packed natural-gradient statistics [data_dm (M'); lam_sum (M'); sum batch_an (1)] (hipgp.py:241-250), plus -
when exact iteration-count parity with the single-GPU stopping rule is requested - one all-reduce(MAX) of a scalar per
PCG iteration (cg.py:70 couples all right-hand sides of the minibatch).
"""
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world():
    return (dist.get_rank(), dist.get_world_size()) if is_dist() else (0, 1)


def shard_slice(n, rank=None, nranks=None):
    """Contiguous, order-preserving split of range(n): the first n % nranks ranks get one extra element."""
    if rank is None or nranks is None:
        rank, nranks = world()
    base, rem = divmod(int(n), int(nranks))
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def allreduce_packed(tensors, group=None):
    """Sum-all-reduce several tensors with ONE collective (they are packed into a flat buffer and unpacked in place)."""
    if not is_dist() or dist.get_world_size(group) == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors


def global_converged(local_max_resid, tol, device, group=None):
    """The reference's stopping test over a minibatch that is sharded across ranks:
    all_b sqrt(r_b.r_b) < tol  <=>  max over ranks of the local max < tol.  NaN never converges (cg.py:70)."""
    v = float("inf") if local_max_resid != local_max_resid else float(local_max_resid)
    t = torch.tensor([v], dtype=torch.float64, device=device)
    if is_dist() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item() < tol)


def sharded_pcg(plan, b_local, maxiter=20, tol=1e-8, precond=True, exact_stop=True, group=None):
    """K^-1 b for this rank's shard of a minibatch with the reference's GLOBAL stopping rule.
    exact_stop=True : one scalar all-reduce(MAX) per iteration, identical iteration count to the unsharded solve.
    exact_stop=False: run `maxiter` iterations (what the SVI defaults 20 / 1e-8 do in practice; no collective)."""
    x = plan.pcg_begin(b_local, tol=-1.0, precond=precond)
    if not exact_stop:
        plan.pcg_step(maxiter, poll=False)
        return x, maxiter
    it = 0
    for it in range(1, maxiter + 1):
        _, _, mx = plan.pcg_step(1, poll=True)
        if global_converged(mx, tol, b_local.device, group):
            break
    return x, it
