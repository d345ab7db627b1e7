"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: order-preserving sharding, the single packed all-reduce of
the natural-gradient statistics, and the global PCG stopping rule.  The arithmetic kernels need a GPU; what is checked
here is that sharded statistics combine to exactly the unsharded ones and that all ranks take the same decision."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hipgp_b200 import dist as hdist


def test_shard_slice_is_contiguous_and_complete():
    for n in (0, 1, 7, 200, 201):
        for w in (1, 2, 3, 8):
            got = []
            for r in range(w):
                s = hdist.shard_slice(n, r, w)
                got += list(range(n))[s]
            assert got == list(range(n))
            sizes = [len(range(n)[hdist.shard_slice(n, r, w)]) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        B, E = 11, 37
        kn = torch.randn(B, E, dtype=torch.float64)          # stands in for k_n; identical on every rank
        w1 = torch.randn(B, dtype=torch.float64); w2 = torch.rand(B, dtype=torch.float64)
        sl = hdist.shard_slice(B)
        dm = (w1[sl, None] * kn[sl]).sum(0); lam = (w2[sl, None] * kn[sl] ** 2).sum(0); an = kn[sl].sum().reshape(1)
        # block family: the per-block statistic sum_n w_n k k^T is a (num_blocks, bs, bs) tensor in the same packed reduction
        from hipgp_b200.util import define_block_chunks
        idx, _, _ = define_block_chunks([torch.arange(6), torch.arange(6)], [3, 2])          # 6 blocks of 6 (E = 36 of the 37 columns)
        kb = kn[sl][:, :36][:, idx].transpose(0, 1)                                               # (num_blocks, rows, bs)
        lam_blk = torch.matmul(kb.transpose(1, 2), w2[sl][None, :, None] * kb)
        hdist.allreduce_packed([dm, lam, an, lam_blk])
        kb_all = kn[:, :36][:, idx].transpose(0, 1)
        ok = torch.allclose(dm, (w1[:, None] * kn).sum(0)) and torch.allclose(lam, (w2[:, None] * kn ** 2).sum(0)) \
            and torch.allclose(an, kn.sum().reshape(1)) \
            and torch.allclose(lam_blk, torch.matmul(kb_all.transpose(1, 2), w2[None, :, None] * kb_all))
        # global stop: converged only when EVERY rank is below tol; NaN on one rank never converges
        c1 = hdist.global_converged(1e-9 if rank == 0 else 1e-3, 1e-8, "cpu")
        c2 = hdist.global_converged(1e-9, 1e-8, "cpu")
        c3 = hdist.global_converged(float("nan") if rank == 1 else 1e-12, 1e-8, "cpu")
        out[rank] = (bool(ok), c1, c2, c3)
    finally:
        dist.destroy_process_group()


def test_packed_allreduce_and_global_stop_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        ok, c1, c2, c3 = out[r]
        assert ok and c1 is False and c2 is True and c3 is False


def test_single_process_is_a_noop():
    t = [torch.ones(3), torch.ones(1)]
    assert hdist.allreduce_packed(t)[0].tolist() == [1.0, 1.0, 1.0]
    assert hdist.world() == (0, 1)
    assert hdist.global_converged(1e-9, 1e-8, "cpu") is True
