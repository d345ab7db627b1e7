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


def _slab_worker(rank, world, port, out):
    """grid-sharded matvec across two PROCESSES: the library's three local stages (CPU emulation build of the same kernel
    sources) around torch.distributed all_to_all_single, exactly as hipgp_b200/slab.py orders them (bins layout, two chunks)"""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emu_build
    from hipgp_b200 import _lib as L
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = emu_build.load()
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        dims = (8, 6, 10)
        g = np.meshgrid(*[np.linspace(0, 1 + d, k) for d, k in enumerate(dims)], indexing="ij")
        r = np.sqrt(sum((x - x.flat[0]) ** 2 for x in g))
        col = ((1 + np.sqrt(3) * r / 0.4) * np.exp(-np.sqrt(3) * r / 0.4)).reshape(-1); col[0] += 1e-2

        def mkplan(slab):
            plan = C.c_void_p(); mm = np.array(dims, dtype=np.int64)
            assert lib.hipgp_plan_create(3, mm.ctypes.data_as(L._pi64), L.F64, 0, C.byref(plan)) == 0
            if slab:
                assert lib.hipgp_plan_set_slab(plan, rank, world) == 0
                assert lib.hipgp_plan_set_slab_chunks(plan, 2) == 0
            assert lib.hipgp_plan_set_first_row(plan, ptr(col), 1e-6, None, None) == 0, lib.hipgp_last_error()
            return plan
        plan = mkplan(True)
        a, b = C.c_int64(), C.c_int64()
        assert lib.hipgp_slab2_sizes(plan, C.byref(a), C.byref(b)) == 0
        slab_elems, exch = a.value, b.value
        M = int(np.prod(dims)); n0 = dims[0] // world
        v = np.random.default_rng(0).standard_normal((1, M))                     # same on both ranks
        mine = np.ascontiguousarray(v.reshape(dims)[rank * n0:(rank + 1) * n0]).reshape(-1)

        def exchange(buf):                                                       # per chunk, like SlabToeplitz._exchange
            recv = np.empty_like(buf); per = exch // 2
            for c in range(2):
                s = torch.view_as_real(torch.from_numpy(buf[c * per:(c + 1) * per]))
                t = torch.view_as_real(torch.from_numpy(recv[c * per:(c + 1) * per]))
                dist.all_to_all_single(t, s)
            return recv
        errs = []
        for mode in (L.MV_K, L.MV_CINV):
            send = np.zeros(exch, dtype=np.complex128)
            assert lib.hipgp_slab2_stage_a(plan, ptr(mine), ptr(send), None) == 0, lib.hipgp_last_error()
            buf = exchange(send)
            for c in range(2):
                assert lib.hipgp_slab2_stage_b_chunk(plan, mode, ptr(buf), c, None) == 0, lib.hipgp_last_error()
            back = exchange(buf)
            o = np.zeros(slab_elems)
            assert lib.hipgp_slab2_stage_c(plan, ptr(back), ptr(o), None) == 0, lib.hipgp_last_error()
            parts = [torch.empty(slab_elems, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(o))
            if rank == 0:
                full = mkplan(False)
                ref = np.zeros((1, M))
                assert lib.hipgp_matvec(full, mode, ptr(v), ptr(ref), 1, None) == 0
                got = torch.cat(parts).numpy().reshape(1, -1)
                errs.append(float(np.linalg.norm(got - ref) / np.linalg.norm(ref)))
        out[rank] = errs
    finally:
        dist.destroy_process_group()


def test_slab_matvec_across_two_processes_gloo():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import emu_build
    emu_build.load()                       # build once in the parent, the workers only load it
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29400 + (os.getpid() % 150)
    mp.spawn(_slab_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out[0]) == 2 and all(e < 1e-12 for e in out[0]), out[0]
