"""Slab-decomposed structured K_uu for grids sharded over several GPUs (BASELINE config 5).

Axis 0 of a 3-D grid is split contiguously over the ranks.  A matvec is three local stages with two all-to-all
transposes between them (NCCL via torch.distributed); the FFT passes next to each transpose write / read the exchange
buffer in its packed layout, so no pack / unpack kernels run.  PCG on slab vectors is the reference's conj_grad2 with the
dot products summed over ranks (one small all-reduce per dot).

`SlabToeplitz.matvec_K/_Cinv(v_slab)`: v_slab is this rank's (n0/P, m1, m2) block, flattened or not.
`emulate=True` runs all ranks inside one process on one GPU (lists of slabs in, lists out) -- used by the tests.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from .plan import Plan, _stream_ptr
from .cg import conj_grad2


class SlabToeplitz:
    def __init__(self, dims, column, dtype, device, rank=None, nranks=None, emulate_ranks=None):
        assert len(dims) == 3, "slab decomposition is implemented for 3-D grids"
        self.dims = tuple(int(d) for d in dims)
        self.emulated = emulate_ranks is not None
        if self.emulated:
            self.nranks = int(emulate_ranks); ranks = range(self.nranks)
        else:
            self.nranks = nranks if nranks is not None else (dist.get_world_size() if dist.is_initialized() else 1)
            ranks = [rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)]
        self.plans = []
        for r in ranks:
            p = Plan(self.dims, dtype, device)
            L.check(p.lib, p.lib.hipgp_plan_set_slab(p._h, r, self.nranks))
            p.set_first_row(column)
            self.plans.append(p)
        p0 = self.plans[0]
        a, b = C.c_int64(), C.c_int64()
        L.check(p0.lib, p0.lib.hipgp_slab_sizes(p0._h, C.byref(a), C.byref(b)))
        self.slab_elems, self.exch_elems = a.value, b.value
        self.dtype, self.device = dtype, p0.device
        self.cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
        self.slab_shape = (self.dims[0] // self.nranks, self.dims[1], self.dims[2])

    # ---- stages -------------------------------------------------------------------------------------
    def _s1(self, p, v):
        send = torch.empty(self.exch_elems, dtype=self.cdtype, device=self.device)
        v = v.reshape(-1).to(self.dtype).contiguous()
        L.check(p.lib, p.lib.hipgp_slab_stage1(p._h, C.c_void_p(v.data_ptr()), C.c_void_p(send.data_ptr()), _stream_ptr(self.device)))
        return send

    def _s2(self, p, mode, buf):
        L.check(p.lib, p.lib.hipgp_slab_stage2(p._h, mode, C.c_void_p(buf.data_ptr()), _stream_ptr(self.device)))
        return buf

    def _s3(self, p, buf):
        out = torch.empty(self.slab_elems, dtype=self.dtype, device=self.device)
        L.check(p.lib, p.lib.hipgp_slab_stage3(p._h, C.c_void_p(buf.data_ptr()), C.c_void_p(out.data_ptr()), _stream_ptr(self.device)))
        return out

    def _exchange(self, bufs):
        """all-to-all of equal blocks; `bufs` is a list with one buffer per local (or emulated) rank"""
        if self.emulated:
            blk = self.exch_elems // self.nranks
            return [torch.cat([bufs[q][r * blk:(r + 1) * blk] for q in range(self.nranks)]) for r in range(self.nranks)]
        if self.nranks == 1:
            return bufs
        recv = torch.empty_like(bufs[0])
        dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(bufs[0]))
        return [recv]

    def matvec(self, mode, v):
        """v: this rank's slab (emulated: list of slabs, one per rank); returns the result slab(s), flattened."""
        vs = v if self.emulated else [v]
        with torch.cuda.device(self.device):
            bufs = self._exchange([self._s1(p, x) for p, x in zip(self.plans, vs)])
            bufs = self._exchange([self._s2(p, mode, b) for p, b in zip(self.plans, bufs)])
            outs = [self._s3(p, b) for p, b in zip(self.plans, bufs)]
        return outs if self.emulated else outs[0]

    def matvec_K(self, v):
        return self.matvec(L.MV_K, v)

    def matvec_Cinv(self, v):
        return self.matvec(L.MV_CINV, v)

    # ---- distributed PCG on slab vectors -----------------------------------------------------------
    def solve(self, b_slab, do_precond=True, maxiter=20, tol=1e-8, callback=None):
        """K^-1 b with b sharded like the grid (one right-hand side); the stopping rule and the iterates are those of
        ziggy/misc/cg.py:44-80, the dot products are all-reduced."""
        assert not self.emulated, "solve() runs one rank per process"
        A = lambda v: self.matvec_K(v.reshape(-1)).reshape(1, -1)
        Pm = (lambda v: self.matvec_Cinv(v.reshape(-1)).reshape(1, -1)) if do_precond else None
        red = (lambda t: dist.all_reduce(t)) if self.nranks > 1 else None
        if red is None:
            red = lambda t: t
        return conj_grad2(A, b_slab.reshape(1, -1).to(self.dtype), precond=Pm, maxiter=maxiter, tol=tol, callback=callback,
                          reduce=red)
