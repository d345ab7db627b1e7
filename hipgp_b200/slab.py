"""Slab-decomposed structured K_uu for grids sharded over several GPUs (BASELINE config 5).

Axis 0 of a 3-D grid is split contiguously over the ranks.  A matvec is three local stages with two all-to-all
transposes between them (NCCL via torch.distributed).  Two layouts of what travels:

  layout="bins" (default, round 2): the UN-PADDED output of the row pass, split along the bins of the last axis; after the
      exchange a rank owns every (i0, i1) for its bins and runs the three ordinary column passes locally.  Half the bytes on
      the wire of the first version.
  layout="axis1" (round 1): the zero-padded axis-1 transform, split along axis-1 positions; the column passes next to each
      transpose write / read the exchange buffer in its packed layout.

PCG on slab vectors is the reference's conj_grad2 with the dot products summed over ranks (one small all-reduce per dot).

`SlabToeplitz.matvec_K/_Cinv(v_slab)`: v_slab is this rank's (n0/P, m1, m2) block, flattened or not.
`emulate=True` runs all ranks inside one process on one GPU (lists of slabs in, lists out) -- used by the tests.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib as L
from .plan import Plan, _stream_ptr
from .cg import conj_grad2


class SlabToeplitz:
    def __init__(self, dims, column, dtype, device, rank=None, nranks=None, emulate_ranks=None, layout=None):
        assert len(dims) == 3, "slab decomposition is implemented for 3-D grids"
        layout = layout or os.environ.get("HIPGP_SLAB_LAYOUT", "bins")
        assert layout in ("bins", "axis1")
        self.layout = layout
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
        sizes = p0.lib.hipgp_slab2_sizes if layout == "bins" else p0.lib.hipgp_slab_sizes
        L.check(p0.lib, sizes(p0._h, C.byref(a), C.byref(b)))
        self.slab_elems, self.exch_elems = a.value, b.value
        self.dtype, self.device = dtype, p0.device
        self.cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
        self.slab_shape = (self.dims[0] // self.nranks, self.dims[1], self.dims[2])

    # ---- stages -------------------------------------------------------------------------------------
    def _s1(self, p, v):
        send = torch.empty(self.exch_elems, dtype=self.cdtype, device=self.device)
        v = v.reshape(-1).to(self.dtype).contiguous()
        f = p.lib.hipgp_slab2_stage_a if self.layout == "bins" else p.lib.hipgp_slab_stage1
        L.check(p.lib, f(p._h, C.c_void_p(v.data_ptr()), C.c_void_p(send.data_ptr()), _stream_ptr(self.device)))
        return send

    def _s2(self, p, mode, buf):
        f = p.lib.hipgp_slab2_stage_b if self.layout == "bins" else p.lib.hipgp_slab_stage2
        L.check(p.lib, f(p._h, mode, C.c_void_p(buf.data_ptr()), _stream_ptr(self.device)))
        return buf

    def _s3(self, p, buf):
        out = torch.empty(self.slab_elems, dtype=self.dtype, device=self.device)
        f = p.lib.hipgp_slab2_stage_c if self.layout == "bins" else p.lib.hipgp_slab_stage3
        L.check(p.lib, f(p._h, C.c_void_p(buf.data_ptr()), C.c_void_p(out.data_ptr()), _stream_ptr(self.device)))
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
    def solve(self, b_slab, do_precond=True, maxiter=20, tol=1e-8, callback=None, check_every=5):
        """K^-1 b with b sharded like the grid (one right-hand side); the stopping rule and the iterates are those of
        ziggy/misc/cg.py:44-80 with the dot products summed over ranks.

        The loop never waits for the device inside an iteration: the reference's `break` is a device-side flag that turns every
        later update into a no-op (so x is the reference's iterate at the iteration it would have stopped), and the host
        looks at the flag every `check_every` iterations only to stop launching.  Two small all-reduces per iteration
        (p.Ap, then r.r and z.r together).  `self.last_iters` holds the iteration count afterwards.
        A `callback` needs the iterate on the host each iteration: that case takes the generic synchronous loop."""
        assert not self.emulated, "solve() runs one rank per process"
        A = lambda v: self.matvec_K(v.reshape(-1)).reshape(1, -1)
        Pm = (lambda v: self.matvec_Cinv(v.reshape(-1)).reshape(1, -1)) if do_precond else None
        multi = self.nranks > 1
        red = (lambda t: dist.all_reduce(t)) if multi else (lambda t: t)
        if callback is not None:
            return conj_grad2(A, b_slab.reshape(1, -1).to(self.dtype), precond=Pm, maxiter=maxiter, tol=tol, callback=callback,
                              reduce=red)
        p0 = self.plans[0]; lib = p0.lib
        dt = L.F32 if self.dtype == torch.float32 else L.F64
        dev = self.device
        ptr = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            st = lambda: _stream_ptr(dev)
            b = b_slab.reshape(1, -1).to(self.dtype).contiguous()
            M = b.shape[1]
            x = torch.zeros_like(b)
            r = b.clone()                                   # r = b - K 0
            z = Pm(r) if Pm is not None else r
            p = z.clone()
            f64 = dict(dtype=torch.float64, device=dev)
            rs = torch.empty(1, **f64); pAp = torch.empty(1, **f64); sc = torch.empty(2, **f64)
            zero = torch.zeros(1, **f64); one = torch.ones(1, **f64)
            active = torch.ones(1, dtype=torch.bool, device=dev)
            iters = torch.zeros(1, dtype=torch.int64, device=dev)
            L.check(lib, lib.hipgp_vec_dot(dt, ptr(r), ptr(z), ptr(rs), 1, M, st()))
            red(rs)
            for n in range(maxiter):
                Ap = A(p)
                L.check(lib, lib.hipgp_vec_dot(dt, ptr(p), ptr(Ap), ptr(pAp), 1, M, st()))
                red(pAp)
                num = torch.where(active, rs, zero); den = torch.where(active, pAp, one)      # alpha = 0 once stopped
                L.check(lib, lib.hipgp_vec_xr_update(dt, ptr(x), ptr(r), ptr(p), ptr(Ap), ptr(num), ptr(den), ptr(sc), 1, M, st()))
                z = Pm(r) if Pm is not None else r
                L.check(lib, lib.hipgp_vec_dot(dt, ptr(z), ptr(r), C.c_void_p(sc.data_ptr() + 8), 1, M, st()))
                red(sc)
                iters += active
                active = active & (torch.sqrt(sc[0:1]) >= tol)
                num = torch.where(active, sc[1:2], zero); den = torch.where(active, rs, one)
                L.check(lib, lib.hipgp_vec_p_update(dt, ptr(p), ptr(z), ptr(num), ptr(den), 1, M, st()))
                rs = sc[1:2].clone()
                if (n + 1) % check_every == 0 and n + 1 < maxiter and not bool(active):
                    break
            self._iters_dev = iters
        return x

    @property
    def last_iters(self):
        return int(self._iters_dev.item())
