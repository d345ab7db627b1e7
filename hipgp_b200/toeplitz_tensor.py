"""Drop-in for `ziggy/misc/toeplitz_tensor.py` : class `ToeplitzTensor`.

Same constructor, attributes and methods as the reference; every matvec / solve runs in libhipgp_b200.so.
The reference evaluates `crop IFFT_N(spec * FFT_N(pad v))` with N_d = 2 m_d - 2.  Each of those four operators is a
(block-)Toeplitz product whose first column is a DCT-I of the clamped spectrum, so the CUDA path evaluates the
identical operator with a smooth embedding length; M' = prod(2 m_d - 2) and every value a caller can
observe keep the reference's meaning.
"""
import numpy as np
import torch

from . import _lib as L
from .plan import Plan, matvec_autograd
from ._inv_matmul import InvMatmul
from .cg import conj_grad2


class ToeplitzTensor:
    def __init__(self, xgrids, kernel, batch_shape=None, jitter_val=1e-3):
        self.device = xgrids[0].device
        if self.device.type != "cuda":
            raise RuntimeError("hipgp_b200.ToeplitzTensor: xgrids must live on a CUDA device (no CPU fallback)")
        self.dims = tuple(len(xg) for xg in xgrids)
        self.ndim = len(self.dims)
        self.M = np.prod(self.dims)
        self.column = self.toeplitz_gram(xgrids, kernel, jitter_val)
        if self.column.dtype not in (torch.float32, torch.float64):
            raise TypeError("kernel must return float32 or float64")
        self._plan = Plan(self.dims, self.column.dtype, self.device)
        self._plan.set_first_row(self.column)
        self._spec_cache = {}
        edims = self._plan.embedded_dims
        # padding slices for results --- batch dim, variable dims, real part (toeplitz_tensor.py:35-39)
        self.res_idx = [slice(None)] + [slice(0, d, 1) for d in self.dims] + [0]
        self.Cc_shape = torch.Size(tuple(edims) + (2,))
        if batch_shape is not None:
            self.batch_shape = batch_shape
            self.cvec_shape = tuple(batch_shape) + tuple(self.Cc_shape)

    # ---- lazily materialised reference attributes ------------------------------------------------
    @property
    def C(self):
        if "C" not in self._spec_cache:
            self._spec_cache["C"] = self.circulant_embed(self.column.view(self.dims))
        return self._spec_cache["C"]

    def _spec(self, name, which):
        if name not in self._spec_cache:
            d0 = self._plan.spectrum(which)
            self._spec_cache[name] = torch.stack([d0, torch.zeros_like(d0)], dim=-1)
        return self._spec_cache[name]

    @property
    def D(self):
        return self._spec("D", L.SPEC_D)

    @property
    def D_sqrt(self):
        return self._spec("D_sqrt", L.SPEC_D_SQRT)

    @property
    def Di(self):
        return self._spec("Di", L.SPEC_DI)

    @property
    def Di_sqrt(self):
        return self._spec("Di_sqrt", L.SPEC_DI_SQRT)

    # ---- solves ------------------------------------------------------------------------------------
    def inv_matmul(self, right_tensor, do_precond=True, maxiter=20, tol=1e-8):
        """compute A^{-1}R, where self = A  (toeplitz_tensor.py:47-52)"""
        return InvMatmul.apply(self, self.column, right_tensor, do_precond, maxiter, tol)

    def _solve(self, vec, do_precond=True, maxiter=100, tol=1e-8, callback=None):
        """vec: (bsz, M) -> d: (bsz, M) = K^{-1} vec by (P)CG (toeplitz_tensor.py:54-68)"""
        assert len(vec.shape) == 2
        self.set_batch_shape(vec.shape[:-1])
        precond = self._matmul_by_Cinv if do_precond else None
        return conj_grad2(self._matmul_by_K, vec, precond=precond, maxiter=maxiter, tol=tol, callback=callback)

    # ---- matvecs -----------------------------------------------------------------------------------
    def _matmul_by_K(self, vec):
        return matvec_autograd(self._plan, L.MV_K, vec, self.column)

    def _matmul_by_RT(self, vec):
        return matvec_autograd(self._plan, L.MV_RT, vec, self.column)

    def _matmul_by_R(self, vec):
        return matvec_autograd(self._plan, L.MV_R, vec.reshape(vec.shape[0], -1), self.column)

    def _matmul_by_Cinv(self, vec):
        return matvec_autograd(self._plan, L.MV_CINV, vec, self.column)

    # ---- construction helpers (same names as the reference) ------------------------------------
    def toeplitz_gram(self, xgrids, kernel, jitter_val):
        """first row k(u_0, u_.) (+ jitter at [0]); `kernel` is any callable (x, y) -> K.  When it was made by
        `hipgp_b200.hipgp` it carries `grid_row`, which evaluates the row from the 1-D grids without a meshgrid."""
        grid_row = getattr(kernel, "grid_row", None)
        if grid_row is not None:
            Krow = grid_row(xgrids).reshape(1, -1)
        else:
            xxs = torch.meshgrid(*xgrids, indexing="ij")
            xs = torch.stack([x.reshape(-1) for x in xxs], dim=-1)
            Krow = kernel(xs[0][None, :], xs)
        Krow = Krow.clone()
        if jitter_val is not None:
            Krow[0, 0] += jitter_val
        return Krow.reshape(-1)

    def circulant_embed(self, Ktoe):
        dims = Ktoe.shape
        for d in range(len(dims)):
            Krev = torch.flip(Ktoe, dims=(d,))
            idx = [slice(None)] * d + [slice(1, -1, 1)]
            Ktoe = torch.cat([Ktoe, Krev[tuple(idx)]], dim=d)
        return Ktoe

    def make_complex(self, vec):
        return torch.stack([vec, torch.zeros_like(vec)], dim=-1)

    def zero_pad_vec_batch_comp(self, vec):
        cvec = torch.zeros(self.cvec_shape, dtype=self.column.dtype, device=self.device)
        cvec[tuple(self.res_idx)] = vec.reshape((vec.shape[0],) + self.dims)
        return cvec

    def complex_mult(self, t1, t2):
        real1, imag1 = t1[..., 0], t1[..., 1]
        real2, imag2 = t2[..., 0], t2[..., 1]
        return torch.stack([real1 * real2 - imag1 * imag2, real1 * imag2 + imag1 * real2], dim=-1)

    def set_batch_shape(self, batch_shape):
        self.batch_shape = batch_shape
        self.cvec_shape = tuple(batch_shape) + tuple(self.Cc_shape)
