"""Drop-in for `ziggy/misc/toeplitz_expanded.py` : `ToeplitzMatmul` (nn.Module, NO jitter, :248) and `gram_solve`."""
import numpy as np
import torch
from torch import nn

from . import _lib as L
from .plan import Plan, matvec_autograd
from .cg import conj_grad

_MODES = {"gram": L.MV_K, "RTv": L.MV_RT, "Rv": L.MV_R, "circ_inv": L.MV_CINV}


class ToeplitzMatmul(nn.Module):
    """(block) Toeplitz structured matrix defined by a list of grids and a kernel function
    (toeplitz_expanded.py:61-250)."""

    def __init__(self, xgrids, kernel, batch_shape=None):
        super(ToeplitzMatmul, self).__init__()
        self.device = xgrids[0].device
        if self.device.type != "cuda":
            raise RuntimeError("hipgp_b200.ToeplitzMatmul: xgrids must live on a CUDA device (no CPU fallback)")
        self.dims = tuple(len(xg) for xg in xgrids)
        self.ndim = len(self.dims)
        self.M = np.prod(self.dims)
        self.xgrids = xgrids
        self.K = self.toeplitz_gram(xgrids, kernel)
        self._plan = Plan(self.dims, self.K.dtype, self.device)
        self._plan.set_first_row(self.K.reshape(-1))
        self._spec_cache = {}
        self.res_idx = [slice(None)] + [slice(0, d, 1) for d in self.dims] + [0]
        self.Cc_shape = torch.Size(tuple(self._plan.embedded_dims) + (2,))
        if batch_shape is not None:
            self.batch_shape = batch_shape
            self.cvec_shape = tuple(batch_shape) + tuple(self.Cc_shape)

    def set_batch_shape(self, batch_shape):
        self.batch_shape = batch_shape
        self.cvec_shape = tuple(batch_shape) + tuple(self.Cc_shape)

    def _spec(self, name, which):
        if name not in self._spec_cache:
            d0 = self._plan.spectrum(which)
            self._spec_cache[name] = torch.stack([d0, torch.zeros_like(d0)], dim=-1)
        return self._spec_cache[name]

    D = property(lambda self: self._spec("D", L.SPEC_D))
    D_sqrt = property(lambda self: self._spec("D_sqrt", L.SPEC_D_SQRT))
    Di = property(lambda self: self._spec("Di", L.SPEC_DI))
    Di_sqrt = property(lambda self: self._spec("Di_sqrt", L.SPEC_DI_SQRT))

    @property
    def C(self):
        if "C" not in self._spec_cache:
            self._spec_cache["C"] = self.circulant_embed(self.K)
        return self._spec_cache["C"]

    def forward(self, vec, multiply_type="gram"):
        """vec: bsz x M (bsz x M' for "Rv"); multiply_type in gram | RTv | Rv | circ_inv (toeplitz_expanded.py:139-189)"""
        if multiply_type not in _MODES:
            raise NotImplementedError("gram|RTv|Rv|circ_inv")
        return matvec_autograd(self._plan, _MODES[multiply_type], vec.reshape(vec.shape[0], -1))

    # methods the fused CG path recognises
    def _matmul_by_K(self, vec):
        return self._plan.matvec(L.MV_K, vec)

    def _matmul_by_Cinv(self, vec):
        return self._plan.matvec(L.MV_CINV, vec)

    def circulant_embed(self, Ktoe):
        dims = Ktoe.shape
        for d in range(len(dims)):
            Krev = torch.flip(Ktoe, dims=(d,))
            idx = [slice(None)] * d + [slice(1, -1, 1)]
            Ktoe = torch.cat([Ktoe, Krev[tuple(idx)]], dim=d)
        return Ktoe

    def make_complex(self, vec):
        return torch.stack([vec, torch.zeros_like(vec)], dim=-1)

    def toeplitz_gram(self, xgrids, kernel):
        dims = [len(xg) for xg in xgrids]
        grid_row = getattr(kernel, "grid_row", None)
        if grid_row is not None:
            return grid_row(xgrids).reshape(dims)
        xxs = torch.meshgrid(*xgrids, indexing="ij")
        xs = torch.stack([x.reshape(-1) for x in xxs], dim=-1)
        return kernel(xs[0][None, :], xs).view(dims)


def gram_solve(xgrids, kernel_fun, vec, K_matmul=None, maxiter=20, do_precond=True, tol=1e-10, callback=None, mult_RT=True):
    """K_uu^{-1/2} v = R^T K_uu^{-1} v (or K_uu^{-1} v with mult_RT=False); vec is bsz x M
    (toeplitz_expanded.py:17-58).  The callback receives x in the reference's (M, bsz) layout."""
    assert len(vec.shape) == 2
    if K_matmul is None:
        K_matmul = ToeplitzMatmul(xgrids, kernel_fun, batch_shape=vec.shape[:-1])
    else:
        K_matmul.set_batch_shape(vec.shape[:-1])
    if isinstance(K_matmul, ToeplitzMatmul):
        precond = K_matmul._matmul_by_Cinv if do_precond else None
        cb = (lambda n, x: callback(n, x.t())) if callback is not None else None
        from .cg import conj_grad2
        d_rows = conj_grad2(K_matmul._matmul_by_K, vec, precond=precond, maxiter=maxiter, tol=tol, callback=cb)
        return K_matmul(d_rows, multiply_type="RTv") if mult_RT else d_rows
    Kmul = lambda x: K_matmul(x.t(), multiply_type="gram").t()
    precond = (lambda x: K_matmul(x.t(), multiply_type="circ_inv").t()) if do_precond else None
    d = conj_grad(Kmul, vec.t(), precond=precond, maxiter=maxiter, tol=tol, callback=callback)
    return K_matmul(d.t(), multiply_type="RTv") if mult_RT else d.t()
