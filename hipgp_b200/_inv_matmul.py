"""Drop-in for `ziggy/misc/_inv_matmul.py` : autograd Function for K^{-1} R.

forward  = fused PCG (hipgp_pcg) under no_grad, as `_inv_matmul.py:10-25`.
backward = a second PCG on grad_output for the right-hand-side gradient (`_inv_matmul.py:35-37,58-60`).
The gradient with respect to the Toeplitz column (kernel hyper-parameter learning, `_inv_matmul.py:39-55`, via the
vendored GPyTorch `sym_toeplitz_derivative_quadratic_form`) is a "next" row of SURVEY.md 8f and raises for now.
"""
import torch
from torch.autograd import Function


class InvMatmul(Function):
    @staticmethod
    def forward(ctx, toeplitz_tensor, column, right_tensor, do_precond, maxiter, tol):
        assert right_tensor.ndimension() == 2, right_tensor.ndimension()
        ctx.toeplitz_tensor = toeplitz_tensor
        with torch.no_grad():
            solves = toeplitz_tensor._solve(right_tensor, do_precond=do_precond, maxiter=maxiter, tol=tol, callback=None)
        ctx.maxiter, ctx.tol = int(maxiter), float(tol)
        ctx.save_for_backward(solves)
        return solves

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError(
                "hipgp_b200: gradient w.r.t. the Toeplitz column (learn_kernel=True) is not built yet "
                "(reference: ziggy/misc/_inv_matmul.py:39-55); SURVEY.md 8f rank 2")
        right_grad = None
        if ctx.needs_input_grad[2]:
            right_grad = InvMatmul.apply(ctx.toeplitz_tensor, ctx.toeplitz_tensor.column, grad_output, True,
                                         ctx.maxiter, ctx.tol)
        return None, None, right_grad, None, None, None
