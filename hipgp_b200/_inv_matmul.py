"""Drop-in for `ziggy/misc/_inv_matmul.py` : autograd Function for K^{-1} R.

forward  = fused PCG (hipgp_pcg) under no_grad, as `_inv_matmul.py:10-25`.
backward = a second PCG on grad_output (the "left solves", `_inv_matmul.py:35-37`), returned as the right-hand-side
gradient (`:58-60`), and -- when the Toeplitz column needs a gradient (learn_kernel=True) -- the quadratic form
`sym_toeplitz_derivative_quadratic_form([L; R], -1/2 [R; L])` of `_inv_matmul.py:39-55`.  That form is bilinear and
symmetric in its two arguments, so it equals `-1 * form(L, R)`; `hipgp_toeplitz_quadform` evaluates it with the plan's
own row / column transform passes (see csrc/corr_api.inl) instead of the reference's 1-D FFTs of length 2M - 1.
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
        right_solves, = ctx.saved_tensors
        left_solves = None
        if any(ctx.needs_input_grad):
            left_solves = InvMatmul.apply(ctx.toeplitz_tensor, ctx.toeplitz_tensor.column, grad_output.contiguous(), True,
                                          ctx.maxiter, ctx.tol)
        column_grad = None
        if ctx.needs_input_grad[1]:
            with torch.no_grad():
                column_grad = ctx.toeplitz_tensor._plan.toeplitz_quadform(left_solves, right_solves, scale=-1.0)
            column_grad = column_grad.view(ctx.toeplitz_tensor.column.shape)
        right_grad = left_solves if ctx.needs_input_grad[2] else None
        return None, column_grad, right_grad, None, None, None
