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
    """K^-1 R as an autograd node; argument order and meaning as in the reference (`_inv_matmul.py:10`)."""

    @staticmethod
    def forward(ctx, toeplitz_tensor, column, right_tensor, do_precond, maxiter, tol):
        if right_tensor.ndimension() != 2:
            raise AssertionError(right_tensor.ndimension())
        ctx.op = toeplitz_tensor
        ctx.cg_args = (int(maxiter), float(tol))
        with torch.no_grad():
            x = toeplitz_tensor._solve(right_tensor, do_precond=do_precond, maxiter=ctx.cg_args[0], tol=ctx.cg_args[1],
                                       callback=None)
        ctx.save_for_backward(x)
        return x

    @staticmethod
    def backward(ctx, grad_output):
        (x,) = ctx.saved_tensors
        want_column, want_rhs = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (want_column or want_rhs):
            return None, None, None, None, None, None
        op = ctx.op
        maxiter, tol = ctx.cg_args
        # left solves K^-1 g, always preconditioned (`_inv_matmul.py:35-37`); differentiable again for double backward
        lhs = InvMatmul.apply(op, op.column, grad_output.contiguous(), True, maxiter, tol)
        g_column = None
        if want_column:
            with torch.no_grad():
                g_column = op._plan.toeplitz_quadform(lhs, x, scale=-1.0).view(op.column.shape)
        return None, g_column, (lhs if want_rhs else None), None, None, None
