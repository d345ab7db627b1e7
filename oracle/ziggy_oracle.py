"""CPU oracle for the HIP-GP structured-kernel hot path.  TEST INFRASTRUCTURE ONLY.

This is a torch-CPU restatement of the reference algorithm (suyashk12/hipgp, package `ziggy`),
op for op, so that it can stand in for the reference on the GPU box where `/root/reference` does not
exist.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it -- and only as the checker or the timed CPU baseline.  Nothing in `hipgp_b200/`
imports it; the product path has no CPU fallback.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the UNMODIFIED reference (under
`oracle/ref_shim.py`) in the build container and commits its outputs as `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this file against those vectors (bit-exact for the FFT matvecs and
PCG iterates, because both sides issue the same torch CPU ops in the same order), and
`tests/test_oracle_vs_reference.py` re-runs the comparison live when the reference tree is present.

The FFT itself is third-party (PyTorch; the reference pins torch==1.4.0, requirements.txt:9, and calls
the legacy `torch.fft(x, signal_ndim)`).  Here it is `torch.fft.fftn/ifftn` on the complex view, which is
what the legacy function computed: unnormalised forward, 1/N inverse.

Each function cites the reference file:line it follows.
"""
import math

import numpy as np
import torch


# --------------------------------------------------------------------------------------------------
# legacy complex helpers (trailing real dimension of size 2)
# --------------------------------------------------------------------------------------------------
def _fft(x, signal_ndim):
    """legacy torch.fft(x, signal_ndim): toeplitz_tensor.py:25,79"""
    xc = torch.view_as_complex(x.contiguous())
    dims = tuple(range(xc.dim() - signal_ndim, xc.dim()))
    return torch.view_as_real(torch.fft.fftn(xc, dim=dims))


def _ifft(x, signal_ndim):
    """legacy torch.ifft(x, signal_ndim): toeplitz_tensor.py:82"""
    xc = torch.view_as_complex(x.contiguous())
    dims = tuple(range(xc.dim() - signal_ndim, xc.dim()))
    return torch.view_as_real(torch.fft.ifftn(xc, dim=dims))


def make_complex(vec):
    """toeplitz_tensor.py:145-147"""
    return torch.stack([vec, torch.zeros_like(vec)], dim=-1)


def circulant_embed(Ktoe):
    """toeplitz_tensor.py:135-143 / toeplitz_expanded.py:191-199: per dim cat([K, flip(K)[1:-1]])."""
    dims = Ktoe.shape
    for d in range(len(dims)):
        Krev = torch.flip(Ktoe, dims=(d,))
        idx = [slice(None)] * d + [slice(1, -1, 1)]
        Krev = Krev[tuple(idx)]
        Ktoe = torch.cat([Ktoe, Krev], dim=d)
    return Ktoe


def toeplitz_gram(xgrids, kernel, jitter_val=None):
    """First row k(u_0, u_.) on the C-order meshgrid.
    toeplitz_tensor.py:127-133 (jitter added at [0,0]) / toeplitz_expanded.py:242-250 (no jitter)."""
    xxs = torch.meshgrid(*xgrids, indexing="ij")
    xs = torch.stack([x.reshape(-1) for x in xxs], dim=-1)
    Krow = kernel(xs[0][None, :], xs)
    if jitter_val is not None:
        Krow[0, 0] += jitter_val
    return Krow.squeeze()


class OracleToeplitz:
    """Restates `ToeplitzTensor` (toeplitz_tensor.py:7-169) and, with `jitter_val=None`,
    `ToeplitzMatmul` (toeplitz_expanded.py:61-250; no jitter, :248)."""

    def __init__(self, xgrids, kernel, batch_shape=None, jitter_val=1e-3):
        self.column = toeplitz_gram(xgrids, kernel, jitter_val)
        self.dims = tuple(len(xg) for xg in xgrids)
        self.ndim = len(self.dims)
        self.M = int(np.prod(self.dims))
        self.C = circulant_embed(self.column.view(self.dims))
        Cc = make_complex(self.C)
        # toeplitz_tensor.py:23-33
        D0 = _fft(Cc, self.ndim)
        D0 = D0[..., 0].clamp(min=1e-6)
        D1 = torch.zeros_like(D0)
        self.D = torch.stack([D0, D1], dim=-1)
        self.D_sqrt = torch.stack([torch.sqrt(self.D[..., 0]), D1], dim=-1)
        Di0 = 1. / self.D[..., 0]
        self.Di = torch.stack([Di0, D1], dim=-1)
        self.Di_sqrt = torch.sqrt(self.Di)
        self.res_idx = tuple([slice(None)] + [slice(0, d, 1) for d in self.dims] + [0])
        self.Cc_shape = Cc.shape
        self.batch_shape = None
        if batch_shape is not None:
            self.set_batch_shape(batch_shape)

    def set_batch_shape(self, batch_shape):
        """toeplitz_tensor.py:167-169"""
        self.batch_shape = batch_shape
        self.cvec_shape = tuple(batch_shape) + tuple(self.Cc_shape)

    def zero_pad_vec_batch_comp(self, vec):
        """toeplitz_tensor.py:149-152"""
        cvec = torch.zeros(self.cvec_shape, dtype=self.column.dtype)
        cvec[self.res_idx] = vec.reshape((vec.shape[0],) + self.dims)
        return cvec

    def complex_mult(self, t1, t2):
        """toeplitz_tensor.py:154-165 (batch_shape set => temp filled through 4 real slices)"""
        real1, imag1 = t1[..., 0], t1[..., 1]
        real2, imag2 = t2[..., 0], t2[..., 1]
        tmp = torch.zeros(self.cvec_shape, dtype=self.column.dtype)
        tmp[..., 0] = real1 * real2 - imag1 * imag2
        tmp[..., 1] = real1 * imag2 + imag1 * real2
        return tmp

    def _apply(self, spec, cvec):
        Fv = _fft(cvec, self.ndim)
        prod = self.complex_mult(spec, Fv)
        return _ifft(prod, self.ndim)

    def matmul_K(self, vec):
        """toeplitz_tensor.py:70-83"""
        self.set_batch_shape(vec.shape[:-1])
        cres = self._apply(self.D, self.zero_pad_vec_batch_comp(vec))
        return cres[self.res_idx].reshape(vec.shape[0], -1)

    def matmul_Cinv(self, vec):
        """toeplitz_tensor.py:114-125 -- the HIP-GP preconditioner"""
        self.set_batch_shape(vec.shape[:-1])
        cres = self._apply(self.Di, self.zero_pad_vec_batch_comp(vec))
        return cres[self.res_idx].reshape(vec.shape[0], -1)

    def matmul_RT(self, vec):
        """toeplitz_tensor.py:85-97"""
        self.set_batch_shape(vec.shape[:-1])
        cres = self._apply(self.D_sqrt, self.zero_pad_vec_batch_comp(vec))
        return cres[..., 0].reshape(vec.shape[0], -1)

    def matmul_R(self, vec):
        """toeplitz_tensor.py:99-112 (input already has the embedded length, no padding)"""
        self.set_batch_shape(vec.shape[:-1])
        cvec = make_complex(vec.reshape((vec.shape[0],) + tuple(self.C.shape)))
        cres = self._apply(self.D_sqrt, cvec)
        return cres[self.res_idx].reshape(vec.shape[0], -1)

    def solve(self, vec, do_precond=True, maxiter=100, tol=1e-8, callback=None):
        """toeplitz_tensor.py:54-68 (`_solve`)"""
        assert len(vec.shape) == 2
        precond = self.matmul_Cinv if do_precond else None
        return conj_grad2(self.matmul_K, vec, precond=precond, maxiter=maxiter, tol=tol, callback=callback)


# --------------------------------------------------------------------------------------------------
# CG / PCG   (ziggy/misc/cg.py)
# --------------------------------------------------------------------------------------------------
def conj_grad(A_mul, b, precond=None, maxiter=20, tol=1e-10, callback=None):
    """cg.py:5-41 -- column layout, b is (M, L)."""
    if precond is None:
        precond = lambda x: x
    x = torch.zeros_like(b)
    r = b - A_mul(x)
    z = precond(r)
    p = z
    for n in range(maxiter):
        rs = torch.sum(r * z, dim=0)
        Ap = A_mul(p)
        alpha = rs / torch.sum(p * Ap, dim=0)
        x = x + alpha[None, :] * p
        r = r - alpha[None, :] * Ap
        rnew = torch.sum(r * r, dim=0)
        if torch.all(torch.sqrt(rnew) < tol):
            break
        z = precond(r)
        beta = torch.sum(z * r, dim=0) / rs
        p = z + beta[None, :] * p
        if callback is not None:
            callback(n, x)
    return x


def conj_grad2(A_mul, b, precond=None, maxiter=20, tol=1e-10, callback=None):
    """cg.py:44-80 -- row layout, b is (bsz, M)."""
    if precond is None:
        precond = lambda x: x
    x = torch.zeros_like(b)
    r = b - A_mul(x)
    z = precond(r)
    p = z
    for n in range(maxiter):
        rs = torch.sum(r * z, dim=1)
        Ap = A_mul(p)
        alpha = rs / torch.sum(p * Ap, dim=1)
        x = x + alpha.unsqueeze(-1) * p
        r = r - alpha.unsqueeze(-1) * Ap
        rnew = torch.sum(r * r, dim=1)
        if torch.all(torch.sqrt(rnew) < tol):
            break
        z = precond(r)
        beta = torch.sum(z * r, dim=1) / rs
        p = z + beta.unsqueeze(-1) * p
        if callback is not None:
            callback(n, x)
    return x


def gram_solve(xgrids, kernel_fun, vec, K=None, maxiter=20, do_precond=True, tol=1e-10,
               callback=None, mult_RT=True):
    """toeplitz_expanded.py:17-58: PCG through `conj_grad` on transposed views, then optional R^T.
    `ToeplitzMatmul` carries no jitter (toeplitz_expanded.py:248)."""
    assert len(vec.shape) == 2
    if K is None:
        K = OracleToeplitz(xgrids, kernel_fun, batch_shape=vec.shape[:-1], jitter_val=None)
    Kmul = lambda x: K.matmul_K(x.t()).t()
    precond = (lambda x: K.matmul_Cinv(x.t()).t()) if do_precond else None
    d = conj_grad(Kmul, vec.t(), precond=precond, maxiter=maxiter, tol=tol, callback=callback)
    if mult_RT:
        return K.matmul_RT(d.t())
    return d.t()


# --------------------------------------------------------------------------------------------------
# kernels   (ziggy/kernels.py, ziggy/misc/stats.py, ziggy/exact_gp_1d_derivatives.py)
# --------------------------------------------------------------------------------------------------
def sqexp(x, y, sig2, ell):
    """kernels.py:73-79"""
    sqdist = torch.sum(((x[:, None, :] - y[None, :, :]) / ell) ** 2, dim=-1)
    return sig2 * torch.exp(-sqdist / 2)


def matern(x, y, sig2, ell, nu):
    """kernels.py:145-158 -- note ell divides the UNSCALED Euclidean distance (:149)."""
    sqdist = torch.sum((x[:, None, :] - y[None, :, :]) ** 2, dim=-1)
    if nu == .5:
        kmat = torch.exp(-torch.sqrt(sqdist) / ell)
    elif nu == 1.5:
        dp = np.sqrt(3) * torch.sqrt(sqdist) / ell
        kmat = (1 + dp) * torch.exp(-dp)
    elif nu == 2.5:
        dp = np.sqrt(5) * torch.sqrt(sqdist) / ell
        kmat = (1 + dp + (5. / 3.) * sqdist / (ell ** 2)) * torch.exp(-dp)
    else:
        raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
    return sig2 * kmat


def gneiting(x, y, sig2, ell, alpha=1.):
    """kernels.py:108-117 -- compact support, t > 1 -> 0."""
    dist = torch.sqrt(torch.sum(((x[:, None, :] - y[None, :, :]) / ell) ** 2, dim=-1))
    t = dist
    cterms = (1 - t) * torch.cos(np.pi * t) + (1 / np.pi) * torch.sin(np.pi * t)
    cij = (1 + t ** alpha) ** (-3) * cterms
    cij[t > 1.] = 0.
    return sig2 * cij


def normal_cdf(x, loc, scale):
    """stats.py:74-76"""
    sqrt2 = np.sqrt(2)
    return .5 * (1. + torch.erf((x - loc) / (scale * sqrt2)))


def semi_integrated_sqe(xintegrated, x, sig2, Sinv):
    """kernels.py:223-237 -- analytic int_0^1 k(u, alpha x) d alpha * |x| for SqExp; integrates over the
    FIRST argument; returns (num_xi, num_x)."""
    sqrt2pi = np.sqrt(2 * np.pi)
    xdists = torch.sqrt(torch.sum(xintegrated * xintegrated, dim=-1))
    a = torch.sum(torch.matmul(xintegrated, Sinv) * xintegrated, dim=-1)
    xint_Si = torch.matmul(xintegrated, Sinv)
    b = torch.matmul(xint_Si[:, None, None, :], x[None, :, :, None]).squeeze()
    c = torch.sum(torch.matmul(x, Sinv) * x, dim=-1)
    scale = torch.sqrt(1 / a[:, None])
    loc = b / a[:, None]
    coef = sig2 * torch.exp((b ** 2) / (2 * a[:, None]) - c / 2) * sqrt2pi * scale
    ca = normal_cdf(1, loc, scale)
    cb = normal_cdf(0, loc, scale)
    return coef * (ca - cb) * xdists[:, None]


def sqexp_k_semi(xpoint, xintegrated, sig2, ell, dtype):
    """kernels.py:85-90 -- returns (Npoint, Nintegrated)."""
    D = xpoint.shape[1]
    Sinv = (1. / (ell ** 2)) * torch.eye(D, dtype=dtype)
    return semi_integrated_sqe(xintegrated, xpoint, sig2, Sinv).transpose(0, 1)


def k_semi_mc(forward, xpoint, xintegrated, alphas):
    """kernels.py:19-39 with the random stratified grid `alphas` (= arange(npts)/npts + U/npts,
    :25-27) passed in, so that callers can share one RNG draw.  Returns (Npoint, Nintegrated)."""
    Np, D = xpoint.shape
    Ni, D = xintegrated.shape
    npts = alphas.shape[0]
    xgrid = xintegrated[:, None, :] * alphas[None, :, None]
    Kpis = forward(xpoint, xgrid.reshape(-1, D))
    Kpis = Kpis.reshape(Np, Ni, npts)
    dists = xintegrated.pow(2.).sum(dim=-1).sqrt()
    return torch.mean(Kpis, dim=-1) * dists[None, :]


def doubly_diag_interp(x, sig2, ell, distance_grid, slopes, knn):
    """kernels.py:199-218 (`KernelDoublyDiagInterpolator.forward`).  The 50-entry table comes from
    scipy dblquad at ctor time (kernels.py:183-197) and is an INPUT here.
    dist == 0 gives lower_i = -1, which wraps to the last table entry -- reproduced."""
    dists = torch.sqrt(torch.sum((x / ell) ** 2, dim=-1))
    lower_i = torch.sum(dists[:, None] > distance_grid, dim=-1) - 1
    diff = dists - distance_grid[lower_i]
    ivals = knn[lower_i] + slopes[lower_i] * diff
    return ell * ell * sig2 * ivals


def deriv_k(x, y, sig2, ell):
    """exact_gp_1d_derivatives.py:9-12"""
    diff = x[:, None] - y[None, ]
    return sig2 * torch.exp(-1 / 2 * diff ** 2 / ell ** 2)


def deriv_kprime(x, y, sig2, ell):
    """exact_gp_1d_derivatives.py:19-23"""
    diff = x[:, None] - y[None, ]
    Kxy = sig2 * torch.exp(-1 / 2 * diff ** 2 / ell ** 2)
    return -diff / (ell ** 2) * Kxy


def deriv_kprime_double_full(x, y, sig2, ell):
    """exact_gp_1d_derivatives.py:32-38"""
    diff = x[:, None] - y[None, ]
    diff_sq = diff ** 2
    ell_sq = ell ** 2
    Kxy = sig2 * torch.exp(-1 / 2 * diff_sq / ell_sq)
    return Kxy / ell_sq * (1 - 1 / ell_sq * diff_sq)


# --------------------------------------------------------------------------------------------------
# callers on the path   (ziggy/hipgp.py:117-146, ziggy/svi_gp.py:48-76)
# --------------------------------------------------------------------------------------------------
def compute_kn(xgrids, kfun, Knm, maxiter_cg=10, tol=1e-8, jitter_val=1e-3):
    """hipgp.py:139-146, ziggy branch: k_n = R^T K_uu^{-1} K_un (PCG with the HIP-GP preconditioner)."""
    Kmm = OracleToeplitz(xgrids, kfun, batch_shape=None, jitter_val=jitter_val)
    d0 = Kmm.solve(Knm, do_precond=True, maxiter=maxiter_cg, tol=tol)
    return Kmm.matmul_RT(d0)


def meshgrid_points(xgrids):
    """hipgp.py:63-65"""
    xxs = torch.meshgrid(*xgrids, indexing="ij")
    return torch.stack([x.reshape(-1) for x in xxs], dim=-1)


def meanfield_elbo_and_grad(xgrids, kfun, Knm, Knn_diag, ybatch, noise_std_batch, theta1, theta2, num_obs, maxiter_cg=10,
                            jitter_val=1e-3):
    """hipgp.py:194-276 (mean-field branch) + compute_batch_an (hipgp.py:370-414) + stats.diag_kl_to_standard.
    Returns (elbo_estimate, theta1.grad, theta2.grad)."""
    kn = compute_kn(xgrids, kfun, Knm, maxiter_cg=maxiter_cg, jitter_val=jitter_val)
    qS = -0.5 * 1 / theta2                      # hipgp.py:499-503
    qm = qS * theta1
    y = ybatch.squeeze()
    Knn = Knn_diag.squeeze()
    knt_kn = torch.sum(kn * kn, dim=-1).squeeze()
    knt_m = kn.matmul(qm).squeeze()
    knSkn = torch.sum((kn * qS.t()) * kn, dim=-1).squeeze()
    ivar = (1 / (noise_std_batch ** 2)).squeeze()
    log_noise_std = torch.log(noise_std_batch)
    mse = (knt_m - y) ** 2
    variance = Knn - knt_kn + knSkn
    batch_an = -0.5 * ivar * (mse + variance) - log_noise_std - 0.5 * np.log(2 * np.pi)
    kl = .5 * (torch.sum(qS) + torch.sum(qm * qm) - torch.sum(torch.log(qS)) - len(qm))
    elbo = torch.mean(batch_an) - (kl / num_obs)
    bscale = num_obs / Knm.shape[0]
    ivar_noise = (1 / (noise_std_batch ** 2))
    knt_m2 = kn.matmul(qm)
    bdiff = ivar_noise * (knt_m2 - ybatch)
    data_dm = -torch.matmul(bdiff.t(), kn).t()
    dm = bscale * data_dm - qm
    lam_diag = bscale * torch.sum(ivar_noise * kn * kn, dim=0) + 1
    dS = -.5 * lam_diag[:, None] - theta2
    deta1 = dm + dS * (-2 * qm)
    return elbo, -deta1, -dS


def batch_indices(n, batch_size):
    """svi_gp.py:81-85 (batch_predict): slices of the prediction batches, restated verbatim."""
    num_batches = int(np.ceil(n / batch_size))

    def one(it):
        idx = it % num_batches
        return slice(idx * batch_size, min((idx + 1) * batch_size, n))
    return [one(i) for i in range(num_batches)]


def toeplitz_matmul_1d(column, row, vec):
    """T v for the Toeplitz matrix with first column `column` and first row `row` (gpt_toeplitz.py:96-154): circulant
    embedding of size 2n-1 = [column, reversed(row[1:])], product in the Fourier domain, first n outputs."""
    n = column.shape[-1]
    c = np.concatenate([column, row[1:][::-1]])
    vp = np.zeros(2 * n - 1, dtype=np.float64); vp[:n] = vec
    return np.real(np.fft.ifft(np.fft.fft(c) * np.fft.fft(vp)))[:n]


def sym_toeplitz_derivative_quadratic_form(left_vectors, right_vectors):
    """sum_j u_j^T (dT/dc_i) v_j for all i (gpt_toeplitz.py:169-209); inputs (M, S) like the reference (columns are the
    vectors), fp64 numpy.  Two Toeplitz products per pair -- an upper-triangular one built from u and one from its
    reversal -- and the diagonal correction of element 0, exactly in the reference's order."""
    L = np.asarray(left_vectors, dtype=np.float64); R = np.asarray(right_vectors, dtype=np.float64)
    if L.ndim == 1:
        L = L[:, None]; R = R[:, None]
    M, S = L.shape
    res = np.zeros(M)
    for j in range(S):
        u, v = L[:, j], R[:, j]
        col = np.zeros(M); col[0] = u[0]
        res += toeplitz_matmul_1d(col, u, v)                      # gpt_toeplitz.py:199-201
        ur = u[::-1]
        col = np.zeros(M); col[0] = ur[0]
        res += toeplitz_matmul_1d(col, ur, v[::-1])               # :202-204
    res[0] -= np.sum(L * R)                                       # :207
    return res


def inv_matmul_backward(left_solves, right_solves):
    """column gradient of InvMatmul.backward (_inv_matmul.py:39-55); left/right solves are (B, M)."""
    Ls = np.asarray(left_solves, dtype=np.float64); Rs = np.asarray(right_solves, dtype=np.float64)
    left_vecs = np.concatenate([Ls, Rs], 0).T
    right_vecs = np.concatenate([Rs, Ls], 0).T * (-0.5)
    return sym_toeplitz_derivative_quadratic_form(left_vecs, right_vecs)


# ---- block-diagonal variational family (hipgp.py:527-690, util.py:79-126, stats.py:15-29) --------------------------
def define_block_chunks(lens, chunk_sizes):
    """util.py:79-117: (num_blocks, block_size) flat indices; blocks enumerated with axis 0 outermost, points inside a
    block in row-major order of the chunk."""
    chunks = [np.split(np.arange(n), n // c) for n, c in zip(lens, chunk_sizes)]
    out = []
    if len(lens) == 2:
        for bx in chunks[0]:
            for by in chunks[1]:
                xx, yy = np.meshgrid(bx, by, indexing="ij")
                out.append((xx * lens[1] + yy).reshape(-1))
    else:
        for bx in chunks[0]:
            for by in chunks[1]:
                for bz in chunks[2]:
                    xx, yy, zz = np.meshgrid(bx, by, bz, indexing="ij")
                    out.append((xx * (lens[1] * lens[2]) + yy * lens[2] + zz).reshape(-1))
    return np.stack(out, 0)


def block_get_lam(blk_idx, ivar_noise, kn, bscale=1.0, add_identity=True):
    """hipgp.py:666-685"""
    blk_kn = kn[..., torch.as_tensor(blk_idx)].transpose(0, 1)                  # (num_blocks, bsz, block_size)
    lam = bscale * torch.matmul(blk_kn.transpose(1, 2), ivar_noise * blk_kn)
    if add_identity:
        lam = lam + torch.eye(blk_idx.shape[1], dtype=kn.dtype)
    return lam


def block_diag_multiply(blk_idx, S_block, v):
    """hipgp.py:640-652: from_blocks(S_block @ to_blocks(v))"""
    idx = torch.as_tensor(blk_idx)
    Sv_block = S_block.matmul(v[..., idx][..., None])
    rev = torch.argsort(idx.flatten())
    return Sv_block.flatten(start_dim=1)[..., rev]


def block_kl_to_standard(blk_m, blk_S):
    """stats.py:15-29"""
    I = torch.eye(blk_S.shape[1], dtype=blk_S.dtype)
    Schol = torch.linalg.cholesky(blk_S + 1e-4 * I)
    lndet = 2.0 * torch.sum(torch.sum(torch.log(torch.diagonal(Schol, dim1=-2, dim2=-1)), dim=-1))
    n_blk, blk_size, _ = blk_S.shape
    return .5 * (torch.sum(torch.diagonal(blk_S, dim1=-2, dim2=-1)) + torch.sum(blk_m * blk_m) - lndet - n_blk * blk_size)


def block_elbo_and_grad(xgrids, kfun, blk_idx, Knm, Knn_diag, ybatch, noise_std_batch, theta1, theta2, num_obs, maxiter_cg=10,
                        jitter_val=1e-3):
    """hipgp.py:194-276 (block branch :251-261) + compute_batch_an (:370-414) + standard_variational_params (:631-638).
    Returns (elbo_estimate, theta1.grad, theta2.grad, kn, qm, qS)."""
    kn = compute_kn(xgrids, kfun, Knm, maxiter_cg=maxiter_cg, jitter_val=jitter_val)
    qS = torch.inverse(-2 * theta2)
    qm = block_diag_multiply(blk_idx, qS, theta1.t()).t()
    y = ybatch.squeeze(); Knn = Knn_diag.squeeze()
    knt_kn = torch.sum(kn * kn, dim=-1).squeeze()
    knt_m = kn.matmul(qm).squeeze()
    knSkn = torch.sum(kn * block_diag_multiply(blk_idx, qS, kn), dim=-1).squeeze()
    ivar = (1 / (noise_std_batch ** 2)).squeeze()
    batch_an = -0.5 * ivar * ((knt_m - y) ** 2 + Knn - knt_kn + knSkn) - torch.log(noise_std_batch).squeeze() - 0.5 * np.log(2 * np.pi)
    elbo = torch.mean(batch_an) - block_kl_to_standard(qm, qS) / num_obs
    bscale = num_obs / Knm.shape[0]
    ivar_noise = 1 / (noise_std_batch ** 2)
    bdiff = ivar_noise * (kn.matmul(qm) - ybatch)
    dm = bscale * (-torch.matmul(bdiff.t(), kn).t()) - qm
    lam_block = block_get_lam(blk_idx, ivar_noise, kn, bscale=bscale, add_identity=True)
    dS = -.5 * lam_block - theta2
    deta1 = dm + block_diag_multiply(blk_idx, dS, -2 * qm[None, :, 0]).squeeze().unsqueeze(-1)
    return elbo, -deta1, -dS, kn, qm, qS
