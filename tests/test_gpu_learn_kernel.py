"""learn_kernel=True / learn_noise=True end to end (ziggy/hipgp.py:208-218, ziggy/svi_gp.py:317-329):
after `elbo_and_grad` the ELBO estimate stays on the autograd tape down to log_sig2 / log_ell / log_noise2 and `(-elbo).backward()`
fills their gradients.  The nodes are custom autograd Functions around the CUDA kernels:

  K_xu, first row   hipgp_kxu_param_grad      (closed-form dk/dsig2, dk/dell reduced on the fly)
  K^-1 .            InvMatmul.backward        (second PCG + hipgp_toeplitz_quadform for the Toeplitz column)
  R^T .             hipgp_rt_column_grad      (the spectrum's dependence on the column through D^(1/2), with torch.clamp's gradient)

Golden vectors: tests/golden/learn_kernel_<dtype>.npz, produced by the UNMODIFIED reference's autograd (make_golden.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DT = {"f32": torch.float32, "f64": torch.float64}
DEV = "cuda:0"


def relerr(a, b):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.mark.parametrize("tag", ["rt2d", "rt3d"])
def test_rt_column_gradient_vs_reference_autograd(tag, golden_dir):
    """d/dc sum(G . R^T v): fp64 against the reference's autograd; the fp32 kernels on the same (fp64 fixture) inputs.
    Both fixtures have CLAMPED eigenvalues (17 of 352 / 174 of 480): torch.clamp passes no gradient there."""
    from hipgp_b200.plan import Plan
    g = np.load(os.path.join(golden_dir, "learn_kernel_f64.npz"))
    dims = [int(x[2]) for x in g[tag + "_grids"]]
    assert int(g[tag + "_nclamped"]) > 0
    for dt, tol in ((torch.float64, 1e-9), (torch.float32, 2e-3)):
        plan = Plan(dims, dt, DEV).set_first_row(torch.from_numpy(g[tag + "_column"]).to(DEV, dt))
        got = plan.rt_column_grad(torch.from_numpy(g[tag + "_v"]).to(DEV, dt), torch.from_numpy(g[tag + "_G"]).to(DEV, dt))
        assert relerr(got, g[tag + "_gcol"]) < tol, (str(dt), relerr(got, g[tag + "_gcol"]))


def test_rt_matvec_autograd_node(golden_dir):
    """ToeplitzTensor._matmul_by_RT under autograd: gradient of the vector (= R g) and of the column in one backward."""
    from hipgp_b200.toeplitz_tensor import ToeplitzTensor
    g = np.load(os.path.join(golden_dir, "learn_kernel_f64.npz"))
    grids = g["rt2d_grids"]
    xg = [torch.linspace(lo, hi, int(m), dtype=torch.float64, device=DEV) for lo, hi, m in grids]
    col = torch.from_numpy(g["rt2d_column"]).to(DEV)
    # build the drop-in from the fixture's column through a callable kernel (first row = column, no jitter added again)
    M = col.numel()
    kfun = lambda x, y: col.reshape(1, M).clone()
    tt = ToeplitzTensor(xg, kfun, batch_shape=None, jitter_val=None)
    tt.column = tt.column.detach().requires_grad_(True)
    v = torch.from_numpy(g["rt2d_v"]).to(DEV).requires_grad_(True)
    G = torch.from_numpy(g["rt2d_G"]).to(DEV)
    y = tt._matmul_by_RT(v)
    (y * G).sum().backward()
    assert relerr(tt.column.grad, g["rt2d_gcol"]) < 1e-9
    assert relerr(v.grad, tt._plan.matvec(3, G).cpu().numpy()) < 1e-12      # adjoint of R^T is R


@pytest.mark.parametrize("dname", ["f64", "f32"])
@pytest.mark.parametrize("tag", ["matern32", "matern52", "sqexp_noise"])
def test_hyperparameter_gradients_vs_reference(tag, dname, golden_dir):
    from hipgp_b200 import hipgp as hh, kernels as hk
    g = np.load(os.path.join(golden_dir, "learn_kernel_%s.npz" % dname))
    dtype = DT[dname]
    sig2, ell, noise2, jitter, nobs = [float(t) for t in g[tag + "_params"]]
    xg = [torch.linspace(lo, hi, int(m), dtype=dtype) for lo, hi, m in g[tag + "_grids"]]
    kern = hk.SqExp(dtype=dtype) if tag.startswith("sqexp") else hk.Matern(nu=1.5 if tag == "matern32" else 2.5, dtype=dtype)
    learn_noise = tag.endswith("noise")
    mod = hh.MeanFieldToeplitzGP(kern, xg, num_obs=int(nobs), sig2_init=sig2, ell_init=ell, noise2_init=noise2, dtype=dtype,
                                 jitter_val=jitter, learn_kernel=True, learn_noise=learn_noise)
    mod.global_theta1.data.copy_(torch.from_numpy(g[tag + "_theta1"])); mod.global_theta2.data.copy_(torch.from_numpy(g[tag + "_theta2"]))
    mod = mod.cuda_params(0)
    x = torch.from_numpy(g[tag + "_x"]).to(DEV); y = torch.from_numpy(g[tag + "_y"]).to(DEV)
    nb = None if learn_noise else torch.from_numpy(g[tag + "_noise_std"]).to(DEV)
    elbo = mod.elbo_and_grad(x, y, nb, maxiter_cg=60)
    assert elbo.requires_grad
    (-elbo).backward()
    # fp64: 1e-6 (both sides solve to tol 1e-8 with 60 iterations); fp32: the gradient passes through 1 / (2 sqrt(D)) of the
    # smallest eigenvalues, which fp32 resolves to a few per cent (north star's 1e-5 holds for the forward operators)
    tol = 1e-6 if dname == "f64" else 5e-2
    assert abs(float(elbo) - float(g[tag + "_elbo"])) <= (1e-8 if dname == "f64" else 2e-3) * abs(float(g[tag + "_elbo"]))
    for name, par in (("g_log_sig2", mod.log_sig2), ("g_log_ell", mod.log_ell)):
        want = float(g[tag + "_" + name])
        assert par.grad is not None and abs(float(par.grad) - want) <= tol * max(abs(want), 1e-3), (name, float(par.grad), want)
    if learn_noise:
        want = float(g[tag + "_g_log_noise2"])
        assert abs(float(mod.log_noise2.grad) - want) <= tol * abs(want)
    # the natural gradients of the variational parameters are the same as without hyper-parameter learning
    assert relerr(mod.global_theta1.grad, g[tag + "_g1"]) < (1e-6 if dname == "f64" else 2e-3)
    assert relerr(mod.global_theta2.grad, g[tag + "_g2"]) < (1e-6 if dname == "f64" else 2e-3)
    # one Adam step on the hyper-parameters as svigp_fit does (svi_gp.py:254-262,327-329) moves them
    opt = torch.optim.Adam([mod.log_ell, mod.log_sig2], lr=1e-2)
    before = float(mod.log_ell)
    opt.step()
    assert float(mod.log_ell) != before


def test_kxu_param_grad_matches_autograd_of_the_formula():
    """hipgp_kxu_param_grad against torch autograd through the oracle's kernel formulas (SqExp with per-axis ell, Matern 1/2, 3/2, 5/2)."""
    from hipgp_b200 import kernels as hk
    from oracle import ziggy_oracle as zo
    dtype = torch.float64
    torch.manual_seed(5)
    xg = [torch.linspace(0, 1, 7, dtype=dtype, device=DEV), torch.linspace(-1, 1, 9, dtype=dtype, device=DEV), torch.linspace(0, 2, 4, dtype=dtype, device=DEV)]
    x = torch.rand(5, 3, dtype=dtype, device=DEV)
    G = torch.randn(5, 7 * 9 * 4, dtype=dtype, device=DEV)
    xi = zo.meshgrid_points([t.cpu() for t in xg])
    for kern, ofun, ell0 in ((hk.SqExp(dtype=dtype), lambda a, b, s, e: zo.sqexp(a, b, s, e), torch.tensor([0.3, 0.5, 0.9], dtype=dtype)),
                             (hk.Matern(nu=0.5, dtype=dtype), lambda a, b, s, e: zo.matern(a, b, s, e, 0.5), torch.tensor(0.4, dtype=dtype)),
                             (hk.Matern(nu=1.5, dtype=dtype), lambda a, b, s, e: zo.matern(a, b, s, e, 1.5), torch.tensor(0.4, dtype=dtype)),
                             (hk.Matern(nu=2.5, dtype=dtype), lambda a, b, s, e: zo.matern(a, b, s, e, 2.5), torch.tensor(0.4, dtype=dtype))):
        s_d = torch.tensor(1.3, dtype=dtype, device=DEV, requires_grad=True); e_d = ell0.to(DEV).requires_grad_(True)
        K = kern.forward_grid(x, xg, (s_d, e_d))
        (K * G).sum().backward()
        s_c = torch.tensor(1.3, dtype=dtype, requires_grad=True); e_c = ell0.clone().requires_grad_(True)
        Kc = ofun(x.cpu(), xi, s_c, e_c)
        (Kc * G.cpu()).sum().backward()
        assert relerr(K, Kc.detach().numpy()) < 1e-12
        assert abs(float(s_d.grad) - float(s_c.grad)) <= 1e-10 * abs(float(s_c.grad))
        assert relerr(e_d.grad, e_c.grad.numpy()) < 1e-10
