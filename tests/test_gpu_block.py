"""GPU parity of the block-diagonal variational family (ziggy/hipgp.py:527-690) against the unmodified reference
(tests/golden/block_step_*.npz): fused get_lam / block_diag_multiply kernels, compute_knSkn, KL, the natural-gradient
step of elbo_and_grad (block branch, hipgp.py:251-261) and predict; plus a larger randomised case against torch indexing
(host-side checker only).  Index maps are bit-exact; tolerances 1e-5 relative in fp32, 1e-10 in fp64 for the kernels."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DT = {"f32": torch.float32, "f64": torch.float64}
TOL = {"f32": 1e-5, "f64": 1e-10}


def relerr(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    a = a.astype(np.float64); b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def make_model(g, dtype):
    from hipgp_b200 import kernels as hk
    from hipgp_b200.hipgp import BlockToeplitzGP
    xgrids = [torch.linspace(lo, hi, int(m), dtype=dtype) for lo, hi, m in g["grids"]]
    sig2, ell, jit, nobs = [float(v) for v in g["params"]]
    mod = BlockToeplitzGP(hk.Matern(nu=1.5, dtype=dtype), xgrids, num_obs=int(nobs), block_sizes=[int(b) for b in g["block_sizes"]],
                          sig2_init=sig2, ell_init=ell, dtype=dtype, jitter_val=jit).cuda_params(0)
    mod.global_theta1.data = torch.from_numpy(g["theta1"]).to(DEV)
    mod.global_theta2.data = torch.from_numpy(g["theta2"]).to(DEV)
    return mod


@pytest.mark.parametrize("dname", ["f32", "f64"])
def test_block_step_vs_reference(dname, golden_dir):
    g = np.load(os.path.join(golden_dir, "block_step_%s.npz" % dname))
    dtype = DT[dname]
    mod = make_model(g, dtype)
    assert mod.name == 'block' and (mod.num_blocks, mod.block_size) == g["block_idx"].shape
    assert np.array_equal(mod.block_idx.numpy(), g["block_idx"])                  # bit-exact index map
    kn = torch.from_numpy(g["kn"]).to(DEV); qS = torch.from_numpy(g["qS"]).to(DEV)
    nb = torch.from_numpy(g["noise_std"]).to(DEV)
    tol = TOL[dname]
    assert relerr(mod.get_lam(1 / nb ** 2, kn, bscale=500 / 8), g["lam"]) < tol
    assert relerr(mod.block_diag_multiply(qS, kn), g["Sv"]) < tol
    assert relerr(mod.compute_knSkn(kn, qS), g["knSkn"]) < tol
    qm, qS2 = mod.standard_variational_params()
    lim = 1e-8 if dname == "f64" else 2e-3           # through the PCG solves / small inverses, as for the mean-field step
    assert relerr(qm, g["qm"]) < lim and relerr(qS2, g["qS"]) < lim
    assert abs(float(mod.get_kl_to_prior(qm, qS2)) - float(g["kl"])) <= lim * abs(float(g["kl"]))
    x = torch.from_numpy(g["x"]).to(DEV); y = torch.from_numpy(g["y"]).to(DEV)
    elbo = mod.elbo_and_grad(x, y, nb, maxiter_cg=20)
    assert abs(float(elbo) - float(g["elbo"])) <= lim * abs(float(g["elbo"]))
    assert relerr(mod.global_theta1.grad, g["g1"]) < lim and relerr(mod.global_theta2.grad, g["g2"]) < lim
    mu, sig = mod.predict(x, maxiter_cg=50)
    assert relerr(mu, g["mu"]) < lim and relerr(sig, g["sig"]) < lim


@pytest.mark.parametrize("dname", ["f32", "f64"])
def test_block_kernels_larger_random(dname):
    """300x300-grid shapes (M' = 598^2, 13x13 chunks -> 169-point blocks, 2116 blocks), 40 rows (several row chunks in
    fp64), and a 3-D map with 1000-point blocks; checker = torch gather + bmm in fp64."""
    from hipgp_b200.util import define_block_chunks
    from hipgp_b200.plan import block_lam, block_diag_multiply
    dtype = DT[dname]
    gen = torch.Generator(device=DEV); gen.manual_seed(3)
    for lens, chunks, B in (((598, 598), (13, 13), 40), ((20, 30, 20), (10, 10, 10), 70), ((26, 20), (13, 5), 1)):
        idx, to_b, from_b = define_block_chunks([torch.arange(n) for n in lens], list(chunks))
        idxd = idx.to(DEV)
        E = int(np.prod(lens)); nblk, bs = idx.shape
        kn = torch.randn(B, E, device=DEV, dtype=dtype, generator=gen)
        w = torch.rand(B, device=DEV, dtype=dtype, generator=gen) + 0.5
        lam = block_lam(kn, w, idx, scale=2.5, diag=1.0)
        kb = kn.double()[:, idxd].transpose(0, 1)                                # (nblk, B, bs)
        want = 2.5 * torch.matmul(kb.transpose(1, 2), w.double()[None, :, None] * kb) + torch.eye(bs, device=DEV, dtype=torch.float64)
        assert relerr(lam, want.cpu().numpy()) < TOL[dname]
        S = torch.randn(nblk, bs, bs, device=DEV, dtype=dtype, generator=gen)
        out = block_diag_multiply(S, kn, idx)
        wantv = torch.matmul(S.double(), kn.double()[:, idxd][..., None]).flatten(start_dim=1)[:, torch.argsort(idxd.flatten())]
        assert relerr(out, wantv.cpu().numpy()) < TOL[dname]
