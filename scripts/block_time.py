"""Times the block-diagonal family's kernels at the BASELINE config 3 shape (300x300 grid -> M' = 598^2, batch 200) and
one full BlockToeplitzGP natural-gradient step beside the mean-field one.  Prints JSON lines."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hipgp_b200 import kernels as hk  # noqa: E402
from hipgp_b200.hipgp import BlockToeplitzGP, MeanFieldToeplitzGP  # noqa: E402
from hipgp_b200.plan import block_lam, block_diag_multiply  # noqa: E402
from hipgp_b200.util import define_block_chunks  # noqa: E402

DEV = "cuda:0"


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dt = torch.float32
    B = 200
    for chunks in ((13, 13), (23, 23), (26, 26)):
        idx, _, _ = define_block_chunks([torch.arange(598), torch.arange(598)], list(chunks))
        idxd = idx.to(DEV)
        nblk, bs = idx.shape
        kn = torch.randn(B, 598 * 598, device=DEV, dtype=dt); w = torch.rand(B, device=DEV, dtype=dt)
        S = torch.randn(nblk, bs, bs, device=DEV, dtype=dt)
        t_lam = timeit(lambda: block_lam(kn, w, idxd, 2.0, 1.0))
        t_mul = timeit(lambda: block_diag_multiply(S, kn, idxd))
        flops = 2.0 * nblk * bs * bs * B
        print(json.dumps({"block": list(chunks), "num_blocks": nblk, "block_size": bs, "batch": B,
                          "block_lam_ms": t_lam, "block_lam_TFLOPs": flops / t_lam / 1e9,
                          "block_diag_multiply_ms": t_mul, "block_diag_multiply_TFLOPs": flops / t_mul / 1e9}), flush=True)
    xgrids = [torch.linspace(0, 1, 300, dtype=dt), torch.linspace(0, 1, 300, dtype=dt)]
    x = torch.rand(B, 2, device=DEV, dtype=dt); y = torch.randn(B, 1, device=DEV, dtype=dt)
    for name, mod in (("mean-field", MeanFieldToeplitzGP(hk.Matern(nu=1.5, dtype=dt), xgrids, num_obs=10 ** 6, ell_init=0.02, dtype=dt)),
                      ("block 13x13", BlockToeplitzGP(hk.Matern(nu=1.5, dtype=dt), xgrids, num_obs=10 ** 6, block_sizes=[13, 13],
                                                      ell_init=0.02, dtype=dt))):
        mod = mod.cuda_params(0)
        t = timeit(lambda: mod.elbo_and_grad(x, y, maxiter_cg=20), n=5, warm=2)
        print(json.dumps({"family": name, "svi_step_ms": t, "grid": [300, 300], "batch": B}), flush=True)


if __name__ == "__main__":
    main()
