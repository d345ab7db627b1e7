"""Times hipgp_toeplitz_quadform (the Toeplitz-column gradient of InvMatmul.backward) at BASELINE config 2 / 4 shapes and,
beside it, the oracle port of the reference's route (gpt_toeplitz.py:169-209: two 1-D FFT Toeplitz products of length
2M-1 per pair) for ONE pair on the host.  Prints one JSON line per shape."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hipgp_b200.plan import Plan  # noqa: E402

DEV = "cuda:0"


def main():
    from oracle import ziggy_oracle as zo      # checker / CPU baseline only
    for dims, S in (((1000, 1000), 16), ((128, 128, 64), 16), ((300, 300), 16)):
        M = int(np.prod(dims))
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            col = torch.zeros(M, device=DEV, dtype=dt); col[0] = 1.0
            plan = Plan(list(dims), dt, DEV).set_first_row(col)
            g = torch.Generator(device=DEV); g.manual_seed(0)
            u = torch.randn(S, M, device=DEV, dtype=dt, generator=g); v = torch.randn(S, M, device=DEV, dtype=dt, generator=g)
            for _ in range(3):
                out = plan.toeplitz_quadform(u, v)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n):
                out = plan.toeplitz_quadform(u, v)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            rec = {"dims": list(dims), "pairs": S, "dtype": name, "gpu_ms": ms, "gpu_ms_per_pair": ms / S}
            if name == "f64":
                u1 = u[0].cpu().numpy(); v1 = v[0].cpu().numpy()
                t0 = time.perf_counter()
                ref = zo.sym_toeplitz_derivative_quadratic_form(u1, v1)
                rec["cpu_port_s_per_pair"] = time.perf_counter() - t0
                one = plan.toeplitz_quadform(u[:1], v[:1]).cpu().numpy()
                rec["relerr_vs_port"] = float(np.linalg.norm(one - ref) / np.linalg.norm(ref))
            print(json.dumps(rec), flush=True)
            del plan


if __name__ == "__main__":
    main()
