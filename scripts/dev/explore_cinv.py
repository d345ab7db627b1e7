"""fp32 preconditioner error against the fp64 truth next to kappa * eps (developer tool; bounds of the parity tests)."""
import sys, os, glob
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from hipgp_b200.plan import Plan
from hipgp_b200 import _lib as L
import test_gpu_sizes as tz
def relerr(a, b): return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / np.linalg.norm(np.asarray(b, np.float64))
eps = 2.0 ** -24
for path in sorted(glob.glob(os.path.join(ROOT, "tests/golden/toeplitz_*_f32.npz"))):
    g = np.load(path, allow_pickle=True); g64 = np.load(path.replace("_f32", "_f64"), allow_pickle=True)
    dims = [int(x[2]) for x in g["grids"]]
    plan = Plan(dims, torch.float32, "cuda:0").set_first_row(torch.from_numpy(g["column"]).cuda())
    D = plan.spectrum(L.SPEC_D); kappa = float(D.max() / D.min())
    got = plan.matvec(L.MV_CINV, torch.from_numpy(g["v"]).cuda()).cpu().numpy()
    print(os.path.basename(path)[9:-8], "kappa %.3g kappa*eps %.3g | dev vs truth64 %.3g | ref32 vs truth64 %.3g | dev vs ref32 %.3g" %
          (kappa, kappa * eps, relerr(got, g64["Cinv_v"]), relerr(g["Cinv_v"], g64["Cinv_v"]), relerr(got, g["Cinv_v"])), flush=True)
for dims, B in tz.CASES:
    col = tz.first_col(dims); M = int(np.prod(dims))
    v64 = np.random.default_rng(1).standard_normal((B, M))
    plan = Plan(list(dims), torch.float32, "cuda:0").set_first_row(torch.from_numpy(col).to("cuda:0", torch.float32))
    D = plan.spectrum(L.SPEC_D); kappa = float(D.max() / D.min())
    v = torch.from_numpy(v64).to("cuda:0", torch.float32)
    Pv = plan.matvec(L.MV_CINV, v).cpu().numpy()
    print(dims, "kappa %.3g kappa*eps %.3g | dev32 vs dense64 %.3g" % (kappa, kappa * eps, relerr(Pv, tz.dense_apply(col, dims, v.double().cpu().numpy(), lambda d: 1.0 / d))), flush=True)
