"""Developer check: run the CPU emulation build of the kernel sources on grids that hit given embedding lengths and
compare K / C^-1 / R^T / R matvecs and a short PCG with a dense numpy FFT evaluation.  (Checker only.)"""
import ctypes as C, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hipgp_b200 import _lib as L
import emu_build
lib = emu_build.load()

def ptr(a): return a.ctypes.data_as(C.c_void_p)

def ref_ops(col, m):
    # reference semantics: circulant embedding of size N = 2m-2 per axis, D = max(Re FFT C, 1e-6)
    D = len(m)
    Cc = col.reshape(m).astype(np.float64)
    for d in range(D):
        if m[d] > 1:
            sl = [slice(None)] * D; sl[d] = slice(m[d] - 2, 0, -1)
            Cc = np.concatenate([Cc, Cc[tuple(sl)]], axis=d)
    Dg = np.maximum(np.fft.fftn(Cc).real, 1e-6)
    N = Cc.shape
    def apply(v, f, pad_in=True, crop=True):
        B = v.shape[0]
        if pad_in:
            x = np.zeros((B,) + N); x[(slice(None),) + tuple(slice(0, k) for k in m)] = v.reshape((B,) + tuple(m))
        else:
            x = v.reshape((B,) + N)
        y = np.fft.ifftn(f * np.fft.fftn(x, axes=range(1, D + 1)), axes=range(1, D + 1)).real
        if crop: y = y[(slice(None),) + tuple(slice(0, k) for k in m)]
        return y.reshape(B, -1)
    return Dg, apply

def run(m, dt, B=2, do_wide=True):
    m = list(m)
    rng = np.random.default_rng(0)
    g = np.meshgrid(*[np.linspace(0, 1, k) for k in m], indexing="ij")
    r = np.sqrt(sum((x - x.flat[0]) ** 2 for x in g))
    col = (np.exp(-r / 0.2) * (1 + r / 0.2)).reshape(-1); col[0] += 1e-2
    col = col.astype(dt)
    plan = C.c_void_p()
    mm = np.array(m, dtype=np.int64)
    assert lib.hipgp_plan_create(len(m), mm.ctypes.data_as(L._pi64), L.F32 if dt == np.float32 else L.F64, 0, C.byref(plan)) == 0
    ncl = C.c_int64()
    assert lib.hipgp_plan_set_first_row(plan, ptr(col), 1e-6, C.byref(ncl), None) == 0, lib.hipgp_last_error()
    Ln = np.zeros(len(m), dtype=np.int64); Lw = np.zeros(len(m), dtype=np.int64)
    lib.hipgp_plan_embedding(plan, Ln.ctypes.data_as(L._pi64), Lw.ctypes.data_as(L._pi64))
    Dg, apply = ref_ops(col.astype(np.float64), m)
    M = int(np.prod(m)); E = int(np.prod([2 * k - 2 if k > 1 else 1 for k in m]))
    v = rng.standard_normal((B, M)).astype(dt); w = rng.standard_normal((B, E)).astype(dt)
    tol = 2e-5 if dt == np.float32 else 1e-10
    res = {}
    modes = [(0, "K", v, M, lambda: apply(v.astype(np.float64), Dg)), (1, "Cinv", v, M, lambda: apply(v.astype(np.float64), 1 / Dg))]
    if do_wide:
        modes += [(2, "RT", v, E, lambda: apply(v.astype(np.float64), np.sqrt(Dg), crop=False)),
                  (3, "R", w, M, lambda: apply(w.astype(np.float64), np.sqrt(Dg), pad_in=False))]
    for mode, name, inp, osz, ref in modes:
        out = np.zeros((B, osz), dtype=dt)
        t0 = time.time()
        assert lib.hipgp_matvec(plan, mode, ptr(inp), ptr(out), B, None) == 0, lib.hipgp_last_error()
        rr = ref()
        res[name] = np.linalg.norm(out - rr) / np.linalg.norm(rr)
    # short PCG: compare with a numpy PCG on the same operators
    x = np.zeros((B, M), dtype=dt); it = C.c_int(); cb = C.c_int()
    assert lib.hipgp_pcg(plan, ptr(v), ptr(x), B, 5, 1e-12, 1, C.byref(it), C.byref(cb), None, L.ITER_CB(0), None, None) == 0, lib.hipgp_last_error()
    b64 = v.astype(np.float64); xr = np.zeros_like(b64); rres = b64.copy(); z = apply(rres, 1 / Dg); p = z.copy()
    for _ in range(5):
        rs = (rres * z).sum(1, keepdims=True); Ap = apply(p, Dg); al = rs / (p * Ap).sum(1, keepdims=True)
        xr += al * p; rres -= al * Ap; z = apply(rres, 1 / Dg); p = z + ((z * rres).sum(1, keepdims=True) / rs) * p
    res["pcg5"] = np.linalg.norm(x - xr) / np.linalg.norm(xr)
    lib.hipgp_plan_destroy(plan)
    ok = all(e < (tol * (50 if k in ("pcg5", "Cinv") and dt == np.float32 else 1)) for k, e in res.items())
    print(m, dt.__name__, "Ln", list(Ln), "Lw", list(Lw), {k: "%.1e" % e for k, e in res.items()}, "OK" if ok else "FAIL", flush=True)
    return ok

if __name__ == "__main__":
    cases = eval(sys.argv[1]) if len(sys.argv) > 1 else [((9, 12), True), ((30, 60), True), ((5, 7, 9), True)]
    allok = True
    for m, wide in cases:
        for dt in (np.float32, np.float64):
            allok &= run(m, dt, do_wide=wide)
    sys.exit(0 if allok else 1)
