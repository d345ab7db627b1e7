#!/usr/bin/env python
"""bench.py -- headline benchmark of the HIP-GP structured-kernel hot path on B200.

Workload (BASELINE.json configs[1]): 2-D inducing grid 1000 x 1000 (M = 10^6), Matern-5/2, ell = 0.01,
jitter 1e-3, fp32; one STEP = one PCG solve K_uu^-1 b with the HIP-GP preconditioner exactly as the SVI
loop issues it (maxiter 20, tol 1e-8 -- it does not converge, so 20 iterations = 20 K-matvecs + 21
preconditioner matvecs + the fused vector updates) on B = 16 right-hand sides per GPU.

metric  : Toeplitz matvec GB/s = algorithmic bytes of the step's 41 structured matvecs,
          w (2 M B + E_h) each (SURVEY.md 8d contract figure, E_h = 1998*1000), divided by the step time.
          `pcg_solve_s` (the other half of BASELINE.json's metric) is reported beside it.
value   : inputs resident in HBM (plan.pcg on device tensors).
e2e     : the same step through the C-ABI host entry point (hipgp_pcg_host): pinned host b -> H2D -> solve ->
          D2H x, every step.
roofline: per-kernel-class CUDA-event timing inside this script (a second pass of the same steps with the
          library's event hooks on); achieved = matvec algorithmic bytes / (sum of the three pass kernels'
          average durations); the dominant kernel and its share of the matvec are named.
cpu_baseline / --impl reference: the CPU oracle (oracle/ziggy_oracle.py, a restatement of the reference pinned
          to golden vectors from the unmodified reference) on the host cores, bounded sample B = 1.

N > 1 (torchrun): right-hand sides are independent, so each rank solves its own B = 16 shard with no data-path
collective ("weak" scaling); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = (1000, 1000)
ELL, SIG2, JITTER = 0.01, 1.0, 1e-3
MAXITER, TOL = 20, 1e-8
B_PER_GPU = 16
N_MATVEC = 2 * MAXITER + 1


def alg_bytes_matvec(B, w=4):
    M = GRID[0] * GRID[1]
    E_h = (2 * GRID[0] - 2) * ((2 * GRID[1] - 2) // 2 + 1)
    return w * (2 * M * B + E_h)


def first_row(dtype, device):
    """k(u_0, u_.) for Matern-5/2 on linspace(0,4,1000) x linspace(-2,2,1000) (+ jitter at [0])."""
    import torch
    from hipgp_b200 import kernels as hk
    g1 = torch.linspace(0, 4, GRID[0], dtype=dtype, device=device)
    g2 = torch.linspace(-2, 2, GRID[1], dtype=dtype, device=device)
    return hk.first_row([g1, g2], hk.Matern(nu=2.5, dtype=dtype), (SIG2, ELL), jitter=JITTER)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([t.strip() for t in line.split(",")])
        except Exception:
            pass

    def mark(self):
        return len(self.samples)

    def wait_running(self, timeout=5.0):
        t0 = time.time()
        while not self.samples and time.time() - t0 < timeout:
            time.sleep(0.01)

    def finish(self, lo=0, hi=None, lo_fallback=None, hi_fallback=None):
        """Samples [lo, hi) = the timed region; if that window holds fewer than 3 samples (a very short region) it is
        widened to [lo_fallback, hi_fallback) -- the same steps repeated under the same load -- and that is reported."""
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        window = "timed region"
        sel = self.samples[lo:hi]
        if len(sel) < 3 and lo_fallback is not None:
            sel = self.samples[lo_fallback:hi_fallback]
            window = "timed region + the e2e / roofline passes of the same step (timed region shorter than 3 sampling periods)"
        self.samples = sel
        self.window = window
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": getattr(self, "window", "")}


def host_threads():
    """threads for the CPU arm: the box's physical cores (what torch picks by default), within this process's affinity"""
    try:
        import psutil
        phys = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        phys = os.cpu_count() or 1
    try:
        phys = min(phys, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    return max(1, int(phys))


def oracle_step(B, threads=None):
    """One bounded CPU step: oracle PCG (20 iterations) on B right-hand sides of the 10^6-point grid."""
    import torch
    from oracle import ziggy_oracle as zo
    if threads:
        torch.set_num_threads(threads)
    g1 = torch.linspace(0, 4, GRID[0]); g2 = torch.linspace(-2, 2, GRID[1])
    kfun = lambda x, y: zo.matern(x, y, SIG2, ELL, 2.5)
    t0 = time.perf_counter()
    K = zo.OracleToeplitz([g1, g2], kfun, jitter_val=JITTER)
    t_setup = time.perf_counter() - t0
    torch.manual_seed(42)
    v = torch.randn(B, GRID[0] * GRID[1])

    def step():
        t = time.perf_counter()
        K.solve(v, do_precond=True, maxiter=MAXITER, tol=TOL)
        return time.perf_counter() - t
    return step, t_setup


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 1
    cores = host_threads()
    torch.set_num_threads(cores)               # torchrun exports OMP_NUM_THREADS=1: the CPU arm still uses every core
    step, t_setup = oracle_step(B, threads=cores)
    for _ in range(min(args.warmup, 1)):
        step()
    ts = [step() for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    val = N_MATVEC * alg_bytes_matvec(B) / t / 1e9
    line = {
        "impl": "reference", "metric": "toeplitz_matvec_GBps_in_pcg_1e6grid", "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "pcg_solve_s": t,
        "config": {"workload": "cfg2: 2D grid 1000x1000 (M=1e6), Matern-5/2 ell=0.01 jitter=1e-3, PCG maxiter=20 tol=1e-8 + HIP-GP preconditioner",
                   "rhs_per_step": B, "note": "CPU oracle port of the reference (torch CPU ops, all host threads); bounded sample B=1"},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": "1 rhs x 20 PCG iterations per step (full 10^6 grid); setup %.2fs not timed" % t_setup},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from hipgp_b200.plan import Plan
    from hipgp_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dtype = torch.float32
    B = B_PER_GPU
    M = GRID[0] * GRID[1]
    plan = Plan(GRID, dtype, dev)
    plan.set_first_row(first_row(dtype, dev))
    gen = torch.Generator(device=dev); gen.manual_seed(42 + rank)
    b = torch.randn(B, M, dtype=dtype, device=dev, generator=gen)
    b_host = b.cpu().pin_memory()
    x_host = torch.empty_like(b_host).pin_memory()

    def step_dev():
        return plan.pcg(b, maxiter=MAXITER, tol=TOL, precond=True)

    def step_e2e():
        return plan.pcg_host(b_host, x_host, maxiter=MAXITER, tol=TOL, precond=True)

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_running()
    for _ in range(max(args.warmup, 3)):
        step_dev()
    m0 = sampler.mark() if sampler else 0
    l0 = plan.launch_count()
    ms_total = timed(step_dev, args.steps)
    launches = plan.launch_count() - l0
    m1 = sampler.mark() if sampler else 0

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # roofline pass: same steps with per-kernel-class events on
    plan.profile(True)
    plan.profile_read(reset=True)
    for _ in range(args.steps):
        step_dev()
    prof = plan.profile_read(reset=True)
    plan.profile(False)
    m2 = sampler.mark() if sampler else 0
    clocks = sampler.finish(m0, m1, m0, m2) if sampler else None

    # fp64 and B=1 companions (not the headline; same step definition)
    extra = {}
    if world == 1 and not args.quick:          # single-GPU run only: `timed` synchronises all ranks
        b1 = b[:1].contiguous()
        for _ in range(3):
            plan.pcg(b1, maxiter=MAXITER, tol=TOL)
        extra["pcg_solve_s_B1_f32"] = timed(lambda: plan.pcg(b1, maxiter=MAXITER, tol=TOL), args.steps) / args.steps / 1e3
        for mode, nm in ((L.MV_K, "K"), (L.MV_CINV, "Cinv"), (L.MV_RT, "RT")):
            for _ in range(3):
                plan.matvec(mode, b)
            ms = timed(lambda: plan.matvec(mode, b), args.steps) / args.steps
            extra["matvec_ms_B16_f32_" + nm] = ms
        extra["matvec_GBps_B16_f32_K"] = alg_bytes_matvec(B) / extra["matvec_ms_B16_f32_K"] / 1e6
    if world > 1:
        dist.barrier()

    if rank == 0:
        t_step = ms_total / args.steps / 1e3
        alg_step = N_MATVEC * alg_bytes_matvec(B) * world
        value = alg_step / t_step / 1e9
        e2e_val = alg_step / (ms_e2e / args.steps / 1e3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        per = {k: (v[0] / v[1] if v[1] else 0.0) for k, v in prof.items()}
        mv_ms = per["rows_fwd"] + per["cols_pass"] + per["rows_inv"]
        dom = max(("rows_fwd", "cols_pass", "rows_inv"), key=lambda k: per[k])
        achieved = alg_bytes_matvec(B) / (mv_ms / 1e3) / 1e9 if mv_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_matvec_B16_f32")
            except Exception:
                traffic = None
        # bounded CPU sample on this box's host cores (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu:
            import torch as _t
            step, t_setup = oracle_step(1)
            step()
            tc = step()
            cpu = {"value": N_MATVEC * alg_bytes_matvec(1) / tc / 1e9, "unit": "GB/s", "cores": _t.get_num_threads(),
                   "kind": "port", "pcg_solve_s": tc,
                   "sample": "CPU oracle, 1 rhs x 20 PCG iterations on the full 10^6 grid (1 warm-up + 1 timed solve)"}
        # per-pass streaming bytes (what each pass moves when its input / output do not stay on chip),
        # averaged over the launches of one PCG iteration, against the same HBM peak
        Mv = 4 * GRID[0] * GRID[1] * B                                   # one vector, all right-hand sides
        Wb = 8 * GRID[0] * ((2 * GRID[1] - 2) // 2 + 1) * B              # half-spectrum workspace of the pruned rows
        spec_b = 4 * plan.embedding()[0][0] * ((2 * GRID[1] - 2) // 2 + 1)
        stream = {"rows_fwd": (3 * Mv + Wb + 6 * Mv + Wb) / 2.0,         # p = z + beta p | x += a p, r -= a Ap, r.r
                  "cols_pass": 2 * Wb + spec_b,
                  "rows_inv": Wb + 2 * Mv}                               # write result, read the dot operand
        per_kernel = {k: {"streaming_bytes": stream[k], "ms": per[k],
                          "achieved_GBps": stream[k] / (per[k] / 1e3) / 1e9 if per[k] > 0 else None,
                          "frac_of_hbm_peak": stream[k] / (per[k] / 1e3) / 1e9 / peak if per[k] > 0 else None} for k in stream}
        line = {
            "metric": "toeplitz_matvec_GBps_in_pcg_1e6grid", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "pcg_solve_s": t_step,
            "config": {"workload": "cfg2: 2D grid 1000x1000 (M=1e6), Matern-5/2 ell=0.01 jitter=1e-3, PCG maxiter=20 tol=1e-8 + HIP-GP preconditioner",
                       "rhs_per_gpu": B, "matvecs_per_step": N_MATVEC, "embedding": list(plan.embedding()[0]),
                       "l2": "working set per GPU (5 vectors x 64 MB + 135 MB of half-spectra) exceeds the 126 MB L2; no flush needed",
                       "parallelism": "rhs-sharded x%d, no data-path collective" % world},
            "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": int(b_host.numel() * 4) * world,
                    "d2h_bytes_per_step": int(x_host.numel() * 4) * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                         "kernel": dom, "kernel_share_of_matvec": per[dom] / mv_ms if mv_ms else None,
                         "per_launch_ms": per, "algorithmic_bytes_per_matvec": alg_bytes_matvec(B), "per_kernel_streaming": per_kernel,
                         "note": "achieved = contract bytes w(2MB+E_h) of one matvec / (rows_fwd + cols_pass + rows_inv average launch durations), CUDA events around every launch in a second pass of the same steps; the contract figure assumes the half-spectrum never leaves the chip -- per_kernel_streaming gives each pass against the bytes it actually has to stream; the column pass is shared-memory-bandwidth bound (DESIGN.md 4a: 79 % of that roofline)"},
            "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="skip the companion measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_gpu(args)


if __name__ == "__main__":
    main()
