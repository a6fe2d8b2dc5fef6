#!/usr/bin/env python
"""bench.py -- sparse-GP objective+gradient evaluations/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" is one evaluation of the headline path: Gaussian VI objective (elbo_fun) + gradient
(delbo_dcov_par) at fixed knots -- what one iteration of the reference's norm_grad_ascent_vi performs
(R/vi_functions.R:1089-1128) -- on BASELINE.json configs[4]: synthetic ARD, n = 1,000,000, d = 8, m = 1024
(SURVEY.md section 8d recipe).  The n rows are sharded over the ranks (strong scaling, total work fixed); each
pass ends in one NCCL sum-allreduce.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


if "reference" in sys.argv[1:] or "--impl=reference" in sys.argv[1:]:
    # The reference arm owns all host cores whatever the launcher exported: torch.distributed.run sets
    # OMP_NUM_THREADS=1 for its workers, which silently made the r01 arm single-threaded at N >= 2.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(host_cores())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sparse-GP obj+grad evals/sec (n=1M,m=1024,d=8 ARD)"
UNIT = "evals/s"


def workload_name(n, m, d):
    return "synthetic sparse ARD GP n=%d d=%d m=%d, VI objective+gradient (BASELINE configs[4])" % (n, d, m)


def workload(n, m, d, seed=1312):
    """X, U ~ N(0, I_d); sigma 1, l_c = 0.8 + 0.05 c, tau 0.5, delta 1e-6, mu = 0."""
    rng = np.random.default_rng(seed)
    x = np.asfortranarray(rng.normal(size=(n, d)))
    xu = np.asfortranarray(rng.normal(size=(m, d)))
    l = np.array([0.8 + 0.05 * (c + 1) for c in range(d)])
    y = np.sin(x[:, 0]) + 0.5 * x[:, 1] + 0.5 * rng.normal(size=n)
    return x, y, xu, dict(sigma=1.0, l=l, tau=0.5, delta=1e-6)


def shard_bounds(n, world, rank):
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak_tflops():
    """FP64 roofline denominator: cuBLAS DGEMM 8192^3 burst, measured in this run (MEASURED_PEAKS.json has no
    FP64 entry; SURVEY.md 8d asks for exactly this)."""
    import torch
    N = 8192
    a = torch.randn(N, N, dtype=torch.float64, device="cuda")
    b = torch.randn(N, N, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * N ** 3 / (best * 1e-3) / 1e12


def ncu_traffic(kernel):
    """DRAM bytes per launch (and tensor-pipe activity) of a kernel from the committed ncu --set full capture (profiles/)."""
    try:
        return dict(json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel])
    except Exception:
        return None


def cpu_reference_time(n_sample, m, d, threads=None):
    """Seconds for ONE obj+grad evaluation of the oracle (literal NumPy/OpenBLAS transcription of elbo_fun +
    delbo_dcov_par, with the per-element assembly in single-threaded C as Rcpp is) at n = n_sample."""
    from oracle import ref_model as rm
    x, y, xu, th = workload(n_sample, m, d, seed=1312)
    cp = {"sigma": th["sigma"]}
    for c in range(d):
        cp["l%d" % (c + 1)] = float(th["l"][c])
    cp["tau"] = th["tau"]
    t0 = time.perf_counter()
    obj, g = rm.vi_obj_grad(cp, "ard", xu, x, y, np.zeros(n_sample), th["delta"])
    return time.perf_counter() - t0, obj


def blas_threads(n=None):
    """Pin (n given) and report the BLAS thread count actually in use (threadpoolctl), not os.cpu_count()."""
    try:
        import threadpoolctl
        if n is not None:
            threadpoolctl.threadpool_limits(limits=int(n), user_api="blas")
        info = [p for p in threadpoolctl.threadpool_info() if p.get("user_api") == "blas"]
        if info:
            return int(info[0]["num_threads"]), "%s %s" % (info[0].get("internal_api", "blas"), info[0].get("version", "?"))
    except Exception:
        pass
    return None, "BLAS (threadpoolctl unavailable)"


def affine(n1, t1, n2, t2, n_full):
    """t(n) = a + b n through two samples: the reference's cost at fixed m is a fixed part (m x m factorisations and
    per-parameter m^3 products) plus a part linear in n, so a proportional extrapolation t(n2) n_full / n2 would charge
    the reference its fixed cost n_full / n2 times over."""
    b = max((t2 - t1) / (n2 - n1), 0.05 * t2 / n2)     # guard: timing noise must not produce a ~zero slope
    a = max(t2 - b * n2, 0.0)
    return a, b, a + b * n_full


def cpu_reference_estimate(n_sample, m, d, n_full):
    """Seconds per evaluation of the CPU port at n_full rows from TWO bounded samples (n_sample / 4 and n_sample rows)."""
    n1, n2 = max(256, n_sample // 4), n_sample
    t1, _ = cpu_reference_time(n1, m, d)
    t2, _ = cpu_reference_time(n2, m, d)
    a, b, sec = affine(n1, t1, n2, t2, n_full)
    return sec, ("%d rows %.2f s, %d rows %.2f s -> %.2f s + %.3g s/row, affine extrapolation to %d rows = %.0f s per "
                 "evaluation" % (n1, t1, n2, t2, a, b, n_full, sec))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores.  R is not installed on
    this image, so the timed code is the oracle port (kind "port"); n = 1e6 is out of reach for the literal
    algebra (>= 12 live 8.2 GB matrices, ~286 TFLOP), so the value is an affine EXTRAPOLATION a + b n:
      * once, before the steps: 16384 and 65536 rows on all cores, plus 512 and 2048 rows on ONE
        BLAS thread (the stand-in for R's default single-threaded reference BLAS);
      * every step: a bounded 2048-row sample; the step's value is the line through (2048, this step) and (65536, calibration)
        evaluated at n; the 16384-row calibration point is the linearity check (its residual from that line is printed).
    The thread count is pinned explicitly and reported from threadpoolctl, so the arm is the same at every --gpus N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, m, d = args.n, args.m, args.d
    cores = host_cores()
    threads, blas = blas_threads(cores)
    n_small, n_mid, n_big = args.ref_sample, 8 * args.ref_sample, 32 * args.ref_sample
    cpu_reference_time(1024, m, d)                                   # page in BLAS / build the C assembly
    t_mid, _ = cpu_reference_time(n_mid, m, d)
    t_big, _ = cpu_reference_time(n_big, m, d)
    one_thread = None
    if not args.no_one_thread:
        blas_threads(1)
        t1a, _ = cpu_reference_time(n_small // 4, m, d)
        t1b, _ = cpu_reference_time(n_small, m, d)
        a1, b1, sec1 = affine(n_small // 4, t1a, n_small, t1b, n)
        one_thread = {"value": 1.0 / sec1, "unit": UNIT, "cores": 1,
                      "sample": "%d rows %.2f s, %d rows %.2f s on 1 BLAS thread -> %.2f s + %.3g s/row = %.0f s per evaluation "
                                "(extrapolated)" % (n_small // 4, t1a, n_small, t1b, a1, b1, sec1)}
        threads, blas = blas_threads(cores)
    times, notes, resid = [], [], []
    for i in range(args.warmup + args.steps):
        t_small, _ = cpu_reference_time(n_small, m, d)
        a, b, sec = affine(n_small, t_small, n_big, t_big, n)
        if i >= args.warmup:
            times.append(sec)
            resid.append((t_mid - (a + b * n_mid)) / t_mid)
            notes.append("%d rows %.2f s (this step), %d rows %.2f s and %d rows %.2f s (calibration) -> %.2f s + %.3g s/row, "
                         "linearity residual at %d rows %+.1f %%, affine extrapolation to %d rows = %.0f s per evaluation"
                         % (n_small, t_small, n_mid, t_mid, n_big, t_big, a, b, n_mid, 100 * resid[-1], n, sec))
    sec_per_eval_full = float(np.mean(times))
    value = 1.0 / sec_per_eval_full
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_eval_full * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, m, d), "extrapolated": True},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads or cores, "kind": "port", "host_cores": cores,
                         "blas": blas, "extrapolated": True, "linearity_residual": float(np.mean(resid)),
                         "one_thread": one_thread,
                         "sample": "oracle (NumPy literal transcription of elbo_fun + delbo_dcov_par on %s with %s threads "
                                   "[threadpoolctl] + single-threaded C assembly as Rcpp is); last step: %s"
                                   % (blas, threads, notes[-1])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


GOLDEN = os.path.join(ROOT, "tests", "golden", "stated_sizes.json")


def load_golden():
    try:
        with open(GOLDEN) as f:
            return json.load(f)
    except Exception:
        return {}


def rel_errs(obj, grad, ref_obj, ref_grad):
    ref_grad = np.asarray(ref_grad, dtype=np.float64)
    grad = np.asarray(grad, dtype=np.float64)
    return {"obj_rel": float(abs(obj - ref_obj) / abs(ref_obj)),
            "grad_rel": float(np.max(np.abs(grad - ref_grad) / np.abs(ref_grad))),
            "grad_rel_to_max": float(np.max(np.abs(grad - ref_grad)) / np.max(np.abs(ref_grad)))}


def knot_errs(kg, summ):
    kg = np.asarray(kg, dtype=np.float64).reshape(summ["shape"])
    probe = np.array([kg[k, c] for k, c in summ["probe_idx"]])
    return float(max(abs(np.linalg.norm(kg) - summ["fro"]) / summ["fro"],
                     abs(np.abs(kg).sum() - summ["abs_sum"]) / summ["abs_sum"],
                     np.max(np.abs(probe - np.asarray(summ["probe"]))) / np.max(np.abs(summ["probe"]))))


def parity_checks(ctx, world, rank, golden):
    """Sharded (world ranks) evaluations of the secondary paths against the committed CPU-oracle goldens, before the
    timed region: FIC objective + gradient + knot gradient (n = 250k, m = 1024), the sparse-Laplace Newton search and
    gradient of BASELINE configs[3] (Bernoulli, n = 100k, m = 512).  Every multi-GPU driver run is thereby a parity
    check of the NCCL paths of gauss_fic.cu, laplace.cu and the knot-gradient allreduce."""
    from sparsergps_b200.vi_functions import knot_bounds
    out = {}
    if "cfg5_fic" in golden:
        g = golden["cfg5_fic"]
        x, y, xu, th = workload(g["n"], g["m"], g["d"])
        lo, hi = shard_bounds(g["n"], world, rank)
        ctx.set_data(np.asfortranarray(x[lo:hi]), np.ascontiguousarray(y[lo:hi]), None)
        ctx.timer_start()
        obj, grad, kg, _ = ctx.gauss_obj_grad_knots("fic", "ard", xu, th["sigma"], th["l"], th["tau"], th["delta"], knot_bounds(x))
        ms = ctx.timer_stop_ms()
        out["fic_n250k_m1024"] = dict(rel_errs(obj, grad, g["obj"], g["grad"]), knot_rel=knot_errs(kg, g["knot"]), ms=ms)
    if "cfg4" in golden:
        from sparsergps_b200 import laplace as Lp
        from tests import cases
        g = golden["cfg4"]
        c = cases.config4(n=g["n"], d=g["d"], m=g["m"])
        cp = c["cov_par"]
        lo, hi = shard_bounds(g["n"], world, rank)
        xs, ys = np.asfortranarray(c["x"][lo:hi]), np.ascontiguousarray(c["y"][lo:hi])
        t0 = time.perf_counter()
        fit = Lp.newtrap_sparseGP(np.zeros(hi - lo), "bernoulli", cp, "ard", xs, c["xu"], ys, np.zeros(hi - lo),
                                  np.zeros(g["m"]), maxit=g["maxit"], tol=g["tol"], delta=c["delta"], ctx=ctx)
        sec = time.perf_counter() - t0
        h = fit["objective_function_values"]
        ff = (1.5 * np.sin(c["x"][:, 0]) + c["x"][:, 1] - 0.5 * c["x"][:, 2])[lo:hi]        # tests/tools/make_golden_sizes.py
        t0 = time.perf_counter()
        got = Lp.dlogq_dcov_par(cp, "ard", c["xu"], xs, ys, ff, "bernoulli", np.zeros(hi - lo), c["delta"], ctx=ctx)["gradient"]
        gsec = time.perf_counter() - t0
        gerr = rel_errs(0.0, [got[k] for k in g["names"]], 1.0, g["grad_at_closed_form_ff"])
        out["laplace_n100k_m512"] = {
            "newton_iterations": int(len(h)), "newton_iterations_golden": g["iterations"],
            # the stopping test compares |grad_psi| with tol row by row, so the iteration count may move by a few with
            # the summation order of a different sharding; the histories must agree on their common prefix
            "hist_rel": float(np.max(np.abs(h[:min(len(h), len(g["hist"]))] - np.asarray(g["hist"])[:min(len(h), len(g["hist"]))])
                                     / np.abs(np.asarray(g["hist"])[:min(len(h), len(g["hist"]))]))),
            "u_mean_rel_to_max": float(np.max(np.abs(fit["u_posterior_mean"] - np.asarray(g["u_mean"]))) / np.max(np.abs(g["u_mean"]))),
            "grad_rel": gerr["grad_rel"], "newton_it_per_s": (len(h) - 1) / sec, "grad_ms": gsec * 1e3}
    return out


def secondary_timings(ctx, hbm_gbs):
    """The other kernels of SURVEY.md section 8 at the stated size, on one GPU, outside the headline timed region (CUDA events
    on the library's stream, median of 5 after 2 warm-ups): FIC objective + gradient (a18 / a20), and the materialising
    HBM-class kernels K1 make_cov_mat_ardC, K2 dsig_dtheta_ardC (one length-scale, tau) and K5 sum Omega o dK over 8.2 GB."""
    from sparsergps_b200 import _lib as L
    n, m, d = 1_000_000, 1024, 8
    x, y, xu, th = workload(n, m, d)
    out = {}
    ctx.set_data(x, y, None)
    ts = []
    for r in range(5):
        ctx.timer_start()
        ctx.gauss_obj_grad("fic", "ard", xu, th["sigma"], th["l"], th["tau"], th["delta"])
        ts.append(ctx.timer_stop_ms())
    out["fic_obj_grad_n1M_m1024"] = {"ms": float(np.median(ts[2:]))}
    xd, ud, big = ctx.dev_alloc(8 * n * d), ctx.dev_alloc(8 * m * d), ctx.dev_alloc(8 * n * m)
    ctx.fill_normal(xd, n * d, 1312)
    ctx.fill_normal(ud, m * d, 1313)
    l = np.ascontiguousarray(th["l"], dtype=np.float64)
    lib, h = ctx._lib, ctx.handle
    gb = 8.0 * n * m / 1e9
    res = np.zeros(d + 2)
    kernels = (
        ("K1_make_cov_mat_ardC", lambda: lib.srgp_make_cov_mat_dev(h, L.ARD, xd, n, ud, m, d, 1.0, L.ptr(l), 0.5, 1e-6, big), "write"),
        ("K2_dsig_dtheta_ardC_l3", lambda: lib.srgp_dsig_dtheta_dev(h, L.ARD, L.PAR_LC, 2, xd, n, ud, m, d, 1.0, L.ptr(l), 0.5, big), "write"),
        ("K2_dsig_dtheta_ardC_tau", lambda: lib.srgp_dsig_dtheta_dev(h, L.ARD, L.PAR_TAU, 0, xd, n, ud, m, d, 1.0, L.ptr(l), 0.5, big), "write"),
        ("K5_omega_dk_reduce", lambda: lib.srgp_omega_dk_reduce_dev(h, L.ARD, xd, n, ud, m, d, 1.0, L.ptr(l), 0.5, big, L.ptr(res)), "read"))
    for name, fn, rw in kernels:
        ts = []
        for r in range(7):
            ctx.timer_start()
            L.check(fn())
            ts.append(ctx.timer_stop_ms())
        ms = float(np.median(ts[2:]))
        out[name] = {"ms": ms, "GBps": gb / (ms * 1e-3), "frac_of_measured_hbm": gb / (ms * 1e-3) / hbm_gbs,
                     "algorithmic_bytes": "8 n m (%s once), n = 1e6, m = 1024" % rw}
    for p_ in (xd, ud, big):
        ctx.dev_free(p_)
    out["note"] = ("FP64-pipe bound, not HBM bound: 36 - 54 FP64 instructions per 8-byte entry against a ridge of 20; ncu counters in "
                   "profiles/r02_hbm_kernels_ncu.txt")
    return out


def measured_hbm_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=1024)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--ref-sample", type=int, default=2048, help="reference arm: rows of the per-step CPU sample (calibration: 8x and 32x)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-one-thread", action="store_true", help="reference arm: skip the 1-BLAS-thread figure")
    ap.add_argument("--no-check", action="store_true", help="skip the sharded parity checks before the timed region")
    ap.add_argument("--no-secondary", action="store_true", help="skip the FIC / K1 / K2 / K5 timings after the timed regions (1 GPU only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from sparsergps_b200.context import Context
    n, m, d = args.n, args.m, args.d
    x, y, xu, th = workload(n, m, d)
    lo, hi = shard_bounds(n, world, rank)
    xs, ys = np.asfortranarray(x[lo:hi]), np.ascontiguousarray(y[lo:hi])
    ctx = Context(local_rank)
    if world > 1:
        # NCCL unique id from rank 0 to everyone over the torch process group
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(Context.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        ctx.comm_init(world, rank, bytes(uid.cpu().tolist()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def eval_resident():
        return ctx.gauss_obj_grad("vi", "ard", xu, th["sigma"], th["l"], th["tau"], th["delta"])

    # pinned host copies of this rank's rows for the end-to-end arm
    xs_pin = torch.empty(xs.shape[::-1], dtype=torch.float64).pin_memory()     # (d, n) C-order == (n, d) F-order
    xs_pin.numpy()[...] = xs.T
    ys_pin = torch.empty(ys.shape, dtype=torch.float64).pin_memory()
    ys_pin.numpy()[...] = ys
    xs_host = xs_pin.numpy().T          # F-ordered view on pinned memory
    ys_host = ys_pin.numpy()

    def eval_e2e():
        return ctx.gauss_obj_grad_host("vi", "ard", xs_host, ys_host, None, xu, th["sigma"], th["l"], th["tau"],
                                       th["delta"])

    golden = load_golden()
    checks = None
    if not args.no_check and (n, m, d) == (1_000_000, 1024, 8):
        checks = parity_checks(ctx, world, rank, golden)

    # ---------------- device-resident arm (value) ----------------
    ctx.set_data(xs, ys, None)
    for _ in range(max(3, args.warmup)):
        ctx.flush_l2()
        obj, grad = eval_resident()
    sampler = ClockSampler(local_rank)
    ctx.prof_enable(True)
    ctx.prof_reset()
    launches0 = ctx.launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    for _ in range(args.steps):
        ctx.flush_l2()                  # L2 flush between timed iterations (256 MiB write, inside the timed region)
        obj, grad = eval_resident()
    ms = ctx.timer_stop_ms()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0
    prof = {k: ctx.prof_get(k) for k in ("gen", "gram", "km", "dense", "reduce", "comm")}
    ctx.prof_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---------------- end-to-end arm (host buffers, H2D + D2H inside the timed region) ----------------
    for _ in range(2):
        eval_e2e()
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        obj_e, grad_e = eval_e2e()
    ms_e = ctx.timer_stop_ms()
    barrier()
    t = torch.tensor([ms_e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e_max = float(t.item())

    parity = None
    g5 = golden.get("cfg5_vi")
    if g5 and (n, m, d) == (g5["n"], g5["m"], g5["d"]):
        parity = dict(rel_errs(obj, grad, g5["obj"], g5["grad"]), e2e=rel_errs(obj_e, grad_e, g5["obj"], g5["grad"]),
                      golden="tests/golden/stated_sizes.json cfg5_vi: oracle/reduced_model.vi_obj_grad on the CPU, 100 row shards, "
                             "float64 (tests/tools/make_golden_sizes.py)", tolerance=1e-8)
        if checks is not None:
            from sparsergps_b200.vi_functions import knot_bounds
            ctx.set_data(xs, ys, None)
            ctx.timer_start()
            _, _, kg, _ = ctx.gauss_obj_grad_knots("vi", "ard", xu, th["sigma"], th["l"], th["tau"], th["delta"], knot_bounds(x))
            checks["vi_knot_gradient_n1M_m1024"] = {"knot_rel": knot_errs(kg, g5["knot"]), "ms": ctx.timer_stop_ms()}

    if rank == 0:
        steps = args.steps
        value = steps / (ms_max * 1e-3)
        e2e_value = steps / (ms_e_max * 1e-3)
        nloc = hi - lo
        # Roofline of the dominant kernel.  Both row passes run on the INT8 tensor cores (tcgen05.mma.kind::i8, Ozaki
        # splitting: NS (NS + 1) / 2 exact INT8 slice-pair products per FP64 product), so the pipe
        # that bounds them is the INT8 tensor pipe: achieved = EXECUTED INT8 operations / kernel time against the INT8
        # peak measured in this run.  The algorithmic FP64 rate (2 n m^2 flop / kernel time) is given beside it against
        # the cuBLAS DGEMM rate measured in this run -- it exceeds 1 because no FP64 unit is used.
        peak64 = fp64_peak_tflops()
        km_launches, km_ms = prof["km"]
        gram_launches, gram_ms = prof["gram"]
        mp = (m + 127) // 128 * 128
        km_flops = 2.0 * nloc * m * m * steps
        gram_flops = 1.0 * nloc * m * (m + 1) * steps          # SYRK count: lower triangle incl. diagonal
        # INT8 peak: measured in this run by the resident-operand issue-rate probe of the library (csrc/probe.cu,
        # tcgen05.mma.kind::i8 128x128x32 on every SM; MEASURED_PEAKS.json has no INT8 entry).  It is a burst figure at
        # the clocks the probe ran at; the row-pass kernels run inside a long step at the lower clocks in `clocks`.
        slices = ctx.i8_slices()
        pairs = slices * (slices + 1) // 2
        peak8, cyc_mma = ctx.probe_i8_peak()
        peak8_src = ("measured in this run: csrc/probe.cu, resident-operand tcgen05.mma.kind::i8 128x128x32 on all SMs, "
                     "%.1f clk per MMA (64 = 8192 MAC/clk/SM), burst clocks" % cyc_mma)

        def rate(x, ms_):
            return x / (ms_ * 1e-3) / 1e12 if ms_ > 0 else None

        def frac(x, ms_, pk):
            return rate(x, ms_) / pk if ms_ > 0 else None

        sms = torch.cuda.get_device_properties(0).multi_processor_count
        rows2 = (sms // 2) * 128                                 # pass-2 chunk: (SMs / 2 column groups) row blocks x 128, rows are padded to it
        km_rows = -(-nloc // rows2) * rows2
        km_ops8 = pairs * 2.0 * km_rows * mp * mp * steps        # slice pairs x every 128-column block of every row block
        gram_quant = 128
        gram_rows = -(-nloc // gram_quant) * gram_quant
        gram_ops8 = pairs * 2.0 * gram_rows * (mp * (mp + 128) / 2.0) * steps   # 128 x 128 tiles of the lower block triangle
        ncu = ncu_traffic("i8_km2_kernel") or {}
        roof = {"bound": "tensor", "kernel": "i8_km2_kernel (K*Mop^T on tcgen05 kind::i8, 128 x 128 tiles in two sweeps, %d x %d digit "
                "slices = %d exact INT8 products per FP64 product, fused dK reductions)" % (slices, slices, pairs),
                "achieved": rate(km_ops8, km_ms), "peak": peak8, "unit": "TOP/s (INT8, executed)", "frac": frac(km_ops8, km_ms, peak8),
                "peak_source": peak8_src,
                "fp64_equivalent": {"achieved": rate(km_flops, km_ms), "unit": "TFLOP/s", "algorithmic_flops": "2 n m^2",
                                    "vs_cublas_dgemm": frac(km_flops, km_ms, peak64), "cublas_dgemm_tflops": peak64},
                "tensor_pipe_active_ncu": ncu.get("tensor_pipe_active"),
                "traffic": ncu.get("bytes_per_launch"), "traffic_detail": ncu or None}
        roof_gram = {"kernel": "i8_gram2_kernel (K^T K on tcgen05 kind::i8, 128 x 128 tiles of the lower block triangle)",
                     "achieved": rate(gram_ops8, gram_ms), "peak": peak8, "unit": "TOP/s (INT8, executed)",
                     "frac": frac(gram_ops8, gram_ms, peak8),
                     "tensor_pipe_active_ncu": (ncu_traffic("i8_gram2_kernel") or {}).get("tensor_pipe_active"),
                     "fp64_equivalent": {"achieved": rate(gram_flops, gram_ms), "unit": "TFLOP/s",
                                         "algorithmic_flops": "n m (m+1)", "vs_cublas_dgemm": frac(gram_flops, gram_ms, peak64)}}
        roof.update({"launches_per_step": km_launches / steps, "avg_launch_ms": km_ms / max(1, km_launches),
                     "share_of_step": km_ms / ms_max})
        roof_gram.update({"launches_per_step": gram_launches / steps, "share_of_step": gram_ms / ms_max})
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            threads, blas = blas_threads(host_cores())
            tsec, note = cpu_reference_estimate(16384, m, d, n)
            cpu = {"value": 1.0 / tsec, "unit": UNIT, "cores": threads or host_cores(), "kind": "port", "extrapolated": True,
                   "sample": "oracle (NumPy literal transcription on %s, %s threads [threadpoolctl] + single-threaded C assembly): %s"
                             % (blas, threads, note)}
        secondary = None
        if world == 1 and not args.no_secondary and (n, m, d) == (1_000_000, 1024, 8):
            hbm, hbm_src = measured_hbm_gbs()
            secondary = secondary_timings(ctx, hbm)
            secondary["hbm_peak_gbs"] = hbm
            secondary["hbm_peak_source"] = hbm_src
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, m, d), "rows_per_gpu": nloc, "l2": "flushed between timed iterations "
                       "(256 MiB write inside the timed region)", "parallelism": "rows sharded x%d, 2 NCCL allreduces/eval" % world},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(xs.nbytes + ys.nbytes + xu.nbytes + 8 * (d + 3)),
                    "d2h_bytes_per_step": int(8 * 256), "ms_per_step": ms_e_max / steps},
            "roofline": roof, "roofline_gram": roof_gram, "cpu_baseline": cpu,
            "kernel_ms_per_step": {k: v[1] / steps for k, v in prof.items()},
            "objective": obj, "grad_norm": float(np.linalg.norm(grad)),
            "parity": parity, "parity_checks": checks, "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
