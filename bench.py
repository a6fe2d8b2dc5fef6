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
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]
        return {"bytes_per_launch": t["bytes_per_launch"], "algorithmic_bytes_per_launch": t["algorithmic_bytes_per_launch"],
                "source": t["source"]}
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


def cpu_reference_estimate(n_sample, m, d, n_full):
    """Seconds per evaluation of the CPU port at n_full rows from TWO bounded samples (n_sample / 4 and n_sample rows):
    the reference's cost is a + b n at fixed m (m x m factorisations and per-parameter m^3 products do not depend on n,
    every other term is linear in n), so the estimate is the affine fit, not t(n_sample) * n_full / n_sample, which
    would charge the reference n_full / n_sample times its fixed cost."""
    n1, n2 = max(256, n_sample // 4), n_sample
    t1, _ = cpu_reference_time(n1, m, d)
    t2, _ = cpu_reference_time(n2, m, d)
    b = max((t2 - t1) / (n2 - n1), 0.05 * t2 / n2)     # guard: timing noise must not produce a ~zero slope
    a = max(t2 - b * n2, 0.0)
    sec = a + b * n_full
    return sec, ("%d rows %.2f s, %d rows %.2f s -> %.2f s + %.3g s/row, affine extrapolation to %d rows = %.0f s per "
                 "evaluation" % (n1, t1, n2, t2, a, b, n_full, sec))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  R is not installed on this image
    (SURVEY.md 8c), so the timed code is the oracle port; n = 1e6 is out of reach for the literal algebra
    (>= 12 live 8.2 GB matrices, ~286 TFLOP), so each step evaluates two bounded row samples and the value is
    the affine extrapolation a + b n (see cpu_reference_estimate)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, m, d = args.n, args.m, args.d
    n_sample = args.ref_sample
    cores = os.cpu_count() or 1
    times, notes = [], []
    for i in range(args.warmup + args.steps):
        t, note = cpu_reference_estimate(n_sample, m, d, n)
        if i >= args.warmup:
            times.append(t)
            notes.append(note)
    sec_per_eval_full = float(np.mean(times))
    value = 1.0 / sec_per_eval_full
    blas = "OpenBLAS (numpy scipy-openblas), %d threads" % cores
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_eval_full * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, m, d)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "oracle (NumPy literal transcription + single-threaded C assembly, %s); each step "
                                   "times two row samples; last step: %s" % (blas, notes[-1])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=1024)
    ap.add_argument("--d", type=int, default=8)
    ap.add_argument("--ref-sample", type=int, default=4096, help="rows per CPU-baseline evaluation")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    if rank == 0:
        steps = args.steps
        value = steps / (ms_max * 1e-3)
        e2e_value = steps / (ms_e_max * 1e-3)
        nloc = hi - lo
        # Roofline of the dominant kernel.  Both row passes run on the INT8 tensor cores (tcgen05.mma.kind::i8, Ozaki
        # splitting: 36 exact INT8 slice-pair products per FP64 product), so the pipe that bounds
        # them is the INT8 tensor pipe: achieved = EXECUTED INT8 operations / kernel time, peak = 2 x the measured dense
        # bf16 rate of MEASURED_PEAKS.json (sustained figure: the kernels are timed inside a long step; INT8 : bf16 is
        # 2 : 1 on B200, 4.5 vs 2.25 POP/s nominal).  The algorithmic FP64 rate (2 n m^2 flop / kernel time) is given
        # beside it against the cuBLAS DGEMM rate measured in this run -- it exceeds 1 because no FP64 unit is used.
        peak64 = fp64_peak_tflops()
        km_launches, km_ms = prof["km"]
        gram_launches, gram_ms = prof["gram"]
        mp = (m + 127) // 128 * 128
        km_flops = 2.0 * nloc * m * m * steps
        gram_flops = 1.0 * nloc * m * (m + 1) * steps          # SYRK count: lower triangle incl. diagonal
        dmma = os.environ.get("SRGP_TENSOR", "").lower().startswith("d")
        try:
            mpk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak8, peak8_src = 2.0 * float(mpk["bf16_tflops_sustained"]), "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (measured)"
        except Exception:
            peak8, peak8_src = 2.0 * 1400.0, "2 x 1.4 PFLOP/s sustained bf16 (fallback of B200_PROFILING.md)"

        def rate(x, ms_):
            return x / (ms_ * 1e-3) / 1e12 if ms_ > 0 else None

        def frac(x, ms_, pk):
            return rate(x, ms_) / pk if ms_ > 0 else None

        sms = torch.cuda.get_device_properties(0).multi_processor_count
        rows2 = (sms // 2) * 128                                 # pass-2 chunk: (SMs / 2 column groups) row blocks x 128, rows are padded to it
        km_rows = -(-nloc // rows2) * rows2
        km_ops8 = 36 * 2.0 * km_rows * mp * mp * steps           # 36 slice pairs, every 64-column tile of every row block
        gram_quant = 128
        gram_rows = -(-nloc // gram_quant) * gram_quant
        gram_ops8 = 36 * 2.0 * gram_rows * (mp * (mp + 128) / 2.0) * steps   # 128 x 64 tiles of the lower block triangle
        if dmma:
            roof = {"bound": "tensor", "kernel": "km_reduce_kernel (K*M on DMMA + fused dK reductions)",
                    "achieved": rate(km_flops, km_ms), "peak": peak64, "unit": "TFLOP/s", "frac": frac(km_flops, km_ms, peak64),
                    "traffic": (ncu_traffic("km_reduce_kernel") or {}).get("bytes_per_launch"),
                    "traffic_detail": ncu_traffic("km_reduce_kernel"),
                    "peak_source": "cuBLAS DGEMM 8192^3 burst measured in this run (no FP64 entry in MEASURED_PEAKS.json)"}
            roof_gram = {"kernel": "syrk_chunk_kernel (K^T K on DMMA, SYRK flop count n m (m+1))",
                         "achieved": rate(gram_flops, gram_ms), "peak": peak64, "unit": "TFLOP/s",
                         "frac": frac(gram_flops, gram_ms, peak64)}
        else:
            roof = {"bound": "tensor", "kernel": "i8_km_kernel (K*Mop^T on tcgen05 kind::i8, 8 x 8 digit slices, fused dK reductions)",
                    "achieved": rate(km_ops8, km_ms), "peak": peak8, "unit": "TOP/s (INT8, executed)", "frac": frac(km_ops8, km_ms, peak8),
                    "peak_source": peak8_src,
                    "fp64_equivalent": {"achieved": rate(km_flops, km_ms), "unit": "TFLOP/s", "algorithmic_flops": "2 n m^2",
                                        "vs_cublas_dgemm": frac(km_flops, km_ms, peak64), "cublas_dgemm_tflops": peak64},
                    "traffic": (ncu_traffic("i8_km_kernel") or {}).get("bytes_per_launch"),
                    "traffic_detail": ncu_traffic("i8_km_kernel")}
            roof_gram = {"kernel": "i8_gram_kernel (K^T K on tcgen05 kind::i8, lower block triangle)",
                         "achieved": rate(gram_ops8, gram_ms), "peak": peak8, "unit": "TOP/s (INT8, executed)",
                         "frac": frac(gram_ops8, gram_ms, peak8),
                         "fp64_equivalent": {"achieved": rate(gram_flops, gram_ms), "unit": "TFLOP/s",
                                             "algorithmic_flops": "n m (m+1)", "vs_cublas_dgemm": frac(gram_flops, gram_ms, peak64)}}
        roof.update({"launches_per_step": km_launches / steps, "avg_launch_ms": km_ms / max(1, km_launches),
                     "share_of_step": km_ms / ms_max})
        roof_gram.update({"launches_per_step": gram_launches / steps, "share_of_step": gram_ms / ms_max})
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            tsec, note = cpu_reference_estimate(args.ref_sample, m, d, n)
            cpu = {"value": 1.0 / tsec, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "oracle (NumPy/OpenBLAS %d threads + single-threaded C assembly): %s" % (cores, note)}
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
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
