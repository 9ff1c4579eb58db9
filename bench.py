#!/usr/bin/env python
"""
Benchmark of the hot path: predict_y + UCB + arg-max over leaf candidates (BASELINE.json metric, config C3:
N=4096 training points, d=10, 1e7 candidates, Matern-5/2, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one full scoring pass over the candidate set (a `gp_eval_best_ucb` call at C3 size).  With N > 1 GPUs
the candidates are sharded contiguously over the ranks (total work fixed -> "scaling": "strong"); the fit state is
broadcast once with NCCL before the timed region (reported as broadcast_ms), and every step ends with the all-gather
of one 32-byte record per rank.  Prints ONE JSON line on rank 0 (keys described in DESIGN.md section "Measurement").

--impl reference times the CPU restatement of the reference's own path (oracle/gpr_oracle.py, GPflow op order incl.
the Cholesky inside every predict_y call) on the box's host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
from scipy.special import erfcinv

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240517
WORKLOADS = {
    # name: (N, d, M, description)
    "c3": (4096, 10, 10_000_000, "C3: UCB scoring, N=4096 train, d=10, 1e7 leaf candidates, Matern-5/2 fp64"),
    "c2": (512, 2, 100_000, "C2: predict_y+UCB microbench, N=512 train, d=2, 1e5 candidates, Matern-5/2 fp64"),
}
FP64_PEAK_TFLOPS = 37.03  # measured DMMA.8x8x4 issue peak of this pool's B200 (profiles/r01_fp64_probe.txt)
INT8_PEAK_TOPS = 4528.7   # measured tcgen05.mma kind::i8 issue peak, M=128 N=256, 148 SMs (profiles/r01_i8_tcgen05_probe.txt)
VARSIGMA = float(erfcinv(0.01))  # UCB multiplier of the reference (gp_surrogate.py:397)
CPU_SAMPLE = 8192         # candidates per CPU-baseline step (bounded sample of the same workload)
LOGICAL_SHARDS = 64       # the candidate matrix is generated in 64 seeded pieces, so it is the same for every GPU count


def synthetic_training(N, d):
    rng = np.random.default_rng(SEED)
    X = rng.random((N, d))
    y = np.sin(3.0 * X.sum(axis=1)) + 0.01 * rng.standard_normal(N)
    return X, y[:, None]


def fill_candidates(out, start, stop, M, d):
    """Rows [start, stop) of the M x d synthetic candidate matrix into ``out``.  Logical shard s (rows s*ceil(M/64) ...) comes
    from ``default_rng([SEED, s])``, whatever the number of ranks: every GPU count scores the same matrix."""
    per = -(-M // LOGICAL_SHARDS)
    for s in range(start // per, (stop - 1) // per + 1 if stop > start else 0):
        lo, hi = s * per, min(M, (s + 1) * per)
        block = np.random.default_rng([SEED, s]).random((hi - lo, d))
        a, b = max(lo, start), min(hi, stop)
        out[a - start:b - start] = block[a - lo:b - lo]


def pack_unconstrained(ls, variance, noise, c):
    """Unconstrained L-BFGS-B variables of (lengthscale, kernel variance, noise variance, constant mean): softplus^-1, with
    GPflow's 1e-6 floor under the noise variance."""
    inv = lambda v: float(np.log(np.expm1(v)))
    return np.array([inv(ls), inv(variance), inv(noise - 1.0e-6), c])


def fixed_theta(d):
    return np.array([0.25 * np.sqrt(d), 1.0, 1.0e-3, 0.0])


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        infos = [i for i in threadpool_info() if i.get("user_api") == "blas"]
        if infos:
            return int(max(i["num_threads"] for i in infos)), infos[0].get("internal_api", "blas")
    except Exception:
        pass
    return os.cpu_count() or 1, "unknown"


def cpu_reference_step(X, y, theta, Xc, varsigma):
    """The reference path on the CPU: predict_y (Cholesky inside, Kmn materialised, two TRSMs) + UCB + argmax."""
    from oracle import gpr_oracle as go

    h = go.Hyper(theta[0], theta[1], theta[2], theta[3])
    mean, var = go.predict_y("Matern52", X, y, h, Xc, chunk=65536)
    return go.ucb_argmax(mean, var, varsigma)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with nvidia-smi while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, smax, reasons = [], 0.0, set()
        for row in self.rows:
            try:
                sm.append(float(row[0]))
                smax = max(smax, float(row[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), row[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def bench_lml_grad(cuda, cpu=True, shapes=((4096, 10), (8192, 20)), evals=5):
    """LML + gradient evaluations per second through the C ABI (gpso_neg_lml_grad: Gram -> Cholesky -> L^-1 -> K_y^-1 ->
    fused gradient reduction), host in / host out, at the shapes of configs C3 and C4.  Algorithmic work: N^3 flops."""
    out = []
    for N, d in shapes:
        X, y = synthetic_training(N, d)
        u = pack_unconstrained(0.25 * np.sqrt(d), 1.0, 1.0e-3, 0.0)
        sess = cuda.open_session("Matern52", 1, True)
        sess.set_data(X, y)
        f, g = sess.neg_lml_and_grad(u)  # warm-up (allocations)
        t0 = time.perf_counter()
        dev_ms = 0.0
        for i in range(evals):
            f, g = sess.neg_lml_and_grad(u + 1e-3 * (i + 1))
            dev_ms += sess.last_timing_ms()[0]
        wall = time.perf_counter() - t0
        # the same closure with the step-by-step launches (gpso_set_factor_mode 0) beside the persistent tile scheduler
        sess.set_factor_mode(False)
        sess.neg_lml_and_grad(u)
        step_ms = 0.0
        for i in range(evals):
            sess.neg_lml_and_grad(u + 1e-3 * (i + 1))
            step_ms += sess.last_timing_ms()[0]
        sess.set_factor_mode(True)
        # ... and with K_y^-1 = L^-T L^-1 on the FP64 DMMA tiles instead of the int8 tensor-core product (automatic from N=512)
        sess.set_kinv_mode(1)
        sess.neg_lml_and_grad(u)
        kinv_dmma_ms = 0.0
        for i in range(evals):
            sess.neg_lml_and_grad(u + 1e-3 * (i + 1))
            kinv_dmma_ms += sess.last_timing_ms()[0]
        sess.set_kinv_mode(0)
        rec = {"N": N, "d": d, "evals_per_s": evals / wall, "device_ms_per_eval": dev_ms / evals,
               "device_ms_per_eval_stepwise_launches": step_ms / evals,
               "device_ms_per_eval_kinv_on_fp64_dmma": kinv_dmma_ms / evals,
               "schedule": "blocked Cholesky as ONE persistent kernel on FP64 DMMA (one CTA per SM, host-built task list with "
                           "per-tile dependency counters, two-level blocking + look-ahead); then L^-1 by recursive doubling (two "
                           "products per level, 8-digit / 62-bit fixed-point operands) and K_y^-1 = L^-T L^-1 (7 digits / 54 bit) as "
                           "exact-integer products on the int8 tensor cores (tcgen05.mma kind::i8)",
               "fp64_tflops_note": "N^3 fp64-equivalent flops per evaluation / device time: N^3/3 (Cholesky) run on the FP64 pipe, "
                                   "2N^3/3 (L^-1, K_y^-1) as int8 tensor-core products, so the ratio to the FP64 pipe peak can exceed 1",
               "library_bar_ms": {4096: {"cusolver_dpotrf": 1.675, "cusolver_dpotri": 14.005},
                                  8192: {"cusolver_dpotrf": 7.753, "cusolver_dpotri": 61.528}}.get(N),
               "fp64_tflops": (float(N) ** 3 / (dev_ms / evals * 1e-3)) / 1e12, "fp64_peak_tflops": FP64_PEAK_TFLOPS,
               "frac_of_fp64_peak": (float(N) ** 3 / (dev_ms / evals * 1e-3)) / 1e12 / FP64_PEAK_TFLOPS, "neg_lml": f}
        if cpu and N <= 4096:
            from oracle import gpr_oracle as go  # CPU leg of the LML side measurement (checker + baseline)

            t0 = time.perf_counter()
            f_ref, g_ref = go.neg_lml_and_grad("Matern52", X, y, u + 1e-3 * evals, 1, True)
            rec["cpu_evals_per_s"] = 1.0 / (time.perf_counter() - t0)
            rec["lml_rel_err_vs_cpu"] = abs(f - f_ref) / max(abs(f_ref), N)
        sess.close()
        out.append(rec)
    return out


def run_reference(args, rank):
    """--impl reference: the CPU path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    N, d, M, desc = WORKLOADS[args.workload]
    X, y = synthetic_training(N, d)
    theta = fixed_theta(d)
    sample = min(M, CPU_SAMPLE)
    Xc = np.empty((sample, d))
    fill_candidates(Xc, 0, sample, M, d)
    for _ in range(args.warmup):
        cpu_reference_step(X, y, theta, Xc, VARSIGMA)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(X, y, theta, Xc, VARSIGMA)
    dt = time.perf_counter() - t0
    cores, blas = blas_threads()
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "predict_y+UCB candidates/sec", "value": value, "unit": "candidates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "N": N, "d": d, "M": M, "kernel": "Matern52", "theta": theta.tolist()},
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of the {M} candidates per step (numpy/scipy restatement of GPflow's predict_y op "
                                   f"sequence incl. the per-call Cholesky; {blas} BLAS, {cores} threads; GPflow itself is not "
                                   "installable here)"},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--candidates", type=int, default=0, help="override the number of candidates (smoke runs only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lml", action="store_true", help="skip the LML+grad evaluations/s side measurement")
    ap.add_argument("--engine", default="auto", choices=["auto", "dmma", "int8"],
                    help="variance-product engine: FP64 DMMA or the exact-integer int8 tcgen05 emulation (auto picks int8 at this size)")
    ap.add_argument("--no-overlap", action="store_true", help="int8 engine: run cross-covariance and product back to back")
    ap.add_argument("--slices", type=int, default=0, help="8-bit digits per operand for the int8 engine (0 = automatic)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from pygpso_b200 import backend, gpmodel
    from pygpso_b200.distributed import ShardedScorer, gather_records, pick_best, shard_bounds

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (pygpso_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("GPSO_DEVICE", str(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    N, d, M, desc = WORKLOADS[args.workload]
    if args.candidates:
        M = args.candidates
    varsigma = VARSIGMA
    X, y = synthetic_training(N, d)
    theta = fixed_theta(d)

    # the model object of the public API (GPR == the reference's gpflow_model); rank 0 fits, the others import
    kernel = gpmodel.Matern52(lengthscales=theta[0], variance=theta[1])
    model = gpmodel.GPR(data=(X, y), kernel=kernel, mean_function=gpmodel.Constant(theta[3]), noise_variance=theta[2],
                        backend=backend.CudaBackend(device=local_rank))
    session = model._session
    session.set_predict_mode({"auto": 0, "dmma": 1, "int8": 2}[args.engine], args.slices)
    session.set_overlap(not args.no_overlap)
    t0 = time.perf_counter()
    if rank == 0:
        model._ensure_factor()  # Gram -> Cholesky -> L^-1 -> alpha on rank 0 only
    factor_ms = (time.perf_counter() - t0) * 1e3
    broadcast_ms = None
    scorer = None
    if world > 1:
        scorer = ShardedScorer(session)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        scorer.broadcast_fit(N, d, src=0)  # one NCCL broadcast of (scaled X, alpha, L^-1, theta) per fit
        torch.cuda.synchronize()
        dist.barrier()
        broadcast_ms = (time.perf_counter() - t0) * 1e3
        model._factor_key = model._theta().tobytes()

    # this rank's shard of the candidates, in pinned host memory (for the end-to-end leg) and resident in HBM
    start, stop = shard_bounds(M, world, rank)
    m_local = stop - start
    host = torch.empty((max(m_local, 1), d), dtype=torch.float64, pin_memory=True)
    xc_host = host.numpy()[:m_local]
    fill_candidates(xc_host, start, stop, M, d)
    xc_dev = host[:m_local].to("cuda", non_blocking=False)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        rec = [-np.inf, -1.0, 0.0, 0.0]
        if m_local:
            idx, mean, var, ucb = session.ucb_argmax_dev(xc_dev.data_ptr(), m_local, varsigma, stream)
            rec = [ucb, float(start + idx), mean, var]
        return pick_best(gather_records(rec)) if world > 1 else (int(rec[1]), rec[2], rec[3], rec[0])

    def step_e2e():
        # public API with HOST buffers: H2D of the candidates + D2H of the result inside the call
        if world > 1:
            return scorer.ucb_argmax(xc_host, start, varsigma)
        return model.ucb_argmax(xc_host, varsigma)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident measurement -------------------------------------------------------------------------------
    for _ in range(args.warmup):
        result = step_device()
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = session.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(4)
    windows = 0
    ev0.record()
    for _ in range(args.steps):
        result = step_device()
        stage_ms += session.last_timing_ms()
        windows += session.last_windows()
    ev1.record()
    sync_all()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = session.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    product_ms = stage_ms[2]  # summed duration of the variance-product launches (event pairs on their stream)
    # per-stage breakdown: one extra, untimed step with per-stage events (stages run back to back, no stream overlap)
    session.set_profile(True)
    step_device()
    stage_profile = session.last_timing_ms()
    session.set_profile(False)
    engine = session.predict_info()

    # ---- end-to-end measurement (host buffers through the public API) ------------------------------------------------
    for _ in range(min(args.warmup, 1)):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        result_e2e = step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if world > 1:
        dist.barrier()

    if rank == 0:
        assert result_e2e[0] == result[0], "end-to-end and device-resident passes selected different candidates"
        value = M * args.steps / (dev_ms * 1e-3)
        e2e_value = M * args.steps / e2e_s
        flops64 = float(N) * N * m_local * args.steps  # algorithmic fp64 work: N^2 per candidate
        windows = max(windows, 1)
        if engine["engine"] == "int8-tcgen05":
            S = engine["slices"]
            pairs = S * (S + 1) // 2
            # the same N^2/2 multiply-adds per candidate, once per retained digit pair, on the int8 tensor pipe
            ops = pairs * flops64
            achieved = ops / (product_ms * 1e-3) / 1e12
            roofline = {
                "bound": "tensor",
                "kernel": f"ozaki_trmm_kernel<{S}> (tcgen05.mma kind::i8, {S} 8-bit digits per operand, {pairs} digit pairs, "
                          "TMEM accumulators, exact int32 sums recombined to fp64 in the epilogue)",
                "achieved": achieved, "peak": INT8_PEAK_TOPS, "unit": "TFLOP/s", "frac": achieved / INT8_PEAK_TOPS,
                "op_kind": "int8 tensor op (2 per multiply-add); algorithmic = digit_pairs * N^2 per candidate",
                "traffic": None,
                "peak_source": "measured tcgen05 kind::i8 issue peak on this pool's B200 at 1965 MHz (profiles/"
                               "r01_i8_tcgen05_probe.txt, nominal 4500); in this kernel the SM clock settles near 1.6-1.7 GHz "
                               "(see clocks); MEASURED_PEAKS.json has no int8 entry.  The FP64 cross-covariance kernel cannot "
                               "hide behind it: DFMA shares the tensor-core datapath on B200 (x7.9 slower beside int8 MMAs, "
                               "profiles/r01s4_corun_probe.txt), so a step is product + cross-covariance in sequence",
                "fp64_equivalent": {"achieved_tflops": flops64 / (product_ms * 1e-3) / 1e12, "fp64_pipe_peak_tflops": FP64_PEAK_TFLOPS,
                                    "ratio_to_fp64_peak": flops64 / (product_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS},
                "digits": S, "error_estimate_over_tolerance": engine["error_estimate_over_tol"],
            }
        else:
            achieved = flops64 / (product_ms * 1e-3) / 1e12
            roofline = {
                "bound": "tensor", "kernel": "predict_trmm_kernel (FP64 DMMA triangular product + column sum of squares)",
                "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS, "traffic": None,
                "peak_source": "measured DMMA.8x8x4 issue peak on this pool's B200 (profiles/r01_fp64_probe.txt; cuBLAS dgemm "
                               "8192^3 reaches 36.06); MEASURED_PEAKS.json has no FP64 entry",
            }
        prof_json = os.path.join(ROOT, "profiles", "product_kernel_traffic.json")
        if os.path.exists(prof_json):
            try:
                roofline["traffic"] = json.load(open(prof_json)).get(engine["engine"], {}).get("dram_bytes_per_launch")
            except Exception:
                pass
        roofline.update({
            "launches": int(windows), "avg_launch_ms": product_ms / windows,
            "algorithmic_ops_per_launch": (roofline["achieved"] * 1e12 * product_ms * 1e-3) / windows,
            "product_share_of_step": product_ms / dev_ms,
            "hbm_algorithmic_gbs": (80.0 * M * args.steps / (dev_ms * 1e-3)) / 1e9,
            "stage_ms_one_step_serialised": {"crosscov": stage_profile[1], "product": stage_profile[2], "finalize": stage_profile[3]},
        })
        line = {
            "metric": "predict_y+UCB candidates/sec", "value": value, "unit": "candidates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "N": N, "d": d, "M": M, "kernel": "Matern52", "theta": theta.tolist(),
                       "varsigma": varsigma, "parallelism": f"candidates sharded over {world} GPU(s)",
                       "engine": engine,
                       "l2": "inputs larger than L2: 80 B/candidate x M candidates in HBM plus two 2 GiB rolling windows of "
                             "cross-covariance digit tiles per GPU vs 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": int(M) * d * 8,
                    "d2h_bytes_per_step": 32 * world, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "setup": {"factorize_ms": factor_ms, "broadcast_ms": broadcast_ms, "state_bytes": session.state_bytes(N, d)},
            "result": {"index": int(result[0]), "mean": result[1], "var": result[2], "ucb": result[3]},
        }
        # ---- second half of the BASELINE metric: LML + gradient evaluations per second (the L-BFGS-B closure) ---------
        if world == 1 and not args.no_lml:
            line["lml_grad"] = bench_lml_grad(backend.CudaBackend(device=local_rank), cpu=not args.no_cpu_baseline)
        # ---- CPU baseline beside it (bounded sample, rank 0, single GPU runs only) -----------------------------------
        if world == 1 and not args.no_cpu_baseline:
            sample = min(M, CPU_SAMPLE)
            xs = xc_host[:sample]
            cpu_reference_step(X, y, theta, xs[:256], varsigma)
            t0 = time.perf_counter()
            ref = cpu_reference_step(X, y, theta, xs, varsigma)
            cpu_s = time.perf_counter() - t0
            cores, blas = blas_threads()
            # the same sample through the CUDA path must select the same candidate
            got = session.ucb_argmax(xs, varsigma)
            line["cpu_baseline"] = {
                "value": sample / cpu_s, "unit": "candidates/s", "cores": cores, "kind": "port",
                "sample": f"first {sample} of the {M} candidates, one pass (numpy/scipy restatement of GPflow's predict_y op "
                          f"sequence incl. the per-call Cholesky; {blas} BLAS, {cores} threads)",
                "same_argmax_as_gpu": bool(got[0] == ref[0]),
            }
        print(json.dumps(line), flush=True)
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
