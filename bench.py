#!/usr/bin/env python3
"""Benchmark of the SympGPR hot path on B200 (driver contract: one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL)

Headline metric (BASELINE.json): NLL+gradient evaluations per second at N = 16 384 training
pairs (n = 32 768, fp64), synthetic standard map K = 0.9, product kernel.  One "step" = one
NLL+gradient evaluation: fill -> potrf (+ forward substitution) -> trtri -> lauum -> gradient contraction.
With N > 1 GPUs every rank evaluates its own multi-start restart (a different theta; the
training Cholesky does not shard -- DESIGN.md "Multi-GPU") and its own slice of the orbit
ensembles / Sobol sample set; NCCL only gathers results ("scaling": "weak").

Also on the same line, each with its own roofline and CPU figure where one exists:
  map        BASELINE config 4: 4096 training pairs, 1e5 orbits x 1000 steps per GPU (second metric), newton_delta and
             the reference's own hybrd1 over the full 1000 steps, CPU oracle timed on a slice; map.big = the 10^7-orbit
             ensemble of config 5 split over the GPUs (strong scaling)
  sweep      config 5: NLL+gradient at N = 2048 ... 32 768 (n up to 65 536)
  configs    01_pendulum (N = 200: latency), 03_henon_heiles (2-DOF 4 x 4-block kernel, N = 8192), 05_tokamak (16 384
             training pairs, tokamak map kind with its loss test, Sobol sample set of 10^6 rows through
             sympgpr_b200.ensemble.sobol_indices_sharded)
  roofline_fill, stages, cpu_baseline (measured points; anything extrapolated sits under `extrapolated`).

--impl reference: the CPU arm (oracle port of the reference maths on the host cores, rank 0 only): W warm-up + K timed
evaluations at a bounded N_s chosen so that the run ends within minutes; `value` is what was MEASURED at N_s, the
measured points at other sizes and the a n^3 + b n^2 fit through them are separate keys.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nll_grad_evals_per_s"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-train", type=int, default=16384, help="training pairs N (matrix order 2N)")
    ap.add_argument("--map-train", type=int, default=4096, help="training pairs of the map leg (config 04_standard_map)")
    ap.add_argument("--orbits", type=int, default=100000, help="orbits per GPU in the map leg")
    ap.add_argument("--map-steps", type=int, default=1000, help="map steps per launch in the map leg (config 04: 1000)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds the timed region of the CPU arm may take")
    ap.add_argument("--route", default="dmma", choices=["dmma", "int8"],
                    help="dmma (default, what north_star names): FP64 tensor-core DMMA everywhere; int8: the opt-in route of "
                         "DESIGN.md 4.1 (factor + inverse + lauum from INT8 digit products on tcgen05) as the HEADLINE measurement")
    ap.add_argument("--int8-digits", type=int, default=6, help="digits per operand of --route int8 (6: 47 bits, 7: 55 bits)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-map", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 01 / 03 / 05 legs")
    ap.add_argument("--big-orbits", type=int, default=10_000_000,
                    help="TOTAL orbits of the 10^7-orbit prediction leg (BASELINE config 5), split over the GPUs; 0 skips it")
    ap.add_argument("--big-steps", type=int, default=0,
                    help="map steps of the 10^7-orbit leg; 0 = 1000 on 8 GPUs (42 s), 16 otherwise (1000 steps on one GPU take 330 s)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the NLL+gradient size sweep")
    ap.add_argument("--sweep-max", type=int, default=32768, help="largest N of the sweep (32768 -> n = 65 536, 77 GB of workspace)")
    ap.add_argument("--sobol-samples", type=int, default=1_000_000, help="rows of the Sobol sample set (config 5), whole job")
    ap.add_argument("--sobol-steps", type=int, default=16, help="map steps per model run of the Sobol leg")
    ap.add_argument("--tok-train", type=int, default=16384, help="training pairs of the tokamak model (config 5)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU legs (oracle = checker / baseline)
FULLSIZE_RECORD = os.path.join(ROOT, "profiles", "r01_cpu_fullsize_record.json")


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def _all_threads():
    """torchrun exports OMP_NUM_THREADS=1: lift the BLAS/OpenMP limit for the CPU arm (context manager or no-op)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=os.cpu_count() or 1)
    except ImportError:
        import contextlib
        return contextlib.nullcontext()


def cpu_nll_grad_time(N, nrep=1, nwarm=0):
    """Seconds per NLL+gradient evaluation of the oracle (port of python/02_pert_pendulum/func.py:148-162: C/NumPy fill,
    SciPy potrf + potri on all host threads, elementwise contraction) at N training pairs, MEASURED; list of nrep times."""
    from oracle import oracle as O
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    for _ in range(nwarm):
        O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    ts = []
    for _ in range(nrep):
        t = time.perf_counter()
        O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
        ts.append(time.perf_counter() - t)
    return ts


def fit_cubic(points):
    """Least squares t(n) = a n^3 + b n^2 through measured (n, seconds) points, a, b >= 0."""
    n = np.array([p[0] for p in points], float)
    t = np.array([p[1] for p in points], float)
    A = np.stack([n**3, n**2], axis=1)
    w = 1.0 / t                                     # relative residuals
    coef, *_ = np.linalg.lstsq(A * w[:, None], t * w, rcond=None)
    a_c, b_c = float(coef[0]), float(coef[1])
    if a_c <= 0 or b_c < 0:
        a_c, b_c = float(t[-1] / n[-1]**3), 0.0
    return a_c, b_c


def fullsize_record():
    try:
        with open(FULLSIZE_RECORD) as f:
            return json.load(f)
    except Exception:
        return None


def cpu_map_sample(Nt, E_s, steps_s):
    """The reference's map loop on the host cores (C oracle, sympgpr.f90:88-177 restated, OpenMP over orbits, hybrd1 started
    at p + guess like the timed GPU solver) on E_s orbits of the benchmark ensemble for steps_s steps: orbit-steps/s."""
    import scipy.linalg
    from oracle import c_oracle as C, oracle as O
    d = O.standard_map_training(Nt)
    l = 2 * np.pi / np.sqrt(Nt)
    hyp, hypp = np.array([l, l, d["sig"], 1e-8]), np.array([l, l, d["sigp"], 1e-8])
    xt, zt, xtp, ztp = d["xtrain"], d["ztrain"], d["xtrainp"], d["ztrainp"]
    K = C.build_k(xt[:Nt], xt[Nt:], xt[:Nt], xt[Nt:], hyp[:3])
    K[np.diag_indices_from(K)] += hyp[3]
    alpha = scipy.linalg.cho_solve(scipy.linalg.cho_factor(K, lower=True), zt)
    Kp = C.buildkreg(xtp[:Nt], xtp[Nt:], xtp[:Nt], xtp[Nt:], hypp[:3])
    Kp[np.diag_indices_from(Kp)] += hypp[3]
    alphap = scipy.linalg.cho_solve(scipy.linalg.cho_factor(Kp, lower=True), ztp)
    from sympgpr_b200 import workloads as W
    q0a, p0a = W.ensemble(100000)
    idx = np.linspace(0, 99999, E_s).astype(int)
    t = time.perf_counter()
    out = C.applymap_alpha(2, steps_s + 1, q0a[idx], p0a[idx], hyp[:3], hypp[:3], xtp[:Nt], xtp[Nt:], alphap, xt[:Nt], xt[Nt:],
                           alpha, start_delta=True, out_every=steps_s)
    dt = time.perf_counter() - t
    return {"value": E_s * steps_s / dt, "unit": "orbit-steps/s", "cores": C.num_threads(), "kind": "port",
            "sample": f"C oracle (hybrd1 started at p + guess, tol 1e-13, OpenMP over orbits), {E_s} orbits of the benchmark "
                      f"ensemble x {steps_s} steps at Nt={Nt}: {dt:.1f} s, {out[2]:.1f} residual evaluations per orbit-step",
            "seconds": dt}


def run_reference(a):
    """CPU arm.  W warm-up + K timed NLL+gradient evaluations of the oracle at a bounded N_s (the largest power of two whose
    (K + W) evaluations fit --cpu-budget, calibrated on N = 512), then one evaluation each at the larger sizes that
    still fit (the basis of the extrapolation).  `value` / `ms_per_step` are the MEASURED rate at N_s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, a.steps), max(0, a.warmup)
    with _all_threads():
        cores = blas_threads()
        t512 = min(cpu_nll_grad_time(512, 2, 1))
        N_s = 512
        for cand in (1024, 2048, 4096):
            if cand <= a.n_train and (steps + warm) * t512 * (cand / 512.0) ** 2.7 <= a.cpu_budget:
                N_s = cand
        t_run0 = time.perf_counter()
        ts = cpu_nll_grad_time(N_s, steps, warm)
        t_timed = float(np.sum(ts))
        t_s = t_timed / steps
        points = [(2 * 512, t512), (2 * N_s, t_s)]
        extra_budget = 90.0
        for cand in (2048, 4096, 8192):
            if cand <= N_s or cand > a.n_train:
                continue
            est = points[-1][1] * (2.0 * cand / points[-1][0]) ** 2.8
            if est > extra_budget:
                break
            tt = cpu_nll_grad_time(cand, 1, 0)[0]
            extra_budget -= tt
            points.append((2 * cand, tt))
    a_c, b_c = fit_cubic(points)
    nf = 2.0 * a.n_train
    t_full = a_c * nf**3 + b_c * nf**2
    v = 1.0 / t_s
    rec = fullsize_record()
    sample = (f"oracle nll_grad (port of python/02_pert_pendulum/func.py:148-162: C/NumPy fill, SciPy potrf+potri, elementwise "
              f"contraction) on {cores} host threads: {warm} warm-up + {steps} timed evaluations at N={N_s} training pairs "
              f"(n={2 * N_s}), {t_s:.3f} s each -- a bounded sample of the N={a.n_train} workload (1/{(a.n_train / N_s) ** 3:.0f} of "
              f"its n^3 flops); value and ms_per_step are this measurement")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * t_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(a),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "sample_n_train": N_s,
            "measured_points": [{"n": int(n_), "N": int(n_ // 2), "seconds_per_eval": float(t_)} for n_, t_ in points],
            "extrapolated": {"n_train": a.n_train, "seconds_per_eval": t_full, "value": 1.0 / t_full, "unit": UNIT,
                             "how": f"t(n) = a n^3 + b n^2, relative least squares through the measured points: a={a_c:.3e}, "
                                    f"b={b_c:.3e}; NOT a measurement",
                             "fullsize_record": rec},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(a):
    return {"workload": f"synthetic standard map K=0.9, N={a.n_train} training pairs (n={2 * a.n_train} fp64 Hessian-block "
                        f"kernel, periodic x SE), NLL+gradient; map leg: N={a.map_train}, {a.orbits} orbits/GPU x "
                        f"{a.map_steps} steps", "n_train": a.n_train, "matrix_order": 2 * a.n_train,
            "hyp": "lx=ly=0.5*2pi/sqrt(N), sig2n=1e-8", "parallelism": f"restarts+ensemble x{a.gpus}",
            "l2": "inputs larger than L2 (8.6 GB matrix vs 126 MB)"}


# ------------------------------------------------------------------------------------------ b200 arm
_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else that writes to fd 1 while the bench
    runs (NCCL's version banner, library chatter) has been sent to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


class Bench:
    """State shared by the legs of the b200 arm."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        from sympgpr_b200 import _lib, api, ensemble, workloads
        self.a, self.torch, self.dist, self._lib, self.api, self.En, self.W = a, torch, dist, _lib, api, ensemble, workloads
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        if not torch.cuda.is_available() or _lib.device_count() < 1:
            raise RuntimeError("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.L = _lib.lib()
        self.ctx = _lib.context(self.local)
        # One explicit (non-default) stream for everything this process launches: the library's kernels, cuBLAS for
        # the peak probe and the CUDA events that time them.  (The legacy default stream has handle 0, which
        # sgp_set_stream reads as "use your own stream".)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)
        self.ctx.set_stream(self.stream.cuda_stream)
        _lib.check(self.L.sgp_set_profiling(self.ctx.handle, 1), "sgp_set_profiling")
        self.sampler = ClockSampler(self.local)

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def timed(self, fn, reps):
        """Average milliseconds of `reps` back-to-back calls of fn on the bench stream (CUDA events, sync on both sides)."""
        self.torch.cuda.synchronize()
        e0, e1 = self.ev(), self.ev()
        e0.record(self.stream)
        for _ in range(reps):
            fn()
        e1.record(self.stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    def free(self):
        self.ctx.release_workspace()
        self.torch.cuda.empty_cache()

    def nll_dev_fn(self, hyp, x_d, z_d, n, reg=0, ngrad=2, fam=0):
        hyp_c = (ctypes.c_double * 4)(*hyp)
        res_d = self.torch.zeros(16, dtype=self.torch.float64, device=self.dev)

        def step():
            self._lib.check(self.L.sgp_nll_dev(self.ctx.handle, fam, 0.5, reg, hyp_c, x_d.data_ptr(), z_d.data_ptr(), n, ngrad,
                                               res_d.data_ptr()), "sgp_nll_dev")
        return step, res_d

    def stage_times(self):
        stage_ms = (ctypes.c_double * 7)()
        self._lib.check(self.L.sgp_stage_times(self.ctx.handle, stage_ms), "sgp_stage_times")
        return dict(zip(["fill", "potrf", "potrs", "trtri", "lauum", "grad", "finalize"], [float(v) for v in stage_ms]))


FP64_PEAK_MEASURED = None          # DFMA-loop probe of leg_peaks (thread-level DP instructions per second)


def fp64_pipe_peak():
    """(peak, how): the DFMA rate measured in this run, else the nominal 148 SMs x 64 lanes x 1.965 GHz."""
    if FP64_PEAK_MEASURED:
        return FP64_PEAK_MEASURED, "register-resident DFMA loop measured in this run (sgp_bench_dfma); nominal 148 x 64 x 1.965 GHz = 1.861e13"
    return 148 * 64 * 1.965e9, "nominal 148 SMs x 64 FP64 lanes x 1.965 GHz"


def leg_peaks(B):
    """FP64 tensor peak (cuBLAS DGEMM 8192^3) and HBM write-only peak (memset of 8 GiB), both measured in this run."""
    torch = B.torch
    A_ = torch.randn(8192, 8192, dtype=torch.float64, device=B.dev)
    B_ = torch.randn(8192, 8192, dtype=torch.float64, device=B.dev)
    torch.matmul(A_, B_)
    best = 1e9
    for _ in range(4):
        best = min(best, B.timed(lambda: torch.matmul(A_, B_), 1) * 1e-3)
    dgemm = 2 * 8192 ** 3 / best / 1e12
    del A_, B_
    buf = torch.empty(1 << 30, dtype=torch.float64, device=B.dev)        # 8 GiB
    buf.zero_()
    bw = 1e9
    for _ in range(3):
        bw = min(bw, B.timed(lambda: buf.zero_(), 2) * 1e-3)
    write_gbs = buf.numel() * 8 / bw / 1e9
    del buf
    torch.cuda.empty_cache()
    # vector FP64 ceiling: register-resident DFMA loop (thread-level DFMA instructions per second)
    global FP64_PEAK_MEASURED
    r = ctypes.c_double(0.0)
    B._lib.check(B.L.sgp_bench_dfma(B.ctx.handle, 3, ctypes.byref(r)), "sgp_bench_dfma")
    FP64_PEAK_MEASURED = float(r.value)
    return dgemm, write_gbs


def leg_headline(B):
    a, torch, W = B.a, B.torch, B.W
    N = a.n_train
    n = 2 * N
    d = W.standard_map_training(N)
    hyp = W.timing_hyp(N, d["sig"], 1e-8)
    hyp[:2] *= (1.0 + 0.02 * B.rank)                 # multi-start restart of this rank: a different length-scale pair per GPU
    x_h = torch.from_numpy(d["xtrain"].copy()).pin_memory()
    z_h = torch.from_numpy(d["ztrain"].copy()).pin_memory()
    x_d, z_d = x_h.to(B.dev), z_h.to(B.dev)
    step_dev, res_d = B.nll_dev_fn(hyp, x_d, z_d, n)
    if a.route == "int8":
        B.ctx.set_ozaki_ex(a.int8_digits, 3, 4096)
    for _ in range(max(a.warmup, 1)):
        step_dev()
    B.sync_all()
    launches0 = B.L.sgp_launch_count()
    t_wall0 = time.time()
    e0, e1 = B.ev(), B.ev()
    e0.record(B.stream)
    for _ in range(a.steps):
        step_dev()
    e1.record(B.stream)
    B.sync_all()
    t_wall1 = time.time()
    launches = int(B.L.sgp_launch_count() - launches0)
    t_dev = B.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    stages = B.stage_times()                          # of the last timed step
    res = res_d.cpu().numpy()
    if res[4] != 0:
        raise RuntimeError(f"Cholesky failed in the timed region (info={res[4]})")
    # end to end through the public API with host buffers
    xh_np, zh_np = x_h.numpy(), z_h.numpy()
    B.api.nll_grad(hyp, xh_np, zh_np, n)
    B.sync_all()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        B.api.nll_grad(hyp, xh_np, zh_np, n)
    torch.cuda.synchronize()
    t_e2e = B.max_over_ranks(time.perf_counter() - t0)
    if a.route == "int8":
        B.ctx.set_ozaki_ex(0, 1, 0)
    # full (2N x 2N) Hessian-block fill: the HBM-write roofline case (8 n^2 bytes)
    Kbuf = torch.empty(n * n, dtype=torch.float64, device=B.dev)
    hyp3 = (ctypes.c_double * 3)(*hyp[:3])
    xq, yP = x_d[:N], x_d[N:]

    def fill_full():
        B._lib.check(B.L.sgp_build_k_dev(B.ctx.handle, 0, 0.5, xq.data_ptr(), yP.data_ptr(), N, xq.data_ptr(), yP.data_ptr(), N,
                                         hyp3, Kbuf.data_ptr(), n), "sgp_build_k_dev")
    for _ in range(3):
        fill_full()
    t_fill = B.timed(fill_full, 5) * 1e-3
    del Kbuf
    B.free()
    return dict(N=N, n=n, t_dev=t_dev, stages=stages, res=res, launches=launches, wall=(t_wall0, t_wall1),
                value=B.world * a.steps / t_dev, ms_per_step=1e3 * t_dev / a.steps, e2e_value=B.world * a.steps / t_e2e,
                h2d=int(x_h.numel() * 8 + z_h.numel() * 8), d2h=16 * 8, t_fill=t_fill, fill_gbs=8.0 * n * n / t_fill / 1e9)


def leg_ozaki(B, head_ms, head_res, dgemm, bf16_peak):
    """OPT-IN route past the DMMA ceiling (not the headline: north_star asks for DMMA in the Cholesky): FP64 products from the
    INT8 tensor pipe (tcgen05.mma kind::i8, TMEM accumulators, Ozaki slicing; csrc/ozaki.cu, csrc/ozaki_chol.cu).  (a) the raw
    GEMM at 8192^3 against the DMMA GEMM, (b) the headline evaluation with ALL THREE n^3/3 stages on the INT8 pipe (factor +
    triangular inverse as one recursion of sliced products, lauum as one sliced product; 7 slices), device resident and end to
    end from host buffers, with the SM clock sampled during the run, (c) the lauum-only variant of round-2 state "e"."""
    a, torch, W = B.a, B.torch, B.W
    out = {"note": "opt-in (sgp_set_ozaki_ex); headline value/roofline above are the DMMA path north_star names"}
    g = {}
    for ns in (6, 7):
        ms = (ctypes.c_double * 2)()
        B._lib.check(B.L.sgp_ozaki_bench(B.ctx.handle, ns, 8192, 8192, 8192, 3, ms), "sgp_ozaki_bench")
        g[f"{ns}_slices"] = {"ms_slicing_plus_gemm": ms[0], "ms_gemm": ms[1], "fp64_equiv_TFLOP/s": 2 * 8192.0**3 / ms[0] / 1e9,
                             "fp64_equiv_TFLOP/s_gemm_only": 2 * 8192.0**3 / ms[1] / 1e9,
                             "int8_TOP/s_gemm_only": ns * (ns + 1) / 2 * 2 * 8192.0**3 / ms[1] / 1e9}
    msd = ctypes.c_double(0.0)
    B._lib.check(B.L.sgp_bench_gemm(B.ctx.handle, 0, 0, 0, 64, 64, 8192, 3, ctypes.byref(msd)), "sgp_bench_gemm")
    g["dmma_gemm_f64_ws"] = {"ms": msd.value, "TFLOP/s": 2 * 8192.0**3 / msd.value / 1e9}
    g["cublas_dgemm_TFLOP/s"] = dgemm
    # INT8 has no entry in MEASURED_PEAKS.json; the dense INT8 rate of the tensor core is twice its bf16 rate (nominal 4.5 POP/s
    # against 2.25 PFLOP/s), so 2 x the MEASURED bf16 burst figure is the denominator
    if bf16_peak:
        tops = g["6_slices"]["int8_TOP/s_gemm_only"]
        g["roofline"] = {"kernel": "oz_gemm_kernel<6>", "bound": "tensor", "achieved": tops, "peak": 2.0 * bf16_peak, "unit": "INT8 TOP/s",
                         "frac": tops / (2.0 * bf16_peak),
                         "note": "21 digit-pair products of 128 x 64 x K per output tile; peak = 2 x bf16_tflops of MEASURED_PEAKS.json "
                                 "(no INT8 entry there; burst figure, the kernel is timed alone)"}
    out["gemm_8192"] = g
    B.free()
    N = a.n_train
    n = 2 * N
    d = W.standard_map_training(N)
    hyp = W.timing_hyp(N, d["sig"], 1e-8)
    hyp[:2] *= (1.0 + 0.02 * B.rank)
    x_d = torch.from_numpy(d["xtrain"].copy()).to(B.dev)
    z_d = torch.from_numpy(d["ztrain"].copy()).to(B.dev)
    step, res_d = B.nll_dev_fn(hyp, x_d, z_d, n)

    def rel(res):
        return [abs(float(res[1 + k]) - float(head_res[1 + k])) / abs(float(head_res[1 + k])) for k in range(2)]

    # (b) all stages: 6 digits (47 bits of every row's scale, 21 digit pairs) and 7 digits (55 bits, 28 pairs: FP64-grade)
    for ns, key in ((6, "nll_grad_all_stages_int8"), (7, "nll_grad_all_stages_int8_7digits")):
        B.ctx.set_ozaki_ex(ns, 3, 4096)
        try:
            step(); step(); step()
            reps = max(3, a.steps)
            t0 = time.time()
            ms = B.timed(step, reps)
            t1 = time.time()
            st = B.stage_times()
            res = res_d.cpu().numpy()
            # end to end through the public API from pinned host buffers, as the headline's e2e; then the value-only evaluation
            x_h = torch.from_numpy(d["xtrain"].copy()).pin_memory()
            z_h = torch.from_numpy(d["ztrain"].copy()).pin_memory()
            xn, zn = x_h.numpy(), z_h.numpy()
            B.api.nll_grad(hyp, xn, zn, n)
            te = time.perf_counter()
            for _ in range(3):
                ve, ge = B.api.nll_grad(hyp, xn, zn, n)
            te = (time.perf_counter() - te) / 3
            B.api.nll_chol(hyp, xn, zn, n)
            tv = time.perf_counter()
            for _ in range(3):
                vv = B.api.nll_chol(hyp, xn, zn, n)
            tv = (time.perf_counter() - tv) / 3
        finally:
            B.ctx.set_ozaki_ex(0, 1, 0)
            B.free()
        t3 = st["potrf"] + st["trtri"] + st["lauum"]
        out[key] = {
            "digits": ns, "bits_of_row_scale": 8 * ns - 1, "digit_pairs": ns * (ns + 1) // 2, "leaf_rows_on_dmma": 4096,
            "ms_per_eval": ms, "evals_per_s": 1e3 / ms, "speedup_vs_dmma_path": head_ms / ms,
            "e2e": {"value": 1.0 / te, "unit": "evals/s", "ms_per_eval": 1e3 * te, "h2d_bytes_per_step": int(x_h.numel() * 8 + z_h.numel() * 8),
                    "d2h_bytes_per_step": 16 * 8},
            "value_only_e2e": {"ms_per_eval": 1e3 * tv, "nll": float(vv),
                               "note": "api.nll_chol (the objective of the scripts' L-BFGS loops) through the factor-only recursion, 0.38 n^3 flop"},
            "stages_ms": {k: round(v, 3) for k, v in st.items()},
            "stage_note": "potrf + trtri run as ONE recursion (ozaki_factinv) timed under 'trtri', with w = X z and alpha = X^T w",
            "fp64_equiv_TFLOP/s": float(n)**3 / t3 / 1e9, "frac_of_dgemm": float(n)**3 / t3 / 1e9 / dgemm,
            "lauum_fp64_equiv_TFLOP/s": float(n)**3 / 3.0 / st["lauum"] / 1e9,
            "clocks": B.sampler.summary(t0, t1),
            "nll": float(res[0]), "grad": [float(res[1]), float(res[2])],
            "nll_rel_diff_vs_dmma_path": abs(float(res[0]) - float(head_res[0])) / abs(float(head_res[0])),
            "grad_rel_diff_vs_dmma_path": rel(res),
            "parity": "tests/test_gpu_ozaki.py::test_full_size_gradient_with_all_stages_on_the_int8_pipe: this route against the CPU golden "
                      "(tests/golden/fullsize_nll_N16384.json) at 1e-9"}
    # (c) lauum only (7 digits)
    B.ctx.set_ozaki(7)
    try:
        step(); step()
        ms = B.timed(step, 3)
        st = B.stage_times()
        res = res_d.cpu().numpy()
    finally:
        B.ctx.set_ozaki(0)
        B.free()
    out["nll_grad_lauum_int8"] = {"digits": 7, "ms_per_eval": ms, "evals_per_s": 1e3 / ms, "speedup_vs_dmma_path": head_ms / ms,
                                  "stages_ms": {k: round(v, 3) for k, v in st.items()},
                                  "lauum_fp64_equiv_TFLOP/s": float(n)**3 / 3.0 / st["lauum"] / 1e9,
                                  "nll": float(res[0]), "grad": [float(res[1]), float(res[2])],
                                  "grad_rel_diff_vs_dmma_path": rel(res)}
    return out


def leg_sweep(B, dgemm):
    """BASELINE config 5: NLL+gradient at N = 2048 ... sweep-max (n up to 65 536), device resident, with stage times."""
    a, torch, W = B.a, B.torch, B.W
    out = []
    for Ns in (2048, 4096, 8192, 32768):
        if Ns == a.n_train or Ns > a.sweep_max:
            continue
        ds = W.standard_map_training(Ns)
        hs = W.timing_hyp(Ns, ds["sig"], 1e-8)
        xs_d = torch.from_numpy(ds["xtrain"].copy()).to(B.dev)
        zs_d = torch.from_numpy(ds["ztrain"].copy()).to(B.dev)
        step_s, rs_d = B.nll_dev_fn(hs, xs_d, zs_d, 2 * Ns)
        big = Ns > 16384
        try:
            for _ in range(1 if big else 3):
                step_s()
            ms = B.timed(step_s, 2 if big else 5)
        except (MemoryError, RuntimeError) as e:
            out.append({"N": Ns, "n": 2 * Ns, "skipped": str(e)[:120]})
            B.free()
            continue
        st = B.stage_times()
        n = 2.0 * Ns
        t3 = st["potrf"] + st["trtri"] + st["lauum"]
        row = {"N": Ns, "n": 2 * Ns, "ms_per_eval": ms, "TFLOP/s": n ** 3 / ms / 1e9, "frac_of_dgemm": n ** 3 / ms / 1e9 / dgemm,
               "stages_ms": {k: round(v, 4) for k, v in st.items()}, "dmma_stage_TFLOP/s": n ** 3 / t3 / 1e9 if t3 > 0 else None,
               "nll": float(rs_d[0].item())}
        if 2 * Ns > 4096:
            # the same evaluation on the opt-in INT8 route (6 digits, all three stages)
            res_dmma = rs_d.cpu().numpy().copy()
            B.ctx.set_ozaki_ex(6, 3, 4096)
            try:
                for _ in range(1 if big else 2):
                    step_s()
                ms8 = B.timed(step_s, 2 if big else 5)
                r8 = rs_d.cpu().numpy()
                row["int8_route"] = {"ms_per_eval": ms8, "fp64_equiv_TFLOP/s": n ** 3 / ms8 / 1e9, "speedup": ms / ms8,
                                     "nll_rel_diff": abs(float(r8[0]) - float(res_dmma[0])) / abs(float(res_dmma[0])),
                                     "grad_rel_diff": max(abs(float(r8[1 + k]) - float(res_dmma[1 + k])) / abs(float(res_dmma[1 + k])) for k in range(2))}
            except Exception as e:                       # noqa: BLE001 -- an optional leg must not take the bench line with it
                row["int8_route"] = {"skipped": f"{type(e).__name__}: {e}"[:160]}
            finally:
                B.ctx.set_ozaki_ex(0, 1, 0)
        out.append(row)
        del xs_d, zs_d
        B.free()
    return out


def leg_config01(B):
    """BASELINE config 1 (01_pendulum): SE-derivative (product) kernel, 200 training pairs (n = 400): NLL+gradient latency,
    then 100 orbits x 1000 steps with the reference's solver (hybrd1; the pendulum guess GP is trained on P)."""
    a, torch, W, api = B.a, B.torch, B.W, B.api
    N = 200
    d = W.pendulum_training(N)
    hyp = W.aniso_hyp(N, d["sig"], 2 * np.pi, 5.0, 1.0, 1e-8)
    hypp = W.aniso_hyp(N, d["sigp"], 2 * np.pi, 5.0, 1.0, 1e-8)
    x_d = torch.from_numpy(d["xtrain"].copy()).to(B.dev)
    z_d = torch.from_numpy(d["ztrain"].copy()).to(B.dev)
    step, res_d = B.nll_dev_fn(hyp, x_d, z_d, 2 * N)
    for _ in range(20):
        step()
    ms_dev = B.timed(step, 200)
    st = B.stage_times()
    api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    t0 = time.perf_counter()
    for _ in range(100):
        v, g = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    ms_e2e = (time.perf_counter() - t0) * 10.0
    f = api.fit(hyp, d["xtrain"], d["ztrain"], 2 * N)
    fp = api.fit(hypp, d["xtrainp"], d["ztrainp"], N, reg=True)
    E, S = 100, 1000
    q0 = W.halton(E, 5) * 2 * np.pi
    p0 = -2.0 + 4.0 * W.halton(E, 7)
    args = (S + 1, E, hyp[:3], hypp[:3], q0, p0, d["xtrainp"], None, None, d["xtrain"], None, None)
    kw = dict(alphap=fp["alpha"], alpha=f["alpha"], out_every=0, return_stats=True)
    out = {}
    for solver in ("hybrd", "newton"):
        api.applymap(*args, solver=solver, **kw)
        t0 = time.perf_counter()
        qf, pf, stt = api.applymap(*args, solver=solver, **kw)
        dt = time.perf_counter() - t0
        out[solver] = {"orbit_steps_per_s": E * S / dt, "seconds": dt, "evaluations_per_orbit_step": stt["evaluations"] / (E * S),
                       "unconverged": stt["unconverged"], "finite": int(np.isfinite(qf).sum())}
    return {"workload": "01_pendulum: kick-drift pendulum map, product kernel, N=200 (n=400); 100 orbits x 1000 steps",
            "nll_grad": {"ms_per_eval_device_resident": ms_dev, "ms_per_eval_e2e_host_buffers": ms_e2e, "nll": v,
                         "grad": [float(g[0]), float(g[1])], "stages_ms": {k: round(v_, 4) for k, v_ in st.items()},
                         "note": "latency-bound: 4 x 4 tiles of 128 (n padded to 512), the diagonal-tile chain of potrf_ll and "
                                 "kernel-launch latencies dominate"},
            "map_e2e": out, "hyp": [float(h) for h in hyp]}


def leg_config03(B, dgemm):
    """BASELINE config 3 (03_henon_heiles as BASELINE.json states it: 2-DOF, 4 x 4-block Hessian kernel, 8k training pairs
    -> n = 32 768; not in the reference): NLL+gradient, then 1e5 orbits x 1000 steps of map4_kernel."""
    a, torch, W, api = B.a, B.torch, B.W, B.api
    N = 8192
    x, z = W.henon_like_training(N)
    hyp = None
    for shrink in (1.0, 0.8, 0.65, 0.5):              # the first length scale the fp64 Cholesky accepts
        h = W.dof2_hyp(N, z, shrink)
        try:
            api.nll_chol4(h, x, z, 4 * N)
            hyp = h
            break
        except np.linalg.LinAlgError:
            continue
    if hyp is None:
        return {"skipped": "no positive definite 2-DOF kernel matrix for the tried length scales"}
    x_d, z_d = torch.from_numpy(x.copy()).to(B.dev), torch.from_numpy(z.copy()).to(B.dev)
    step, res_d = B.nll_dev_fn(hyp, x_d, z_d, 4 * N, reg=4, fam=1)
    step()
    ms = B.timed(step, 2)
    st = B.stage_times()
    res = res_d.cpu().numpy()
    # the same evaluation on the opt-in INT8 route (all three stages).  This matrix is ill-conditioned (the largest
    # length scale the FP64 Cholesky accepts), so the digit count matters: 6, 7 and 8 digits with their distance to the DMMA result
    int8 = {}
    for nd in (6, 7, 8):
        B.ctx.set_ozaki_ex(nd, 3, 4096)
        try:
            step(); step()
            ms8 = B.timed(step, 2)
            r8 = res_d.cpu().numpy()
            int8[f"{nd}_digits"] = {"ms_per_eval": ms8, "fp64_equiv_TFLOP/s": (4.0 * N) ** 3 / ms8 / 1e9, "speedup": ms / ms8,
                                    "nll_rel_diff": abs(float(r8[0]) - float(res[0])) / abs(float(res[0])),
                                    "grad_rel_diff": max(abs(float(r8[1 + k]) - float(res[1 + k])) / abs(float(res[1 + k])) for k in range(2))}
        except Exception as e:                               # noqa: BLE001 -- an optional leg must not take the bench line with it
            int8[f"{nd}_digits"] = {"skipped": f"{type(e).__name__}: {e}"[:160]}
        finally:
            B.ctx.set_ozaki_ex(0, 1, 0)
    f = api.fit(hyp, x, z, 4 * N, reg=4)
    B.free()
    E, S = a.orbits, a.map_steps
    q0 = np.vstack((-0.3 + 0.6 * W.halton(E, 2, 7), -0.3 + 0.6 * W.halton(E, 3, 7)))
    p0 = np.vstack((-0.3 + 0.6 * W.halton(E, 5, 7), -0.3 + 0.6 * W.halton(E, 7, 7)))
    api.applymap4(3, E, hyp[:3], q0, p0, x, f["alpha"], out_every=0)
    t0 = time.perf_counter()
    qf, pf, stt = api.applymap4(S + 1, E, hyp[:3], q0, p0, x, f["alpha"], out_every=0, return_stats=True)
    t_map = time.perf_counter() - t0
    q1, p1 = api.applymap4(2, E, hyp[:3], q0, p0, x, f["alpha"], out_every=0)
    dt = 0.3
    P = np.vstack((p0[0] - dt * (q0[0] + 2 * q0[0] * q0[1]), p0[1] - dt * (q0[1] + q0[0] ** 2 - q0[1] ** 2)))
    n = 4.0 * N
    pair_rate = stt["evaluations"] * N / t_map
    return {"workload": f"03_henon_heiles: 2-DOF 4x4-block kernel, N={N} (n={4 * N}); {E} orbits x {S} steps (e2e, host buffers)",
            "parity": "unpinned (no reference code; oracle twin self-validated, DESIGN.md 1 row X1)",
            "nll_grad": {"ms_per_eval": ms, "TFLOP/s": n ** 3 / ms / 1e9, "frac_of_dgemm": n ** 3 / ms / 1e9 / dgemm,
                         "stages_ms": {k: round(v_, 3) for k, v_ in st.items()}, "nll": float(res[0]), "grad": [float(res[1]), float(res[2])],
                         "int8_route": int8},
            "map_e2e": {"orbit_steps_per_s": E * S / t_map, "seconds": t_map, "passes_per_orbit_step": stt["evaluations"] / (E * S),
                        "unconverged": stt["unconverged"], "pair_evals_per_s": pair_rate,
                        "roofline": {"bound": "fp64 pipe", "achieved": pair_rate * 40.0, "peak": fp64_pipe_peak()[0],
                                     "unit": "DP instr/s", "frac": pair_rate * 40.0 / fp64_pipe_peak()[0], "peak_how": fp64_pipe_peak()[1],
                                     "note": "~43 / 31 DP instructions per pair (Newton pass with the 2x2 Jacobian sums / dQ pass)"},
                        "one_step_error_vs_training_map": float(max(np.abs(p1 - P).max(), np.abs(q1 - (q0 + dt * P)).max())),
                        "finite_final": float(np.isfinite(qf).mean())},
            "hyp": [float(h) for h in hyp]}


def leg_map04(B):
    """BASELINE config 4 / second metric: standard-map model with 4096 training pairs, ensemble sharded over the GPUs."""
    a, torch, W, api, dist = B.a, B.torch, B.W, B.api, B.dist
    dev, world, rank, L = B.dev, B.world, B.rank, B.L
    Nt = a.map_train
    dm = W.standard_map_training(Nt)
    hm = W.timing_hyp(Nt, dm["sig"], 1e-8, factor=1.0)
    hpm = W.timing_hyp(Nt, dm["sigp"], 1e-8, factor=1.0)
    fm = api.fit(hm, dm["xtrain"], dm["ztrain"], 2 * Nt)
    fpm = api.fit(hpm, dm["xtrainp"], dm["ztrainp"], Nt, reg=True)
    E = a.orbits
    q0_all, p0_all = W.ensemble(E * world)
    q0 = torch.from_numpy(q0_all[rank::world].copy()).to(dev)      # interleaved shards (load balance)
    p0 = torch.from_numpy(p0_all[rank::world].copy()).to(dev)
    qf, pf = torch.empty_like(q0), torch.empty_like(p0)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    xtp, xt = dm["xtrainp"], dm["xtrain"]
    model = B.En.DeviceMapModel(hm[:3], hpm[:3], xtp, fpm["alpha"], xt, fm["alpha"], device=dev)

    def map_run(solver, steps):
        B._lib.check(L.sgp_model_applymap_dev(B.ctx.handle, model.handle, 2, solver, steps, E, q0.data_ptr(), p0.data_ptr(),
                                              qf.data_ptr(), pf.data_ptr(), None, None, 0, stats.data_ptr()),
                     "sgp_model_applymap_dev")

    def timed_run(solver, steps):
        B.sync_all()
        stats.zero_()
        m0, m1 = B.ev(), B.ev()
        tw0 = time.time()
        m0.record(B.stream); map_run(solver, steps); m1.record(B.stream)
        B.sync_all()
        return m0.elapsed_time(m1) * 1e-3, tw0, time.time()
    # solver 3 = Newton started at p + guess: the guess GP of this workload is trained on P - p as
    # python/04_standard_map/main.py:89-90 does (the reference starts hybrd1 at the bare difference).  Pinned at this
    # size to the oracle's hybrd1 from the same start: tests/test_gpu_parity.py::test_newton_delta_config4_golden.
    MAP_SOLVER, MAP_SOLVER_NAME = 3, "newton_delta"
    map_run(MAP_SOLVER, 8)                          # warm-up
    t_map_local, t_mwall0, t_mwall1 = timed_run(MAP_SOLVER, a.map_steps)
    t_map = B.max_over_ranks(t_map_local)
    ck = B.sampler.summary(t_mwall0, t_mwall1)
    mine_r = torch.tensor([t_map_local, float(stats[0].item()), float(stats[1].item()), ck["sm_mhz"] or 0.0,
                           1.0 if ck["reasons"] else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.zeros_like(mine_r) for _ in range(world)]
        dist.all_gather(allr, mine_r)
    else:
        allr = [mine_r]
    per_rank = {"seconds": [round(float(r[0]), 3) for r in allr], "evaluations": [int(r[1]) for r in allr],
                "unconverged": [int(r[2]) for r in allr], "sm_mhz": [float(r[3]) for r in allr],
                "throttle_reasons_seen": [bool(r[4] > 0) for r in allr]}
    st = stats.clone()
    npass = ctypes.c_ulonglong(0)
    B._lib.check(L.sgp_map_last_passes(B.ctx.handle, ctypes.byref(npass)), "sgp_map_last_passes")
    passes = B.sum_over_ranks(torch.tensor([npass.value], dtype=torch.int64, device=dev))
    cks = torch.stack([torch.nansum(qf), torch.nansum(pf)])
    B.sum_over_ranks(st)                            # the only collectives of the path: gather statistics
    B.sum_over_ranks(cks)
    evals = int(st[0].item())
    orbit_steps = float(E) * world * a.map_steps
    # the learned map follows the map it was trained on (standard map, K = 0.9): distance after 50 steps
    hsteps = min(a.map_steps, 50)
    map_run(MAP_SOLVER, hsteps)
    B.sync_all()
    qt, pt = q0.clone(), p0.clone()
    for _ in range(hsteps):
        Pn = pt + 0.9 * torch.sin(qt)
        qt = torch.remainder(qt + Pn, 2 * math.pi)
        pt = torch.remainder(Pn, 2 * math.pi)

    def wrapd(x, y):
        d_ = (x - y).abs()
        return torch.minimum(d_, (d_ - 2 * math.pi).abs())
    err = torch.maximum(wrapd(qf, qt), wrapd(pf, pt))
    fin = torch.isfinite(err)
    agree = B.sum_over_ranks(torch.stack([(fin & (err < 1e-2)).sum(), fin.sum()]).to(torch.int64))
    # the reference's own solver (MINPACK hybrd1 from the reference's start) over the FULL step count, and Newton from
    # the reference's start on a short run
    t_h, _, _ = timed_run(0, a.map_steps)
    t_map_h = B.max_over_ranks(t_h)
    st_h = B.sum_over_ranks(stats.clone())
    t_n, _, _ = timed_run(1, hsteps)
    t_map_n = B.max_over_ranks(t_n)
    # end to end through the public API: host arrays in, final states out (model staging, H2D of the initial
    # conditions and D2H of the result inside the timed region)
    q0_h, p0_h = q0_all[rank::world].copy(), p0_all[rank::world].copy()
    B.sync_all()
    t0 = time.perf_counter()
    api.applymap_standard(a.map_steps + 1, E, hm[:3], hpm[:3], q0_h, p0_h, xtp, None, None, xt, None, None,
                          solver=MAP_SOLVER_NAME, alphap=fpm["alpha"], alpha=fm["alpha"], out_every=0, want_pdiff=False)
    t_map_e2e = B.max_over_ranks(time.perf_counter() - t0)
    B.ctx.set_stream(B.stream.cuda_stream)
    pair_evals = (Nt + evals / orbit_steps * Nt)      # per orbit-step: guess sweep + solver/dq sweeps
    dp_instr = 34.0                                   # DP instructions per pair evaluation (SASS of the F/dF sweep)
    fp64_peak, fp64_how = fp64_pipe_peak()            # thread-level DP instructions/s
    info = {"metric": "orbit_map_steps_per_s", "value": orbit_steps / t_map, "unit": "orbit-steps/s",
            "solver": MAP_SOLVER_NAME, "n_train": Nt, "orbits_total": E * world, "steps": a.map_steps,
            "parity": "tests/golden/map_config4_newton_delta.npz: this solver against the oracle's hybrd1 from the same start on "
                      "orbits of THIS ensemble at this Nt over 1000 steps (test_newton_delta_config4_golden)",
            "sweeps_per_orbit_step": 1 + evals / orbit_steps,
            "passes_per_warp_step": int(passes.item()) / (orbit_steps / 32.0),
            "pair_evals_per_s": orbit_steps * pair_evals / t_map,
            "roofline": {"kernel": "map_kernel<product,newton>", "bound": "fp64 pipe",
                         "note": "useful (per-orbit) pair evaluations x 34 DP instr (SASS of the F/dF sweep); lanes that idle "
                                 "during a full pass are not counted",
                         "peak_how": fp64_how,
                         "achieved": orbit_steps * pair_evals * dp_instr / t_map / world, "peak": fp64_peak,
                         "unit": "DP instr/s per GPU", "frac": orbit_steps * pair_evals * dp_instr / t_map / world / fp64_peak},
            "e2e": {"value": orbit_steps / t_map_e2e, "unit": "orbit-steps/s",
                    "h2d_bytes_per_step": int(16 * E + 8 * (3 * Nt + 4 * Nt)), "d2h_bytes_per_step": int(16 * E)},
            "hybrd_value": orbit_steps / t_map_h, "hybrd_steps": a.map_steps,
            "hybrd_evaluations_per_orbit_step": int(st_h[0].item()) / orbit_steps,
            "hybrd_note": "MINPACK hybrd1 started at the bare guess, exactly what sympgpr.f90:103-107 does with this model "
                          "(guess GP trained on P - p): the strictly pinned solver, timed over the full step count",
            "newton_refstart_value": float(E) * world * hsteps / t_map_n,
            "follows_exact_map_after_50_steps": {"within_1e-2": int(agree[0].item()), "finite": int(agree[1].item())},
            "unconverged": int(st[1].item()),
            "clocks": B.sampler.summary(t_mwall0, t_mwall1), "rank0_s": t_map_local, "per_rank": per_rank,
            "checksum": [float(cks[0].item()), float(cks[1].item())]}
    # ---- BASELINE config 5: ONE ensemble of 10^7 orbits split over the GPUs (strong scaling) ----
    if a.big_orbits > 0:
        big_steps = a.big_steps if a.big_steps > 0 else (1000 if world >= 8 else 16)
        Eb_tot = int(a.big_orbits)
        qb_all, pb_all = W.ensemble(Eb_tot)
        qb = torch.from_numpy(qb_all[rank::world].copy()).to(dev)
        pb = torch.from_numpy(pb_all[rank::world].copy()).to(dev)
        del qb_all, pb_all
        Eb = int(qb.numel())
        qbf, pbf = torch.empty_like(qb), torch.empty_like(pb)

        def big_run(steps):
            B._lib.check(L.sgp_model_applymap_dev(B.ctx.handle, model.handle, 2, MAP_SOLVER, steps, Eb, qb.data_ptr(),
                                                  pb.data_ptr(), qbf.data_ptr(), pbf.data_ptr(), None, None, 0,
                                                  stats.data_ptr()), "sgp_model_applymap_dev")
        big_run(1)
        B.sync_all()
        stats.zero_()
        b0, b1 = B.ev(), B.ev()
        b0.record(B.stream); big_run(big_steps); b1.record(B.stream)
        B.sync_all()
        t_big_local = b0.elapsed_time(b1) * 1e-3
        t_big = B.max_over_ranks(t_big_local)
        stb = B.sum_over_ranks(stats.clone())
        tl = torch.tensor([t_big_local], dtype=torch.float64, device=dev)
        if world > 1:
            tls = [torch.zeros_like(tl) for _ in range(world)]
            dist.all_gather(tls, tl)
        else:
            tls = [tl]
        info["big"] = {"metric": "orbit_map_steps_per_s", "value": float(Eb_tot) * big_steps / t_big,
                       "unit": "orbit-steps/s", "scaling": "strong", "orbits_total": Eb_tot, "steps": big_steps,
                       "solver": MAP_SOLVER_NAME, "n_train": Nt,
                       "sweeps_per_orbit_step": 1 + int(stb[0].item()) / (float(Eb_tot) * big_steps),
                       "unconverged": int(stb[1].item()), "per_rank_s": [round(float(t[0]), 3) for t in tls],
                       "note": "BASELINE config 5: one 10^7-orbit ensemble split over the GPUs (interleaved shards, model replicated, "
                               "no collective on the data path).  1000 steps on 8 GPUs; on fewer GPUs a stated reduced step count "
                               "(1000 steps of 10^7 orbits take ~330 s on one GPU) -- the rate per orbit-step is what compares"}
        del qb, pb, qbf, pbf
    model.close()
    return info


def leg_config05(B):
    """BASELINE config 5 (05_tokamak): learned field-line map with --tok-train training pairs (default 16 384, n = 32 768),
    (a) the tokamak map kind with its loss test on 1e5 orbits per GPU, (b) the Sobol sample set of --sobol-samples rows
    split over the GPUs through sympgpr_b200.ensemble.sobol_indices_sharded(on_device=True): d + 2 = 4 map ensembles per
    row block, estimator sums accumulated on the device, ONE all_reduce (NCCL) of 10 doubles."""
    a, torch, W, api = B.a, B.torch, B.W, B.api
    dev, world, rank = B.dev, B.world, B.rank
    Nt = a.tok_train
    d = W.tokamak_training(Nt)
    hyp = hypp = None
    for factor in (1.0, 0.8, 0.65, 0.5):
        h = W.aniso_hyp(Nt, d["sig"], 2 * np.pi, 9.4, factor, 1e-8)
        hp = W.aniso_hyp(Nt, d["sigp"], 2 * np.pi, 9.4, factor, 1e-8)
        try:
            f = api.fit(h, d["xtrain"], d["ztrain"], 2 * Nt)
            fp = api.fit(hp, d["xtrainp"], d["ztrainp"], Nt, reg=True)
            hyp, hypp = h, hp
            break
        except np.linalg.LinAlgError:
            continue
    if hyp is None:
        return {"skipped": "tokamak model: kernel matrix not positive definite for the tried length scales"}
    B.free()
    B.ctx.set_stream(B.stream.cuda_stream)
    model = B.En.DeviceMapModel(hyp[:3], hypp[:3], d["xtrainp"], fp["alpha"], d["xtrain"], f["alpha"], device=dev)
    # (a) tokamak map kind: q wrapped, orbit lost (NaN) where compute_r(...) > 0.5 or P < 0 (05_tokamak/SympGPR/func.py:182-211)
    E, S = a.orbits, 200
    q0_all = W.halton(E * world, 5) * 2 * np.pi
    p0_all = 0.2 + 10.3 * W.halton(E * world, 7)
    q0 = torch.from_numpy(q0_all[rank::world].copy()).to(dev)
    p0 = torch.from_numpy(p0_all[rank::world].copy()).to(dev)
    model.applymap(q0, p0, 4, kind="tokamak", solver="newton_delta")
    B.sync_all()
    model.stats.zero_()
    m0, m1 = B.ev(), B.ev()
    m0.record(B.stream)
    qf, pf = model.applymap(q0, p0, S, kind="tokamak", solver="newton_delta")
    m1.record(B.stream)
    B.sync_all()
    t_tok = B.max_over_ranks(m0.elapsed_time(m1) * 1e-3)
    lost = B.sum_over_ranks(torch.isnan(pf).sum().to(torch.int64).reshape(1))
    st = B.sum_over_ranks(model.stats.clone())
    # one step against the map the model was trained on
    q1, p1 = model.applymap(q0, p0, 1, kind="tokamak", solver="newton_delta")
    qe, pe = W.tokamak_exact_step(q0.cpu().numpy(), p0.cpu().numpy(), d["par"])
    dq = np.abs(q1.cpu().numpy() - qe)
    dq = np.minimum(dq, np.abs(dq - 2 * np.pi))
    inside = (p0.cpu().numpy() > 0.5) & (p0.cpu().numpy() < 9.5) & np.isfinite(p1.cpu().numpy())
    one_step = float(max(dq[inside].max(), np.abs(p1.cpu().numpy() - pe)[inside].max())) if inside.any() else None
    tok = {"metric": "orbit_map_steps_per_s", "value": float(E) * world * S / t_tok, "unit": "orbit-steps/s", "kind": "tokamak",
           "solver": "newton_delta", "n_train": Nt, "orbits_total": E * world, "steps": S,
           "lost_orbits": int(lost.item()), "sweeps_per_orbit_step": 1 + int(st[0].item()) / (float(E) * world * S),
           "unconverged": int(st[1].item()), "one_step_error_vs_training_map": one_step,
           "note": "lost orbits stop costing sweeps once NaN; orbit-steps are counted for all orbits as the reference loop does"}
    # (b) Sobol indices of the action after S_s steps w.r.t. the initial conditions (theta0, p0)
    S_s = a.sobol_steps
    nrows = int(a.sobol_samples)

    def model_fn(X):
        qx, px = model.applymap(X[:, 0].contiguous(), X[:, 1].contiguous(), S_s, kind="tokamak", solver="newton_delta")
        return px
    bounds = [(0.0, 2 * np.pi), (0.5, 9.0)]
    B.En.sobol_indices_sharded(model_fn, bounds, min(nrows, 4096 * world), device=dev, on_device=True)      # warm-up
    B.sync_all()
    s0, s1 = B.ev(), B.ev()
    s0.record(B.stream)
    r = B.En.sobol_indices_sharded(model_fn, bounds, nrows, device=dev, on_device=True, block=1 << 18)
    s1.record(B.stream)
    B.sync_all()
    t_sob = B.max_over_ranks(s0.elapsed_time(s1) * 1e-3)
    sob = {"metric": "sobol_samples_per_s", "value": nrows / t_sob, "unit": "rows/s", "rows": nrows, "inputs": ["theta0", "p0"],
           "output": f"action P after {S_s} map steps (lost orbits dropped)", "model_runs": r["model_runs"],
           "orbit_steps_per_s": r["model_runs"] * S_s / t_sob, "S1": [float(v) for v in r["S1"]], "ST": [float(v) for v in r["ST"]],
           "mean": r["mean"], "var": r["var"], "rows_used": r["n_used"], "seconds": t_sob, "n_train": Nt,
           "collectives": "one all_reduce(sum) of 10 doubles (+ one int flag) per call; sample rows generated on the device",
           "parity": "unpinned (no reference code, SURVEY 8a row X2); estimator validated against the analytic Ishigami indices "
                     "on this device path (tests/test_gpu_parity.py::test_sobol_on_device_ishigami)"}
    model.close()
    B.free()
    return {"workload": f"05_tokamak: field-line-like twist map, product kernel, Nt={Nt} (n={2 * Nt})", "hyp": [float(h) for h in hyp],
            "map": tok, "sobol": sob}


def leg_cpu(B, head, sweep):
    """cpu_baseline of the headline metric (rank 0, single-GPU runs): measured points at bounded sizes."""
    a = B.a
    with _all_threads():
        pts = []
        for Ns in (1024, 2048, 4096):
            if Ns > a.n_train:
                break
            pts.append((2 * Ns, min(cpu_nll_grad_time(Ns, 2 if Ns <= 2048 else 1, 1 if Ns <= 1024 else 0))))
        cores = blas_threads()
    a_c, b_c = fit_cubic(pts)
    nf = 2.0 * a.n_train
    t_full = a_c * nf**3 + b_c * nf**2
    gpu_ms = {s["n"]: s["ms_per_eval"] for s in (sweep or []) if "ms_per_eval" in s}
    like = [{"n": int(n_), "cpu_seconds": float(t_), "gpu_ms": gpu_ms.get(int(n_)),
             "speedup": (t_ * 1e3 / gpu_ms[int(n_)]) if int(n_) in gpu_ms else None} for n_, t_ in pts]
    n_l, t_l = pts[-1]
    return {"value": 1.0 / t_l, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle nll_grad (port of python/02_pert_pendulum/func.py:148-162) on the host cores, MEASURED at "
                      f"N={int(n_l // 2)} (n={int(n_l)}): {t_l:.2f} s per evaluation; value is that measurement (a bounded "
                      f"sample, 1/{(nf / n_l) ** 3:.0f} of the headline size's n^3 flops)",
            "measured_points": like,
            "extrapolated": {"n_train": a.n_train, "seconds_per_eval": t_full, "value": 1.0 / t_full,
                             "how": f"t(n) = a n^3 + b n^2 through the measured points (a={a_c:.3e}, b={b_c:.3e}); NOT a measurement",
                             "fullsize_record": fullsize_record()}}


def main():
    global _JSON_FD
    a = parse()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the duration of the run
    if a.impl == "reference":
        run_reference(a)
        return
    B = Bench(a)
    torch, dist, world, rank = B.torch, B.dist, B.world, B.rank
    dgemm_tflops, write_gbs = leg_peaks(B)
    B.sampler.start()
    head = leg_headline(B)
    n = head["n"]
    single = world == 1
    ozaki = None
    if single and not a.no_configs and a.route == "dmma":
        try:
            ozaki = leg_ozaki(B, head["ms_per_step"], head["res"], dgemm_tflops, read_bf16_peak())
        except Exception as e:                           # noqa: BLE001 -- the opt-in leg must not take the headline line with it
            ozaki = {"skipped": f"{type(e).__name__}: {e}"[:200]}
            B.ctx.set_ozaki_ex(0, 1, 0)
            B.free()
    sweep = leg_sweep(B, dgemm_tflops) if (single and not a.no_sweep) else None
    configs = {}
    if not a.no_configs:
        if single:
            configs["01_pendulum"] = leg_config01(B)
            configs["03_henon_heiles_2dof"] = leg_config03(B, dgemm_tflops)
        configs["05_tokamak"] = leg_config05(B)
    B.ctx.set_stream(B.stream.cuda_stream)
    map_info = leg_map04(B) if not a.no_map else None
    B.sampler.stop()
    clocks = B.sampler.summary(*head["wall"])

    cpu = None
    if rank == 0 and single and not a.no_cpu_baseline:
        cpu = leg_cpu(B, head, sweep)
        if map_info is not None:
            with _all_threads():
                map_info["cpu_baseline"] = cpu_map_sample(a.map_train, 512, 40)
        if "01_pendulum" in configs:
            with _all_threads():
                ts = cpu_nll_grad_small(200)
            configs["01_pendulum"]["cpu_baseline"] = ts

    if rank == 0:
        stages = head["stages"]
        t_gemm = (stages["potrf"] + stages["trtri"] + stages["lauum"]) * 1e-3
        achieved = float(n) ** 3 / t_gemm / 1e12
        hbm = read_hbm_peak()
        tf = lambda fl, ms: fl / (ms * 1e-3) / 1e12 if ms > 0 else None
        gb = lambda by, ms: by / (ms * 1e-3) / 1e9 if ms > 0 else None
        stage_roof = {
            "fill_sym": {"bound": "hbm write", "bytes": 4.0 * n * n, "ms": stages["fill"], "GB/s": gb(4.0 * n * n, stages["fill"]),
                         "frac_of_write_peak": (gb(4.0 * n * n, stages["fill"]) or 0.0) / write_gbs},
            "potrf": {"bound": "tensor", "flops": n ** 3 / 3.0, "ms": stages["potrf"], "TFLOP/s": tf(n ** 3 / 3.0, stages["potrf"]),
                      "note": "includes the forward substitution L w = z (fused into the diagonal tasks)"},
            "potrs": {"ms": stages["potrs"], "note": "empty stage: the forward half rides in potrf, alpha = X^T w is timed with trtri"},
            "trtri": {"bound": "tensor", "flops": n ** 3 / 3.0, "ms": stages["trtri"], "TFLOP/s": tf(n ** 3 / 3.0, stages["trtri"])},
            "lauum": {"bound": "tensor", "flops": n ** 3 / 3.0, "ms": stages["lauum"], "TFLOP/s": tf(n ** 3 / 3.0, stages["lauum"])},
            "grad": {"bound": "hbm read", "bytes": 4.0 * n * n, "ms": stages["grad"], "GB/s": gb(4.0 * n * n, stages["grad"]),
                     "frac_of_hbm_peak": (gb(4.0 * n * n, stages["grad"]) or 0.0) / hbm},
        }
        traffic, traffic_src = read_traffic(n)
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(a),
                "e2e": {"value": head["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": head["d2h"]},
                "gpu_launches": head["launches"], "clocks": clocks,
                "roofline": {"kernel": "potrf_ll_kernel + gemm_f64_ws_kernel (DMMA m8n8k4: potrf 1 launch, trtri 2 log2(n/128) + 1 "
                                       "grouped launches, lauum 1 launch; taken together)",
                             "bound": "tensor", "achieved": achieved, "peak": dgemm_tflops, "unit": "TFLOP/s",
                             "frac": achieved / dgemm_tflops, "traffic": traffic,
                             "note": "achieved = n^3 algorithmic flops / (potrf+trtri+lauum stage time of the last timed "
                                     "step, CUDA events on the launch stream); peak = cuBLAS DGEMM 8192^3 measured in "
                                     "this run (MEASURED_PEAKS.json has no fp64 entry; DMMA issue limit 37.2); traffic = "
                                     f"DRAM bytes read + written by those kernels in one evaluation (ncu, {traffic_src}), "
                                     "against 3 x 8 n^2 / 2 algorithmic matrix bytes: operand panels are re-streamed per "
                                     "tile task, far below the HBM peak"},
                "stages": stage_roof,
                "roofline_fill": {"kernel": "fill_hess_kernel", "bound": "hbm write", "achieved": head["fill_gbs"],
                                  "peak": write_gbs, "unit": "GB/s", "frac": head["fill_gbs"] / write_gbs,
                                  "bytes": 8.0 * n * n, "ms": head["t_fill"] * 1e3,
                                  "peak_note": "write-only peak measured in this run (memset of 8 GiB, best of 3); the read+write "
                                               f"copy peak of MEASURED_PEAKS.json is {hbm} GB/s, frac against it "
                                               f"{head['fill_gbs'] / hbm:.3f} -- a store-only kernel can exceed a copy figure",
                                  "frac_of_copy_peak": head["fill_gbs"] / hbm},
                "peaks": {"dgemm_tflops": dgemm_tflops, "hbm_write_gbs": write_gbs, "hbm_copy_gbs": hbm,
                          "dfma_dp_instr_per_s": FP64_PEAK_MEASURED},
                "result": {"nll": float(head["res"][0]), "grad": [float(head["res"][1]), float(head["res"][2])]},
                }
        if a.route == "int8":
            # the opt-in route as the headline: the products are INT8 digit-pair GEMMs, so the roofline is the INT8 tensor rate
            nd = a.int8_digits
            pairs = nd * (nd + 1) // 2
            t3 = (stages["potrf"] + stages["trtri"] + stages["lauum"]) * 1e-3
            tops = pairs * float(n) ** 3 / t3 / 1e12
            peak = 2.0 * read_bf16_peak(sustained=True)
            line["dtype"] = f"f64 from {nd} int8 digits per operand ({8 * nd - 1} bits of every row's scale)"
            line["config"]["route"] = f"int8 (opt-in, sgp_set_ozaki_ex({nd}, 3, 4096)); the default route is dmma"
            line["roofline"] = {"kernel": f"oz_gemm_kernel<{nd}> (all sliced products of one evaluation; leaf blocks on DMMA included in the time)",
                                "bound": "tensor", "achieved": tops, "peak": peak, "unit": "INT8 TOP/s", "frac": tops / peak, "traffic": None,
                                "fp64_equiv_TFLOP/s": float(n) ** 3 / t3 / 1e12, "frac_of_dgemm": float(n) ** 3 / t3 / 1e12 / dgemm_tflops,
                                "note": f"achieved = {pairs} digit-pair products x n^3 algorithmic operations / (factor + inverse + lauum stage "
                                        "time); peak = 2 x bf16_tflops_sustained of MEASURED_PEAKS.json (no INT8 entry there; nominal INT8 = "
                                        "2 x bf16; the sustained figure because the kernels run inside a 0.3 s step under the power cap)"}
        if ozaki:
            line["ozaki_opt_in"] = ozaki
        if sweep:
            line["sweep"] = sweep
        if configs:
            line["configs"] = configs
        if map_info:
            line["map"] = map_info
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_nll_grad_small(N):
    """Config 1 on the host: the oracle's NLL+gradient and SciPy's dpotrf alone at N = 200 (n = 400), measured."""
    import scipy.linalg
    from oracle import c_oracle as C, oracle as O
    from sympgpr_b200 import workloads as W
    d = W.pendulum_training(N)
    hyp = W.aniso_hyp(N, d["sig"], 2 * np.pi, 5.0, 1.0, 1e-8)
    O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    t0 = time.perf_counter()
    for _ in range(10):
        O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    t_ng = (time.perf_counter() - t0) / 10
    K = C.build_k(d["xtrain"][:N], d["xtrain"][N:], d["xtrain"][:N], d["xtrain"][N:], hyp[:3])
    K[np.diag_indices_from(K)] += hyp[3]
    scipy.linalg.cholesky(K, lower=True)
    t0 = time.perf_counter()
    for _ in range(50):
        L_ = scipy.linalg.cholesky(K, lower=True)
        Ki = scipy.linalg.lapack.dpotri(L_, lower=1)[0]
    t_ch = (time.perf_counter() - t0) / 50
    return {"nll_grad_ms_per_eval": 1e3 * t_ng, "scipy_potrf_potri_ms": 1e3 * t_ch, "cores": blas_threads(), "kind": "port",
            "sample": "oracle nll_grad at N=200 (10 evaluations) and SciPy dpotrf + dpotri of the same 400 x 400 matrix (50 x), measured"}


def read_traffic(n):
    """DRAM bytes of the DMMA kernels per evaluation from the committed ncu capture of this matrix order, else None."""
    for rnd in ("r02", "r01"):
        try:
            with open(os.path.join(ROOT, "profiles", f"{rnd}_traffic_n{n}.json")) as f:
                return float(json.load(f)["dmma_dram_bytes_per_evaluation"]), f"profiles/{rnd}_traffic_n{n}.json"
        except Exception:
            continue
    return None, "no capture committed for this order"


def read_bf16_peak(sustained=False):
    """Measured cuBLAS bf16 rate (TFLOP/s; burst or sustained) of this pool's B200s, else the fallback B200_PROFILING.md states."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained" if sustained else "bf16_tflops"])
    except Exception:
        return 1400.0 if sustained else 1590.0


def read_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0          # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    main()
