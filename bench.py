#!/usr/bin/env python3
"""Benchmark of the SympGPR hot path on B200 (driver contract: one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL)

Headline metric (BASELINE.json): NLL+gradient evaluations per second at N = 16 384 training
pairs (n = 32 768, fp64), synthetic standard map K = 0.9, product kernel.  One "step" = one
NLL+gradient evaluation: fill -> potrf -> potrs -> trtri -> lauum -> gradient contraction.
With N > 1 GPUs every rank evaluates its own multi-start restart (a different theta; the
training Cholesky does not shard -- DESIGN.md "Multi-GPU") and its own slice of the orbit
ensemble; NCCL only gathers the results ("scaling": "weak").

Also reported on the same line: the map leg (orbit map-steps/s, BASELINE's second metric), the
Hessian-block fill against the HBM roofline, the in-run cuBLAS DGEMM FP64 peak, and the CPU
baseline (oracle = port of the reference maths, SciPy LAPACK on the host cores).
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
    ap.add_argument("--cpu-sample", type=int, default=2048, help="training pairs of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-map", action="store_true")
    ap.add_argument("--big-orbits", type=int, default=10_000_000,
                    help="TOTAL orbits of the 10^7-orbit prediction leg (BASELINE config 5), split over the GPUs; 0 skips it")
    ap.add_argument("--big-steps", type=int, default=16, help="map steps of the 10^7-orbit leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the small-size NLL+gradient sweep (N = 2048..8192)")
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


# ------------------------------------------------------------------------------------------ reference arm
def cpu_nll_grad_sample(N_s, N_full, steps, warmup):
    """The reference maths on the host cores (oracle port; SciPy/OpenBLAS LAPACK with all threads): NLL+gradient
    evaluations at N_s and N_s/2 pairs, extrapolated to the full size with  t(n) = a n^3 + b n^2  fitted to the two
    samples (potrf + potri are n^3, fill / dK / contraction n^2 with a large constant in NumPy).  Scaling one sample by
    (n/n_s)^3 alone would overstate the CPU time threefold: here 7.0 s at N = 2048 -> 3580 s, the two-point fit from
    N = 2048 / 4096 gives 640 s, and the blocked full-size evaluation behind tests/golden/fullsize_nll_N16384.json
    (which does ~2x the LAPACK flops) took 1147 s on 8 cores.  Returns (evals/s at N_full, seconds at N_s, description)."""
    from oracle import oracle as O

    def timed(N, nrep, nwarm):
        d = O.standard_map_training(N)
        hyp = O.timing_hyp(N, d["sig"], 1e-8)
        for _ in range(nwarm):
            O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
        ts = []
        for _ in range(nrep):
            t = time.perf_counter()
            O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
            ts.append(time.perf_counter() - t)
        return float(np.mean(ts))

    t1 = timed(N_s, steps, warmup)
    N_h = max(64, N_s // 2)
    t0 = timed(N_h, max(1, steps), 1)
    n1, n0, nf = 2.0 * N_s, 2.0 * N_h, 2.0 * N_full
    # a n1^3 + b n1^2 = t1,  a n0^3 + b n0^2 = t0
    det = n1 ** 3 * n0 ** 2 - n0 ** 3 * n1 ** 2
    a_c = (t1 * n0 ** 2 - t0 * n1 ** 2) / det
    b_c = (n1 ** 3 * t0 - n0 ** 3 * t1) / det
    if N_h < N_s and a_c > 0 and b_c >= 0:
        t_full = a_c * nf ** 3 + b_c * nf ** 2
        how = (f"t(n) = a n^3 + b n^2 fitted to N={N_h} ({t0:.3f} s) and N={N_s} ({t1:.3f} s): a={a_c:.3e}, b={b_c:.3e}")
    else:
        t_full = t1 * (nf / n1) ** 3
        how = f"N={N_s} ({t1:.3f} s) scaled by (n/n_s)^3 (two-point fit degenerate)"
    return 1.0 / t_full, t1, how


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, a.steps), max(1, min(a.warmup, 1))
    ncpu = os.cpu_count() or 1
    try:                                    # torchrun exports OMP_NUM_THREADS=1: lift the BLAS/OpenMP limit again
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=ncpu):
            v, t_s, how = cpu_nll_grad_sample(a.cpu_sample, a.n_train, steps, warm)
            cores = blas_threads()
    except ImportError:
        v, t_s, how = cpu_nll_grad_sample(a.cpu_sample, a.n_train, steps, warm)
        cores = blas_threads()
    sample = (f"oracle nll_grad (C/NumPy fill, SciPy potrf+potri, elementwise contraction) on the host cores, extrapolated to "
              f"N={a.n_train}: {how}")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(a),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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


def main():
    global _JSON_FD
    a = parse()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the duration of the run
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist

    from sympgpr_b200 import _lib, api, workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available() or _lib.device_count() < 1:
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = _lib.lib()
    ctx = _lib.context(local)
    # One explicit (non-default) stream for everything this process launches: the library's
    # kernels, cuBLAS for the peak probe and the CUDA events that time them.  (The legacy default
    # stream has handle 0, which sgp_set_stream reads as "use your own stream".)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    _lib.check(L.sgp_set_profiling(ctx.handle, 1), "sgp_set_profiling")

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N = a.n_train
    n = 2 * N
    d = W.standard_map_training(N)
    # multi-start restart of this rank: a different length-scale pair per GPU
    hyp = W.timing_hyp(N, d["sig"], 1e-8)
    hyp[:2] *= (1.0 + 0.02 * rank)
    hyp_c = (ctypes.c_double * 4)(*hyp)
    x_h = torch.from_numpy(d["xtrain"].copy()).pin_memory()
    z_h = torch.from_numpy(d["ztrain"].copy()).pin_memory()
    x_d = x_h.to(dev)
    z_d = z_h.to(dev)
    res_d = torch.zeros(16, dtype=torch.float64, device=dev)

    def step_dev():
        _lib.check(L.sgp_nll_dev(ctx.handle, 0, 0.5, 0, hyp_c, x_d.data_ptr(), z_d.data_ptr(), n, 2, res_d.data_ptr()),
                   "sgp_nll_dev")

    # ---- FP64 tensor peak of this GPU, measured in-run: cuBLAS DGEMM 8192^3 -----------------
    A = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    B = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(A, B)
    best = 1e9
    for _ in range(4):
        e0, e1 = ev(), ev()
        e0.record(stream); torch.matmul(A, B); e1.record(stream); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    dgemm_tflops = 2 * 8192 ** 3 / best / 1e12
    del A, B
    torch.cuda.empty_cache()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident NLL+grad: W warm-up, K timed steps ---------------------------------
    for _ in range(max(a.warmup, 1)):
        step_dev()
    sync_all()
    launches0 = L.sgp_launch_count()
    t_wall0 = time.time()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(a.steps):
        step_dev()
    e1.record(stream)
    sync_all()
    t_wall1 = time.time()
    launches = int(L.sgp_launch_count() - launches0)
    t_dev = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    stage_ms = (ctypes.c_double * 7)()
    _lib.check(L.sgp_stage_times(ctx.handle, stage_ms), "sgp_stage_times")      # of the last timed step
    stages = dict(zip(["fill", "potrf", "potrs", "trtri", "lauum", "grad", "finalize"], [float(v) for v in stage_ms]))
    res = res_d.cpu().numpy()
    if res[4] != 0:
        raise RuntimeError(f"Cholesky failed in the timed region (info={res[4]})")
    value = world * a.steps / t_dev
    ms_per_step = 1e3 * t_dev / a.steps

    # ---- end to end through the public API with host buffers --------------------------------
    xh_np, zh_np = x_h.numpy(), z_h.numpy()
    api.nll_grad(hyp, xh_np, zh_np, n)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        v_e2e, g_e2e = api.nll_grad(hyp, xh_np, zh_np, n)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * a.steps / t_e2e
    h2d = int(x_h.numel() * 8 + z_h.numel() * 8)
    d2h = 16 * 8

    # ---- Hessian-block fill against the HBM roofline (full 2N x 2N matrix, 8 n^2 bytes) ------
    Kbuf = torch.empty(n * n, dtype=torch.float64, device=dev)
    hyp3 = (ctypes.c_double * 3)(*hyp[:3])
    xq, yP = x_d[:N], x_d[N:]

    def fill_full():
        _lib.check(L.sgp_build_k_dev(ctx.handle, 0, 0.5, xq.data_ptr(), yP.data_ptr(), N, xq.data_ptr(), yP.data_ptr(), N,
                                     hyp3, Kbuf.data_ptr(), n), "sgp_build_k_dev")
    for _ in range(3):
        fill_full()
    torch.cuda.synchronize()
    f0, f1 = ev(), ev()
    f0.record(stream)
    for _ in range(5):
        fill_full()
    f1.record(stream); f1.synchronize()
    t_fill = f0.elapsed_time(f1) * 1e-3 / 5
    fill_gbs = 8.0 * n * n / t_fill / 1e9
    del Kbuf
    ctx.release_workspace()
    torch.cuda.empty_cache()

    # ---- size sweep (BASELINE config 5), small sizes only: N = 2048 .. 8192 on rank 0 of a single-GPU run ------
    sweep = None
    if world == 1 and not a.no_sweep:
        sweep = []
        for Ns in (2048, 4096, 8192):
            if Ns >= N:
                continue
            ds = W.standard_map_training(Ns)
            hs = W.timing_hyp(Ns, ds["sig"], 1e-8)
            hs_c = (ctypes.c_double * 4)(*hs)
            xs_d = torch.from_numpy(ds["xtrain"].copy()).to(dev)
            zs_d = torch.from_numpy(ds["ztrain"].copy()).to(dev)
            rs_d = torch.zeros(16, dtype=torch.float64, device=dev)

            def step_s():
                _lib.check(L.sgp_nll_dev(ctx.handle, 0, 0.5, 0, hs_c, xs_d.data_ptr(), zs_d.data_ptr(), 2 * Ns, 2, rs_d.data_ptr()),
                           "sgp_nll_dev")
            for _ in range(3):
                step_s()
            torch.cuda.synchronize()
            s0, s1 = ev(), ev()
            s0.record(stream)
            for _ in range(5):
                step_s()
            s1.record(stream); s1.synchronize()
            ms = s0.elapsed_time(s1) / 5
            sweep.append({"N": Ns, "n": 2 * Ns, "ms_per_eval": ms, "TFLOP/s": (2.0 * Ns) ** 3 / ms / 1e9,
                          "nll": float(rs_d[0].item())})
        ctx.release_workspace()
        torch.cuda.empty_cache()

    # ---- map leg: ensemble sharded over the GPUs, model replicated -----------------------------
    map_info = None
    if not a.no_map:
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
        model = ctypes.c_void_p()
        dp = _lib.dptr
        xtp, xt = dm["xtrainp"], dm["xtrain"]
        _lib.check(L.sgp_model_create(ctx.handle, 0, 0.5, dp(np.ascontiguousarray(hm[:3])), dp(np.ascontiguousarray(hpm[:3])),
                                      dp(np.ascontiguousarray(xtp[:Nt])), dp(np.ascontiguousarray(xtp[Nt:])), dp(fpm["alpha"]),
                                      Nt, dp(np.ascontiguousarray(xt[:Nt])), dp(np.ascontiguousarray(xt[Nt:])), dp(fm["alpha"]),
                                      Nt, ctypes.byref(model)), "sgp_model_create")

        def map_run(solver, steps):
            _lib.check(L.sgp_model_applymap_dev(ctx.handle, model, 2, solver, steps, E, q0.data_ptr(), p0.data_ptr(),
                                                qf.data_ptr(), pf.data_ptr(), None, None, 0, stats.data_ptr()),
                       "sgp_model_applymap_dev")
        # solver 3 = Newton started at p + guess: the guess GP of this workload is trained on P - p as
        # python/04_standard_map/main.py:89-90 does (the reference starts hybrd1 at the bare difference)
        MAP_SOLVER, MAP_SOLVER_NAME = 3, "newton_delta"
        map_run(MAP_SOLVER, 8)                          # warm-up
        sync_all()
        stats.zero_()
        m0, m1 = ev(), ev()
        t_mwall0 = time.time()
        m0.record(stream); map_run(MAP_SOLVER, a.map_steps); m1.record(stream)
        sync_all()
        t_mwall1 = time.time()
        t_map_local = m0.elapsed_time(m1) * 1e-3
        t_map = max_over_ranks(t_map_local)
        # per-rank evidence for the scaling figure: time, lane-level evaluations, unconverged steps and the median SM clock
        # of every rank's own GPU during its map run (a slow rank is either a slow GPU or a slow shard)
        ck = sampler.summary(t_mwall0, t_mwall1)
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
        _lib.check(L.sgp_map_last_passes(ctx.handle, ctypes.byref(npass)), "sgp_map_last_passes")
        passes = torch.tensor([npass.value], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(passes, op=dist.ReduceOp.SUM)
        cks = torch.stack([torch.nansum(qf), torch.nansum(pf)])
        if world > 1:                                   # the only collectives of the path: gather statistics
            dist.all_reduce(st, op=dist.ReduceOp.SUM)
            dist.all_reduce(cks, op=dist.ReduceOp.SUM)
        evals = int(st[0].item())
        orbit_steps = float(E) * world * a.map_steps
        # the reference's own solver (MINPACK hybrd1) for comparison, on a shorter run
        hsteps = min(a.map_steps, 50)
        stats.zero_()
        h0, h1 = ev(), ev()
        h0.record(stream); map_run(0, hsteps); h1.record(stream)
        sync_all()
        t_map_h = max_over_ranks(h0.elapsed_time(h1) * 1e-3)
        # Newton from the reference's start (solver 1) on the same short run
        n0, n1 = ev(), ev()
        n0.record(stream); map_run(1, hsteps); n1.record(stream)
        sync_all()
        t_map_n = max_over_ranks(n0.elapsed_time(n1) * 1e-3)
        # full-size sanity property: the learned map follows the map it was trained on (standard map, K = 0.9);
        # distance of the headline solver's orbits from the exact map after hsteps steps
        map_run(MAP_SOLVER, hsteps)
        sync_all()
        qt, pt = q0.clone(), p0.clone()
        for _ in range(hsteps):
            Pn = pt + 0.9 * torch.sin(qt)
            qt = torch.remainder(qt + Pn, 2 * math.pi)
            pt = torch.remainder(Pn, 2 * math.pi)
        def wrapd(x, y):
            d = (x - y).abs()
            return torch.minimum(d, (d - 2 * math.pi).abs())
        err = torch.maximum(wrapd(qf, qt), wrapd(pf, pt))
        fin = torch.isfinite(err)
        agree = torch.stack([(fin & (err < 1e-2)).sum(), fin.sum()]).to(torch.int64)
        if world > 1:
            dist.all_reduce(agree, op=dist.ReduceOp.SUM)
        # end to end through the public API: host arrays in, final states out (model staging, H2D of the
        # initial conditions and D2H of the result inside the timed region)
        q0_h, p0_h = q0_all[rank::world].copy(), p0_all[rank::world].copy()
        sync_all()
        t0 = time.perf_counter()
        out = api.applymap_standard(a.map_steps + 1, E, hm[:3], hpm[:3], q0_h, p0_h, xtp, None, None, xt, None, None,
                                    solver=MAP_SOLVER_NAME, alphap=fpm["alpha"], alpha=fm["alpha"], out_every=0, want_pdiff=False)
        t_map_e2e = max_over_ranks(time.perf_counter() - t0)
        pair_evals = (Nt + evals / orbit_steps * Nt)      # per orbit-step: guess sweep + solver/dq sweeps
        dp_instr = 34.0                                   # DP instructions per pair evaluation (SASS of the F/dF sweep)
        fp64_peak = 148 * 64 * 1.965e9                    # thread-level DP instructions/s: SMs x FP64 lanes x max SM clock
        map_info = {"metric": "orbit_map_steps_per_s", "value": orbit_steps / t_map, "unit": "orbit-steps/s",
                    "solver": MAP_SOLVER_NAME, "n_train": Nt, "orbits_total": E * world, "steps": a.map_steps,
                    "sweeps_per_orbit_step": 1 + evals / orbit_steps,
                    "passes_per_warp_step": int(passes.item()) / (orbit_steps / 32.0),
                    "passes_note": "full passes over a training set per 32-orbit step (guess + lock-step residual passes + dQ); a "
                                   "group of cooperative single-orbit passes for the last <= 16 lanes counts as one",
                    "pair_evals_per_s": orbit_steps * pair_evals / t_map,
                    "roofline": {"kernel": "map_kernel<product,newton>", "bound": "fp64 pipe", "note": "useful (per-orbit) pair evaluations x 34 DP instr (SASS of the F/dF sweep); lanes that idle during a full pass are not counted",
                                 "achieved": orbit_steps * pair_evals * dp_instr / t_map / world, "peak": fp64_peak,
                                 "unit": "DP instr/s per GPU", "frac": orbit_steps * pair_evals * dp_instr / t_map / world / fp64_peak},
                    "e2e": {"value": orbit_steps / t_map_e2e, "unit": "orbit-steps/s",
                            "h2d_bytes_per_step": int(16 * E + 8 * (3 * Nt + 4 * Nt)), "d2h_bytes_per_step": int(16 * E)},
                    "hybrd_value": float(E) * world * hsteps / t_map_h, "hybrd_steps": hsteps,
                    "newton_refstart_value": float(E) * world * hsteps / t_map_n,
                    "follows_exact_map_after_hsteps": {"within_1e-2": int(agree[0].item()), "finite": int(agree[1].item())},
                    "unconverged": int(st[1].item()),
                    "clocks": sampler.summary(t_mwall0, t_mwall1), "rank0_s": t_map_local, "per_rank": per_rank,
                    "checksum": [float(cks[0].item()), float(cks[1].item())]}
        # ---- BASELINE config 5: ONE ensemble of 10^7 orbits split over the GPUs (strong scaling), a short step loop ----
        if a.big_orbits > 0:
            Eb_tot = int(a.big_orbits)
            qb_all, pb_all = W.ensemble(Eb_tot)
            qb = torch.from_numpy(qb_all[rank::world].copy()).to(dev)
            pb = torch.from_numpy(pb_all[rank::world].copy()).to(dev)
            del qb_all, pb_all
            Eb = int(qb.numel())
            qbf, pbf = torch.empty_like(qb), torch.empty_like(pb)

            def big_run(steps):
                _lib.check(L.sgp_model_applymap_dev(ctx.handle, model, 2, MAP_SOLVER, steps, Eb, qb.data_ptr(), pb.data_ptr(),
                                                    qbf.data_ptr(), pbf.data_ptr(), None, None, 0, stats.data_ptr()),
                           "sgp_model_applymap_dev")
            big_run(1)
            sync_all()
            stats.zero_()
            b0, b1 = ev(), ev()
            b0.record(stream); big_run(a.big_steps); b1.record(stream)
            sync_all()
            t_big_local = b0.elapsed_time(b1) * 1e-3
            t_big = max_over_ranks(t_big_local)
            stb = stats.clone()
            tl = torch.tensor([t_big_local], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(stb, op=dist.ReduceOp.SUM)
                tls = [torch.zeros_like(tl) for _ in range(world)]
                dist.all_gather(tls, tl)
            else:
                tls = [tl]
            map_info["big"] = {"metric": "orbit_map_steps_per_s", "value": float(Eb_tot) * a.big_steps / t_big,
                               "unit": "orbit-steps/s", "scaling": "strong", "orbits_total": Eb_tot, "steps": a.big_steps,
                               "solver": MAP_SOLVER_NAME, "n_train": Nt,
                               "sweeps_per_orbit_step": 1 + int(stb[0].item()) / (float(Eb_tot) * a.big_steps),
                               "unconverged": int(stb[1].item()), "per_rank_s": [round(float(t[0]), 3) for t in tls],
                               "note": "BASELINE config 5: one 10^7-orbit ensemble split over the GPUs (interleaved shards, model "
                                       "replicated, no collective on the data path); the 1e5-orbit x 1000-step leg above is "
                                       "config 4 replicated per GPU (weak scaling), whose time is set by the slowest 32-orbit "
                                       "batch of a shard"}
            del qb, pb, qbf, pbf
        L.sgp_model_destroy(model)

    sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall1)

    # ---- CPU baseline (rank 0, single GPU runs only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v_cpu, t_s, how = cpu_nll_grad_sample(a.cpu_sample, N, 2, 1)
        cpu = {"value": v_cpu, "unit": UNIT, "cores": blas_threads(), "kind": "port",
               "sample": f"oracle nll_grad on the host cores, extrapolated to N={N}: {how}"}

    if rank == 0:
        # Dominant kernel: gemm_f64_kernel carries every flop of potrf/trtri/lauum (n^3/3 each).  Its
        # launches differ in shape, so the roofline is taken over all of them together: algorithmic
        # n^3 flops / (potrf + trtri + lauum stage time, CUDA events on the launch stream inside the
        # timed region).  The panel kernels (potrf_tile, copy_tile) sit inside those stages too, so
        # the figure is a lower bound of the GEMM kernel's own rate.
        t_gemm = (stages["potrf"] + stages["trtri"] + stages["lauum"]) * 1e-3
        achieved = float(n) ** 3 / t_gemm / 1e12
        hbm = read_hbm_peak()
        tf = lambda fl, ms: fl / (ms * 1e-3) / 1e12 if ms > 0 else None
        gb = lambda by, ms: by / (ms * 1e-3) / 1e9 if ms > 0 else None
        stage_roof = {
            "fill_sym": {"bound": "hbm", "bytes": 4.0 * n * n, "ms": stages["fill"], "GB/s": gb(4.0 * n * n, stages["fill"])},
            "potrf": {"bound": "tensor", "flops": n ** 3 / 3.0, "ms": stages["potrf"], "TFLOP/s": tf(n ** 3 / 3.0, stages["potrf"])},
            "potrs": {"bound": "hbm", "bytes": 8.0 * n * n, "ms": stages["potrs"], "GB/s": gb(8.0 * n * n, stages["potrs"])},
            "trtri": {"bound": "tensor", "flops": n ** 3 / 3.0, "ms": stages["trtri"], "TFLOP/s": tf(n ** 3 / 3.0, stages["trtri"])},
            "lauum": {"bound": "tensor", "flops": n ** 3 / 3.0, "ms": stages["lauum"], "TFLOP/s": tf(n ** 3 / 3.0, stages["lauum"])},
            "grad": {"bound": "hbm", "bytes": 4.0 * n * n, "ms": stages["grad"], "GB/s": gb(4.0 * n * n, stages["grad"])},
        }
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(a),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"kernel": "potrf_ll_kernel + gemm_f64_ws_kernel (DMMA m8n8k4: potrf, trtri, lauum; 512 launches per "
                                       "evaluation taken together)",
                             "bound": "tensor", "achieved": achieved, "peak": dgemm_tflops, "unit": "TFLOP/s",
                             "frac": achieved / dgemm_tflops, "traffic": read_traffic(n),
                             "note": "achieved = n^3 algorithmic flops / (potrf+trtri+lauum stage time of the last timed "
                                     "step, CUDA events on the launch stream); peak = cuBLAS DGEMM 8192^3 measured in "
                                     "this run (MEASURED_PEAKS.json has no fp64 entry; DMMA issue limit 37.2); traffic = "
                                     "DRAM bytes read + written by those kernels in one evaluation (ncu, "
                                     "profiles/r01_traffic_n32768.json), against 3 x 8 n^2 / 2 algorithmic matrix bytes: "
                                     "operand panels are re-streamed per tile task, at 1.2 TB/s = 19 % of the HBM peak"},
                "stages": stage_roof,
                "roofline_fill": {"kernel": "fill_hess_kernel", "bound": "hbm", "achieved": fill_gbs,
                                  "peak": hbm, "unit": "GB/s", "frac": fill_gbs / hbm,
                                  "bytes": 8.0 * n * n, "ms": t_fill * 1e3},
                "result": {"nll": float(res[0]), "grad": [float(res[1]), float(res[2])]},
                }
        if sweep:
            line["sweep"] = sweep
        if map_info:
            line["map"] = map_info
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def read_traffic(n):
    """DRAM bytes of the DMMA kernels per evaluation from the committed ncu capture of this matrix order, else None."""
    try:
        with open(os.path.join(ROOT, "profiles", f"r01_traffic_n{n}.json")) as f:
            return float(json.load(f)["dmma_dram_bytes_per_evaluation"])
    except Exception:
        return None


def read_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0          # fallback stated in B200_PROFILING.md


if __name__ == "__main__":
    main()
