#!/usr/bin/env python3
"""Golden fixture for the map path at BASELINE config 4 scale (04_standard_map: 4096 training pairs, 1000 map steps).

Runs the CPU oracle (oracle/csrc/sympgpr_oracle.c: sympgpr.f90:88-177 restated, hybrd1 n = 1 statement by statement,
pinned by tests/test_oracle.py) on orbits sampled from the benchmark ensemble of bench.py, with hybrd1 started at
p + guess (the oracle twin of the library's "newton_delta" start; the guess GP of this workload is trained on P - p as
python/04_standard_map/main.py:89-90 does), and records

  * the model itself (alpha, alphap: the GPU test must use the SAME model bit for bit -- cond(K) ~ 1e10, so two
    independently fitted models differ by far more than 1e-8 in their orbits),
  * the oracle orbits every 100 steps up to step 1000,
  * the oracle's own sensitivity per orbit: distance after 1000 steps between the orbits from p0 and p0 + 1e-12
    (the regular-orbit subset is "< 1e-9"), and between forward and reversed summation order of the
    training-set sums (rounding sensitivity of the same algorithm),
  * the largest residual |f(P)| an accepted root left (hybrd1's info is ignored by the reference, sympgpr.f90:107).

Usage: python tests/golden/make_golden_map_config4.py [candidates=2048] [threads=all]    (about 45 min on 8 cores)
"""
import os
import sys
import time

import numpy as np
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if len(sys.argv) > 2:
    os.environ["OMP_NUM_THREADS"] = sys.argv[2]
from oracle import c_oracle as C, oracle as O   # noqa: E402

NT, NM, EVERY, E_BENCH = 4096, 1001, 100, 100000
ECAND = int(sys.argv[1]) if len(sys.argv) > 1 else 2048


def wrapdist(a, b):
    d = np.abs(a - b)
    return np.minimum(d, np.abs(d - 2 * np.pi))


def main():
    t0 = time.time()
    d = O.standard_map_training(NT)
    l = 1.0 * 2 * np.pi / np.sqrt(NT)                    # bench.py: timing_hyp(..., factor=1.0)
    hyp = np.array([l, l, d["sig"], 1e-8])
    hypp = np.array([l, l, d["sigp"], 1e-8])
    xt, zt, xtp, ztp = d["xtrain"], d["ztrain"], d["xtrainp"], d["ztrainp"]
    K = C.build_k(xt[:NT], xt[NT:], xt[:NT], xt[NT:], hyp[:3])
    K[np.diag_indices_from(K)] += hyp[3]
    alpha = scipy.linalg.cho_solve(scipy.linalg.cho_factor(K, lower=True), zt)
    Kp = C.buildkreg(xtp[:NT], xtp[NT:], xtp[:NT], xtp[NT:], hypp[:3])
    Kp[np.diag_indices_from(Kp)] += hypp[3]
    alphap = scipy.linalg.cho_solve(scipy.linalg.cho_factor(Kp, lower=True), ztp)
    del K, Kp
    q0a = O.halton(E_BENCH, 5) * 2 * np.pi               # workloads.ensemble(1e5)
    p0a = 1.0 + O.halton(E_BENCH, 7) * 4.0
    idx = np.linspace(0, E_BENCH - 1, ECAND).astype(np.int64)
    q0, p0 = q0a[idx], p0a[idx]

    def run(p0_, **kw):
        t = time.time()
        out = C.applymap_alpha(2, NM, q0, p0_, hyp[:3], hypp[:3], xtp[:NT], xtp[NT:], alphap, xt[:NT], xt[NT:], alpha,
                               want_notconv=True, start_delta=True, out_every=EVERY, **kw)
        print(kw, "%.0f s" % (time.time() - t), "evaluations per orbit-step %.2f" % out[2], flush=True)
        return out

    A = run(p0)
    B = run(p0 + 1e-12)
    R = run(p0, reverse_sum=True)
    dist = lambda X, Y: np.maximum(wrapdist(X[0], Y[0]), wrapdist(X[1], Y[1]))
    sens_pert, sens_sum = dist(A, B), dist(A, R)
    out = os.path.join(ROOT, "tests", "golden", "map_config4_newton_delta.npz")
    np.savez_compressed(out, nt=NT, nm=NM, every=EVERY, e_bench=E_BENCH, idx=idx, hyp=hyp, hypp=hypp, alpha=alpha,
                        alphap=alphap, q0=q0, p0=p0, q=A[0], p=A[1], sens_pert=sens_pert, sens_sum=sens_sum,
                        maxres=A[-1], notconv=A[-2], evals_per_step=A[2], threads=C.num_threads(),
                        seconds=time.time() - t0)
    reg = (sens_pert[-1] < 1e-9) & (A[-1] < 1e-10)
    print("regular subset (p0 + 1e-12 -> < 1e-9 after 1000 steps, every root a root):", int(reg.sum()), "of", ECAND)
    print("  of those, summation-order distance < 1e-9:", int((sens_sum[-1][reg] < 1e-9).sum()),
          " max %.2e" % sens_sum[-1][reg].max())
    print("total %.0f s on %d threads" % (time.time() - t0, C.num_threads()))


if __name__ == "__main__":
    main()
