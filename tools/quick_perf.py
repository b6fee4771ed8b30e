"""Scratch: quick timing of the stages on a GPU box (not part of the product)."""
import ctypes, sys, time
import numpy as np
sys.path.insert(0, ".")  # run from the repo root
from oracle import oracle as O
from sympgpr_b200 import _lib, api

L = _lib.lib(); ctx = _lib.context()
for N in [1024, 4096, 8192, 16384]:
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    for rep in range(2):
        t = time.time(); v = api.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N); t1 = time.time() - t
        t = time.time(); v2, g = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N); t2 = time.time() - t
    n = 2 * N
    print(f"N={N} n={n}: nll {t1*1e3:.1f} ms ({n**3/3/t1/1e12:.2f} TF)  nll+grad {t2*1e3:.1f} ms ({n**3/t2/1e12:.2f} TF) val={v:.10g} grad={g}", flush=True)
    if N <= 4096:
        vr, gr = O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
        print("   oracle", vr, gr, "rel", abs(v2 - vr) / abs(vr), np.abs(g - gr).max() / np.abs(gr).max(), flush=True)
# map
for Nt in [1024, 4096]:
    d = O.standard_map_training(Nt)
    hyp = O.timing_hyp(Nt, d["sig"], 1e-8); hypp = O.timing_hyp(Nt, d["sigp"], 1e-8)
    hyp[:2] *= 2; hypp[:2] *= 2
    f = api.fit(hyp, d["xtrain"], d["ztrain"], 2 * Nt); fp = api.fit(hypp, d["xtrainp"], d["ztrainp"], Nt, reg=True)
    E = 100000
    q0 = O.halton(E, 5) * 2 * np.pi; p0 = 1.0 + O.halton(E, 7) * 4.0
    for solver in ("newton", "hybrd"):
        nm = 11
        t = time.time()
        out = api.applymap(nm, E, hyp[:3], hypp[:3], q0, p0, d["xtrainp"], None, None, d["xtrain"], None, None, solver=solver,
                           alphap=fp["alpha"], alpha=f["alpha"], out_every=0, return_stats=True)
        dt = time.time() - t
        st = out[-1]
        print(f"map Nt={Nt} E={E} steps={nm-1} {solver}: {dt:.3f} s -> {E*(nm-1)/dt:.3e} orbit-steps/s, evals/orbit-step {st['evaluations']/(E*(nm-1)):.2f} unconverged {st['unconverged']}", flush=True)
