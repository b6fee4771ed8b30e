#!/usr/bin/env python3
"""Generate tests/golden/path_expl.npz by running the REFERENCE's own explicit-map Python code
(build container only): nll_expl, calcP_expl and applymap_expl of python/04_standard_map/func.py
and applymap / calcP / calcQ of python/01_pendulum/explicit/func_expl.py, imported unmodified, with
``kernels`` = float64 lambdify of the sum-kernel expressions that
python/01_pendulum/explicit/init_func.py builds (periodic(q) + SE(P); the same kernel the
committed kernels_expl_per_q_sq_p.f90 / kernels_sum.f90 were generated from).

    python tests/golden/make_golden_expl.py
"""
import os
import sys

import numpy as np
import scipy.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_path import REF, halton, load, reference_kernels_module   # noqa: E402


def main():
    kernels = reference_kernels_module(f"{REF}/01_pendulum/explicit/init_func.py", "kernels")
    sys.modules["kernels"] = kernels
    sys.modules["kernels_sum"] = kernels
    f04 = load(f"{REF}/04_standard_map/func.py", "ref04_func_expl")
    f01 = load(f"{REF}/01_pendulum/explicit/func_expl.py", "ref01_func_expl")

    N, kch = 14, 0.9
    q = halton(N, 2) * 2 * np.pi
    p = halton(N, 3) * 2 * np.pi
    P = p + kch * np.sin(q)
    Q = q + P
    # explicit variant of python/04_standard_map/main.py:150-190: xtrain = [q; p] (the OLD momentum), ztrain = [p - P; Q - q]
    xtrain = np.hstack((q, p)); ztrain = np.concatenate((p - P, Q - q))
    sig = 2 * np.amax(np.abs(ztrain))**2
    out = dict(N=np.array([N]), xtrain=xtrain, ztrain=ztrain)
    hyp = np.array([0.9, 1.4, sig]); sig2n = 1e-3
    out["hyp"] = hyp; out["sig2n"] = np.array([sig2n])
    # nll_expl(hyp=[l, sig, sig2n], x, y, N, ind)   func.py:126-141 (np.hstack((hyp[0], 0, hyp[1])): l of the other block = 0)
    with np.errstate(all="ignore"):
        out["nll_expl_0"] = np.array([f04.nll_expl(np.array([l, sig, sig2n]), xtrain, ztrain[:N], 2 * N, 0) for l in (0.9, 0.5)])
        out["nll_expl_1"] = np.array([f04.nll_expl(np.array([l, sig, sig2n]), xtrain, ztrain[N:], 2 * N, 1) for l in (0.6, 0.35)])
    K = np.empty((2 * N, 2 * N)); f04.build_K(xtrain, xtrain, hyp, K)
    out["K"] = K.copy()
    Kyinv = scipy.linalg.inv(K + sig2n * np.eye(2 * N))
    out["Kyinv"] = Kyinv
    E, nm = 5, 6
    q0 = halton(E, 5) * 2 * np.pi
    p0 = halton(E, 7) * 2 * np.pi
    out["q0"] = q0; out["p0"] = p0
    out["calcp_expl"] = np.array([f04.calcP_expl(q0[k], p0[k], hyp, xtrain, ztrain, Kyinv) for k in range(E)])
    qm, pm, pd = f04.applymap_expl(nm, E, hyp, q0, p0, xtrain, ztrain, Kyinv)
    out["std_q"] = qm; out["std_p"] = pm; out["std_pdiff"] = pd
    qm1, pm1 = f01.applymap(hyp, q0, p0, xtrain, ztrain, Kyinv, E, nm)
    out["pen_q"] = qm1; out["pen_p"] = pm1
    np.savez(os.path.join(HERE, "path_expl.npz"), **out)
    print("wrote path_expl.npz")


if __name__ == "__main__":
    main()
