#!/usr/bin/env python3
"""Generate tests/golden/path_*.npz by running the REFERENCE's own pure-Python
GP layer (build container only; the fixtures are what travels to the GPU box).

The reference's compiled f2py modules cannot be built here (no Fortran
compiler), but its GP layer exists twice: as Fortran (sympgpr.f90) and as
Python loops (python/04_standard_map/func.py, python/02_pert_pendulum/func.py)
-- test_sympgpr.py asserts the two agree to 1e-12.  This script imports those
reference Python files unmodified, with

  * ``kernels``         -> float64 ``math``-lambdified versions of the SymPy
                           expressions built by the reference's init_func.py
                           (the same expression trees codegen prints to F95),
  * ``fortran.sympgpr`` -> an adapter that forwards build_k / buildkreg to the
                           reference's Python-loop build_K / buildKreg of
                           python/04_standard_map/func.py,

and records their outputs on (a) the literal inputs of test_sympgpr.py and
(b) a small standard-map training set.  The implicit root P is recorded as the
root of the reference's own ``Pnewton`` residual: located by MINPACK hybrd with
hybrd1's parameters from the GP guess, polished by brentq to ~1 ulp -- the value both
hybrd1(tol=1e-13) and an analytic Newton started at the same guess converge to.

    python tests/golden/make_golden_path.py
"""
import importlib.util
import os
import runpy
import sys
import tempfile
import types

import numpy as np
import scipy.optimize

REF = "/root/reference/python"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

EXPRS = [
    ("kern_num", "kern"), ("dkdx_num", "dkdxa"), ("dkdy_num", "dkdya"),
    ("dkdx0_num", "dkdxb"), ("dkdy0_num", "dkdyb"),
    ("d2kdxdx0_num", "dkdxadxb"), ("d2kdydy0_num", "dkdyadyb"), ("d2kdxdy0_num", "dkdxadyb"),
    ("d3kdxdx0dy0_num", "d3kdxdx0dy0"), ("d3kdydy0dy0_num", "d3kdydy0dy0"),
    ("d3kdxdy0dy0_num", "d3kdxdy0dy0"),
    ("dkdlx_num", "dkdlx"), ("dkdly_num", "dkdly"),
    ("d3kdxdx0dlx_num", "d3kdxdx0dlx"), ("d3kdydy0dlx_num", "d3kdydy0dlx"),
    ("d3kdxdy0dlx_num", "d3kdxdy0dlx"),
    ("d3kdxdx0dly_num", "d3kdxdx0dly"), ("d3kdydy0dly_num", "d3kdydy0dly"),
    ("d3kdxdy0dly_num", "d3kdxdy0dly"),
]


def reference_kernels_module(init_func_path, modname):
    import sympy
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            ns = runpy.run_path(init_func_path)
        finally:
            os.chdir(cwd)
    seq = [ns["xa"], ns["ya"], ns["xb"], ns["yb"], ns["lx"], ns["ly"]]
    mod = types.ModuleType(modname)
    for fname, var in EXPRS:
        setattr(mod, fname, sympy.lambdify(seq, ns[var], modules="math"))
    mod.__all__ = [f for f, _ in EXPRS]
    return mod


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def halton(n, base):
    out = np.zeros(n)
    for k in range(n):
        i, f, r = k + 1, 1.0, 0.0
        while i > 0:
            f /= base
            r += f * (i % base)
            i //= base
        out[k] = r
    return out


def root_of(fun, guess):
    """Root of the reference residual the way sympgpr.f90:107 finds it: MINPACK hybrd
    (SciPy's copy) with hybrd1's parameter block (minpack.f90:1570-1577), started at the
    GP guess -- this also fixes WHICH root is meant where the residual has several --
    then polished by brentq inside +-1e-9 so the stored value is the root to ~1 ulp."""
    sol, _, ier, _ = scipy.optimize.fsolve(lambda v: [fun(float(v[0]))], [guess], xtol=1e-13, maxfev=400,
                                           diag=[1.0], factor=100, epsfcn=0.0, full_output=True)
    r = float(sol[0])
    w = 1e-9 * max(1.0, abs(r))
    if fun(r - w) * fun(r + w) < 0:
        r = scipy.optimize.brentq(fun, r - w, r + w, xtol=1e-300, rtol=8.9e-16, maxiter=500)
    return r


def main(family, init_func, tag):
    kernels = reference_kernels_module(init_func, "kernels")
    sys.modules["kernels"] = kernels
    f04 = load(f"{REF}/04_standard_map/func.py", "ref04_func")

    # adapter so python/02_pert_pendulum/func.py imports: its build_K/buildKreg
    # forward to sympgpr.build_k/buildkreg(x, y, x0, y0, hyp, K)
    class _Sym:
        @staticmethod
        def build_k(x, y, x0, y0, hyp, K):
            f04.build_K(np.hstack((x, y)), np.hstack((x0, y0)), hyp, K)

        @staticmethod
        def buildkreg(x, y, x0, y0, hyp, K):
            f04.buildKreg(np.hstack((x, y)), np.hstack((x0, y0)), hyp, K)

    fortran = types.ModuleType("fortran")
    fs = types.ModuleType("fortran.sympgpr")
    fs.sympgpr = _Sym
    fortran.sympgpr = fs
    sys.modules["fortran"] = fortran
    sys.modules["fortran.sympgpr"] = fs
    f02 = load(f"{REF}/02_pert_pendulum/func.py", "ref02_func")

    out = {}
    # ---------------- (a) literal inputs of test_sympgpr.py:7-10,19,48-68
    x = np.array([1.0, 2.0, 3.0]); y = np.array([0.0, 3.0, 2.0])
    x0 = np.array([1.0, 2.0]); y0 = np.array([0.0, 3.0])
    hyp = np.array([0.5, 2.0, 0.4])
    K = np.empty((3, 2)); f04.buildKreg(np.hstack((x, y)), np.hstack((x0, y0)), hyp, K)
    out["lit_buildkreg"] = K.copy()
    K = np.zeros((1, 2)); f04.buildKreg(np.hstack((x[:1], y[:1])), np.hstack((x0, y0)), hyp, K)
    out["lit_buildkreg_1"] = K.copy()
    K = np.empty((6, 4)); f04.build_K(np.hstack((x, y)), np.hstack((x0, y0)), hyp, K)
    out["lit_build_k"] = K.copy()
    hypp = np.array([0.6, 1.9, 0.3])
    Kyinvp = np.array([[0.9, -0.3], [0.3, 0.9]])
    ztrainp = np.cos(x0 + y0)
    out["lit_guessp"] = np.array(f04.guessP([x[0]], [y[0]], hypp, np.hstack((x0, y0)), ztrainp, Kyinvp, 1)).ravel()
    Kyinv = np.reshape(np.arange(16), (4, 4), order="F").astype(float)
    ztrain = np.hstack((np.cos(x0 + y0), np.sin(x0 + y0)))
    out["lit_calcq"] = np.array([f04.calcQ(x[0], y[0], np.hstack((x0, y0)), hyp, Kyinv, ztrain)])

    def resid(P):
        return float(f04.Pnewton(np.array([P]), np.array([x[0]]), np.array([y[0]]), hyp,
                                 np.hstack((x0, y0)), Kyinv, ztrain))
    out["lit_calcp"] = np.array([root_of(resid, float(out["lit_guessp"][0]))])
    out["lit_resid_P"] = np.linspace(-1.0, 2.0, 7)
    out["lit_resid"] = np.array([resid(P) for P in out["lit_resid_P"]])

    # ---------------- (b) small standard-map set (main.py:27-59,89-92; k=0.9, Halton 2,3)
    N = 12
    kch = 0.9
    q = halton(N, 2) * 2 * np.pi
    p = halton(N, 3) * 2 * np.pi
    P = p + kch * np.sin(q)
    Q = q + P
    xtrain = np.hstack((q, P)); ztrain = np.concatenate((p - P, Q - q))
    xtrainp = np.hstack((q, p)); ztrainp = P - p
    sig = 2 * np.amax(np.abs(ztrain))**2
    sigp = 2 * np.amax(np.abs(ztrainp))**2
    out["N"] = np.array([N]); out["kchaos"] = np.array([kch])
    out["xtrain"] = xtrain; out["ztrain"] = ztrain; out["xtrainp"] = xtrainp; out["ztrainp"] = ztrainp
    hyps = np.array([[0.9, 1.1, sig, 1e-8], [0.5, 0.7, sig, 1e-6], [1.3, 0.8, 0.5 * sig, 1e-8]])
    hypps = np.array([[0.8, 1.2, sigp, 1e-8], [0.5, 0.9, sigp, 1e-6]])
    out["hyps"] = hyps; out["hypps"] = hypps
    K = np.empty((2 * N, 2 * N)); f04.build_K(xtrain, xtrain, hyps[0, :3], K)
    out["K"] = K.copy()
    Kp = np.empty((N, N)); f04.buildKreg(xtrainp, xtrainp, hypps[0, :3], Kp)
    out["Kreg"] = Kp.copy()
    out["nll_chol"] = np.array([f04.nll_chol(h, xtrain, ztrain, 2 * N) for h in hyps])
    out["nll_chol_reg"] = np.array([f04.nll_chol_reg(h, xtrainp, ztrainp, N) for h in hypps])
    dK = f02.build_dK(xtrain, xtrain, hyps[0, :3])
    out["dK_lx"] = dK[0]; out["dK_ly"] = dK[1]
    dKr = f02.build_dKreg(xtrainp, xtrainp, hypps[0, :3])
    out["dKreg_lx"] = dKr[0]; out["dKreg_ly"] = dKr[1]
    vals, grads = [], []
    for h in hyps:
        v, g = f02.nll_grad(h, xtrain, ztrain, 2 * N)
        vals.append(v); grads.append(g)
    out["nll_grad_val"] = np.array(vals); out["nll_grad_grad"] = np.array(grads)
    vals, grads = [], []
    for h in hypps:
        v, g = f02.nll_grad_reg(h, xtrainp, ztrainp, N)
        vals.append(v); grads.append(g)
    out["nll_grad_reg_val"] = np.array(vals); out["nll_grad_reg_grad"] = np.array(grads)

    # model finalisation exactly as main.py:96-118 (scipy.linalg.inv)
    import scipy.linalg
    hyp = hyps[0, :3]; hypp = hypps[0, :3]
    Kyinv = scipy.linalg.inv(K + hyps[0, 3] * np.eye(2 * N))
    Kyinvp = scipy.linalg.inv(Kp + hypps[0, 3] * np.eye(N))
    out["Kyinv"] = Kyinv; out["Kyinvp"] = Kyinvp

    # three map steps for a few orbits: P = root of the reference residual,
    # dq from the reference calcQ, wraps as applymap (04_standard_map/func.py:218-254)
    E = 6
    q0 = halton(E, 5) * 2 * np.pi
    p0 = halton(E, 7) * 2 * np.pi
    S = 4
    qm = np.zeros((S, E)); pm = np.zeros((S, E)); pg = np.zeros((S, E)); praw = np.zeros((S, E))
    qm[0] = q0; pm[0] = p0
    for i in range(S - 1):
        for k in range(E):
            g = float(np.ravel(f04.guessP([qm[i, k]], [pm[i, k]], hypp, xtrainp, ztrainp, Kyinvp, 1))[0])
            pg[i + 1, k] = g

            def resid(Pv, k=k, i=i):
                return float(f04.Pnewton(np.array([Pv]), np.array([qm[i, k]]), np.array([pm[i, k]]),
                                         hyp, xtrain, Kyinv, ztrain))
            Pn = root_of(resid, g)
            praw[i + 1, k] = Pn
            # applymap wraps p first and hands the WRAPPED value to calcQ (func.py:232,244)
            pm[i + 1, k] = np.mod(Pn, 2 * np.pi)
            dq = f04.calcQ(qm[i, k], pm[i + 1, k], xtrain, hyp, Kyinv, ztrain)
            qm[i + 1, k] = np.mod(dq + qm[i, k], 2 * np.pi)
    out["map_q"] = qm; out["map_p"] = pm; out["map_guess"] = pg; out["map_praw"] = praw
    np.savez(os.path.join(HERE, f"path_{tag}.npz"), **out)
    print("wrote", f"path_{tag}.npz")


if __name__ == "__main__":
    main("product", f"{REF}/04_standard_map/init_func.py", "product")
    main("sq", f"{REF}/03_henon_heiles/init_func.py", "sq")
