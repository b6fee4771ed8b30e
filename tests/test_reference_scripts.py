"""The reference's OWN Python layer and example scripts, unmodified, over the C-ABI boundary (the
pattern of python/05_tokamak/SympGPR/test_sympgpr.py:17-100, whose second implementation `func_old` and `test.pickle`
are missing from the reference: the committed goldens tests/golden/path_*.npz, produced from the reference's pure-Python
layer, take their place).

The reference tree is looked for in $SYMPGPR_REFERENCE, baseline/_ref/python (staged by tools/stage_reference.sh -- the
git-ignored location that travels to the GPU box) and /root/reference/python; without one the tests skip (reference sources
are never committed).  Packages the scripts import but the image lacks (ghalton, tkinter, matplotlib, cma, the VODE-based
`henon` extension) are stood in for from THIS side (tests/harness/standins.py); the product ships no fakes.

  not gpu : scripts whose GP layer is Python loops over the 19 scalar `kernels*` functions (host arithmetic of the library:
            01_pendulum/implicit, 04_standard_map) -- they need the shared library but no device
  gpu     : func.py layers that call sympgpr.build_k / buildkreg / guessp / calcq / calcp (02, 05, functions/) against the
            goldens, and the main.py scripts 02_pert_pendulum, 05_tokamak/SympGPR and the remaining host-layer scripts
"""
import contextlib
import importlib.util
import io
import os
import sys
import time

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = os.path.join(HERE, "golden")
sys.path.insert(0, os.path.join(HERE, "harness"))


def _ref_root():
    for cand in (os.environ.get("SYMPGPR_REFERENCE"), os.path.join(ROOT, "baseline", "_ref", "python"), "/root/reference/python"):
        if cand and os.path.isdir(os.path.join(cand, "02_pert_pendulum")):
            return cand
    return None


REF = _ref_root()
needs_ref = pytest.mark.skipif(REF is None, reason="reference tree not available (tools/stage_reference.sh stages it for the GPU box)")


def _run_script(rel, solver="hybrd", zero_empty=False):
    """Execute a reference script unchanged through sympgpr_b200.runner with the test-side stand-ins; returns its globals,
    what it printed and the seconds it took.  zero_empty: numpy.empty returns zeros while the script runs -- script 03
    reads uninitialised memory (its step-1 fit builds a 55 x 55 matrix with a loop that fills 54 rows and columns, SURVEY
    Appendix C.2; whether scipy.linalg.cholesky then sees NaN depends on what the allocator hands out)."""
    import standins
    standins.install()
    from sympgpr_b200 import runner
    real_empty = np.empty
    if zero_empty:
        np.empty = lambda shape, dtype=float, order="C", **kw: np.zeros(shape, dtype=dtype, order=order)
    # a fresh import of every module the scripts share by NAME (each example directory has its own func.py, calc_*.py)
    for name in ("func", "func_expl", "calc_poincare", "calc_fieldlines", "common", "kernels", "kernels_sq", "kernels_sum"):
        sys.modules.pop(name, None)
    buf = io.StringIO()
    t0 = time.time()
    try:
        with contextlib.redirect_stdout(buf):
            ns = runner.run(os.path.join(REF, rel), solver=solver)
    finally:
        np.empty = real_empty
        for name in ("func", "func_expl", "calc_poincare", "calc_fieldlines", "common"):
            sys.modules.pop(name, None)
    return ns, buf.getvalue(), time.time() - t0


def _load_ref_module(rel, name, family="product"):
    """Import one reference func.py unmodified under a private name, with the shims registered."""
    import sympgpr_b200
    import standins
    standins.install()
    mods = sympgpr_b200.install_shims()
    mods["sympgpr"].sympgpr.family = family
    mods["sympgpr"].sympgpr.solver = "hybrd"
    py_root = REF
    if py_root not in sys.path:
        sys.path.insert(1, py_root)                    # `from fortran.sympgpr import sympgpr` resolves through sys.modules
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------- host layer: no device needed
@needs_ref
@pytest.mark.parametrize("rel", ["01_pendulum/implicit/main.py", "04_standard_map/main.py"])
def test_host_layer_scripts_run_unchanged(rel):
    """Scripts whose func.py loops in Python over the scalar `kernels` functions: the library's host closed forms
    (sgp_kernel_scalar) stand behind `from kernels import *`; no device is touched."""
    ns, out, secs = _run_script(rel)
    assert "training error" in out, out[-400:]
    assert np.isfinite(float(ns["outtrain"])) and float(ns["outtrain"]) < 1e-6
    qmap, pmap = np.asarray(ns["qmap"]), np.asarray(ns["pmap"])
    assert qmap.shape == (int(ns["nm"]), int(ns["Ntest"])) and qmap.shape == pmap.shape
    assert np.isfinite(qmap).mean() > 0.9 and np.isfinite(pmap).mean() > 0.9
    print(f"\n{rel}: {secs:.1f} s, training error {float(ns['outtrain']):.1e}")


# ------------------------------------------------------------------------------- GPU: func.py layers against the goldens
@needs_ref
@pytest.mark.gpu
def test_reference_func_layers_over_the_gpu_boundary():
    """python/02_pert_pendulum/func.py, python/functions/func.py and python/05_tokamak/SympGPR/func.py imported UNMODIFIED;
    their build_K / buildKreg / build_dK / nll_chol(_reg) / nll_grad(_reg) / guessP / calcQ / calcP / applymap / applymap_tok
    run on libsympgpr_b200 behind `fortran.sympgpr` / `sympgpr` / `kernels` / `fieldlines` and must reproduce the values the
    reference's pure-Python layer gave (tests/golden/path_product.npz): 1e-12 for matrices and scalars
    (test_sympgpr.py:26-74), NLL 1e-9, the LU/dense-trace gradient 1e-6, map steps 1e-8 (test_sympgpr.py:92-93)."""
    from sympgpr_b200 import _lib
    assert _lib.device_count() >= 1
    g = np.load(os.path.join(G, "path_product.npz"))
    N = int(g["N"][0])
    xt, zt, xtp, ztp = g["xtrain"], g["ztrain"], g["xtrainp"], g["ztrainp"]
    hyp, hypp = g["hyps"][0], g["hypps"][0]
    f02 = _load_ref_module("02_pert_pendulum/func.py", "ref02_func_gpu")
    ffn = _load_ref_module("functions/func.py", "reffn_func_gpu")
    f05 = _load_ref_module("05_tokamak/SympGPR/func.py", "ref05_func_gpu")
    for f in (f02, ffn, f05):
        K = np.empty((2 * N, 2 * N), order="F")
        f.build_K(xt, xt, hyp[:3], K)
        assert np.allclose(K, g["K"], rtol=1e-12, atol=1e-12)
        Kp = np.zeros((N, N), order="F")
        f.buildKreg(xtp, xtp, hypp[:3], Kp)
        assert np.allclose(Kp, g["Kreg"], rtol=1e-12, atol=1e-12)
    # the literal inputs of test_sympgpr.py:7-10,19,48-68 through the reference's own wrappers
    x = np.array([1.0, 2.0, 3.0]); y = np.array([0.0, 3.0, 2.0]); x0 = np.array([1.0, 2.0]); y0 = np.array([0.0, 3.0])
    lh, lhp = np.array([0.5, 2.0, 0.4]), np.array([0.6, 1.9, 0.3])
    K = np.empty((6, 4), order="F")
    f05.build_K(np.hstack((x, y)), np.hstack((x0, y0)), lh, K)
    assert np.allclose(K, g["lit_build_k"], rtol=1e-12, atol=1e-12)
    Kyinvp = np.array([[0.9, -0.3], [0.3, 0.9]])
    Kyinv = np.reshape(np.arange(16), (4, 4), order="F")
    ztl, ztpl = np.hstack((np.cos(x0 + y0), np.sin(x0 + y0))), np.cos(x0 + y0)
    for f in (f02, ffn, f05):
        assert np.allclose(f.guessP(x[0], y[0], lhp, np.hstack((x0, y0)), ztpl, Kyinvp), g["lit_guessp"], rtol=1e-12, atol=1e-12)
        assert np.allclose(f.calcQ(x[0], y[0], np.hstack((x0, y0)), lh, Kyinv, ztl), g["lit_calcq"], rtol=1e-12, atol=1e-12)
        P = f.calcP(x[0], y[0], lh, lhp, np.hstack((x0, y0)), ztpl, Kyinvp, np.hstack((x0, y0)), ztl, Kyinv)
        assert np.allclose(P, g["lit_calcp"], rtol=1e-12, atol=1e-12)
    # NLL / gradient: the reference's own Python (LAPACK calls, build_dK loops over the scalar kernels) on our fills
    for k, h in enumerate(g["hyps"]):
        assert np.isclose(f02.nll_chol(h, xt, zt, 2 * N), g["nll_chol"][k], rtol=1e-9)
        assert np.isclose(f05.nll_chol(h, xt, zt, 2 * N), g["nll_chol"][k], rtol=1e-9)
        v, gr = f02.nll_grad(h, xt, zt, 2 * N)
        assert np.isclose(v, g["nll_grad_val"][k], rtol=1e-9)
        assert np.allclose(gr, g["nll_grad_grad"][k], rtol=1e-6, atol=1e-6)
    for k, h in enumerate(g["hypps"]):
        assert np.isclose(f02.nll_chol_reg(h, xtp, ztp, N), g["nll_chol_reg"][k], rtol=1e-9)
        v, gr = f02.nll_grad_reg(h, xtp, ztp, N)
        assert np.isclose(v, g["nll_grad_reg_val"][k], rtol=1e-9) and np.allclose(gr, g["nll_grad_reg_grad"][k], rtol=1e-6, atol=1e-6)
    dK = f02.build_dK(xt, xt, hyp[:3])
    assert np.allclose(dK[0], g["dK_lx"], rtol=1e-12, atol=1e-12) and np.allclose(dK[1], g["dK_ly"], rtol=1e-12, atol=1e-12)
    # map loops: the reference's applymap (q wrapped, p not) and applymap_tok against a re-derivation from the golden roots
    S, E = g["map_q"].shape
    q0, p0 = g["map_q"][0], g["map_p"][0]
    for f, name in ((f02, "applymap"), (ffn, "applymap"), (f05, "applymap_tok")):
        qm, pm = getattr(f, name)(2, E, hyp[:3], hypp[:3], q0, p0, xtp, ztp, g["Kyinvp"], xt, zt, g["Kyinv"])
        lost = np.isnan(pm[1])
        assert np.allclose(pm[1][~lost], g["map_praw"][1][~lost], rtol=1e-8, atol=1e-8), name       # the unwrapped root
        # dq at the UNWRAPPED momentum here (pendulum loop), so recompute the expected angle with the reference's own calcQ
        for k in np.nonzero(~lost)[0]:
            dq = f.calcQ(q0[k], pm[1, k], xt, hyp[:3], g["Kyinv"], zt)
            assert abs(qm[1, k] - np.mod(dq + q0[k], 2 * np.pi)) < 1e-8
        if name == "applymap_tok":
            r_ok = np.array([f05.fieldlines.compute_r(np.array([g["map_praw"][1][k] * 1e-2, q0[k], 0.0]), 0.3) <= 0.5 and
                             g["map_praw"][1][k] >= 0 for k in range(E)])
            assert np.array_equal(~lost, r_ok)


def _func_old_from_oracle():
    """The missing second implementation of test_sympgpr.py (`func_old`, the pure-Python-loop GP layer): the same six
    functions backed by the CPU oracle (NumPy fills, SciPy's MINPACK hybrd with hybrd1's parameters)."""
    import types
    from oracle import oracle as O
    m = types.ModuleType("func_old")

    def build_K(xin, x0in, hyp, K):
        N, N0 = K.shape[0] // 2, K.shape[1] // 2
        K[:, :] = O.build_k_vec(xin[:N], xin[N:], x0in[:N0], x0in[N0:], hyp)

    def buildKreg(xin, x0in, hyp, K):
        N, N0 = K.shape
        K[:, :] = O.buildkreg_vec(xin[:N], xin[N:], x0in[:N0], x0in[N0:], hyp)

    def guessP(x, y, hypp, xtrainp, ztrainp, Kyinvp):
        n = len(xtrainp) // 2
        return O.guessp(x, y, hypp, xtrainp[:n], xtrainp[n:], ztrainp, Kyinvp)

    def calcQ(x, y, xtrain, l, Kyinv, ztrain):
        n = len(xtrain) // 2
        return O.calcq(x, y, xtrain[:n], xtrain[n:], l, Kyinv, ztrain)

    def calcP(x, y, l, hypp, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv):
        n, np_ = len(xtrain) // 2, len(xtrainp) // 2
        return O.calcp(x, y, l, hypp, xtrainp[:np_], xtrainp[np_:], ztrainp, Kyinvp, xtrain[:n], xtrain[n:], ztrain, Kyinv)

    def applymap_tok(nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv):
        n, np_ = len(xtrain) // 2, len(xtrainp) // 2
        return O.applymap(O.MAP_TOKAMAK, nm, Q0map[:Ntest], P0map[:Ntest], l, hypp, xtrainp[:np_], xtrainp[np_:], ztrainp, Kyinvp,
                          xtrain[:n], xtrain[n:], ztrain, Kyinv)
    for f in (build_K, buildKreg, guessP, calcQ, calcP, applymap_tok):
        setattr(m, f.__name__, f)
    return m


@needs_ref
@pytest.mark.gpu
def test_reference_test_sympgpr_runs_unchanged(tmp_path):
    """python/05_tokamak/SympGPR/test_sympgpr.py -- the reference's only hot-path test -- executed UNCHANGED: `sympgpr` is the
    GPU shim, `func` the reference's own func.py (over the shim), and the two artefacts missing from the reference are
    supplied from the test side: `func_old` (the second implementation it compares against) = the CPU oracle, and
    `test.pickle` (a trained model + initial conditions for two applymap_tok steps) = a small field-line-like model.
    Its assertions are the reference's: 1e-12 for buildkreg / build_k / guessp / calcq / calcp, 1e-8 for applymap_tok,
    each checked once against func.py and once against func_old (test_sympgpr.py:26-100)."""
    import pickle
    import runpy
    import standins
    import sympgpr_b200
    from oracle import oracle as O
    from sympgpr_b200 import workloads as W
    standins.install()
    mods = sympgpr_b200.install_shims()
    mods["sympgpr"].sympgpr.family, mods["sympgpr"].sympgpr.solver = "product", "hybrd"
    N = 30
    d = W.tokamak_training(N)
    hyp = W.aniso_hyp(N, d["sig"], 2 * np.pi, 9.4, 1.5, 1e-8)
    hypp = W.aniso_hyp(N, d["sigp"], 2 * np.pi, 9.4, 1.5, 1e-8)
    xt, zt, xtp, ztp = d["xtrain"], d["ztrain"], d["xtrainp"], d["ztrainp"]
    Kyinv = np.linalg.inv(O.build_k_vec(xt[:N], xt[N:], xt[:N], xt[N:], hyp[:3]) + hyp[3] * np.eye(2 * N))
    Kyinvp = np.linalg.inv(O.buildkreg_vec(xtp[:N], xtp[N:], xtp[:N], xtp[N:], hypp[:3]) + hypp[3] * np.eye(N))
    Ntest = 9
    Q0 = O.halton(Ntest, 5) * 2 * np.pi
    P0 = np.array([2.0, 3.5, 5.0, 6.5, 7.5, 1.0, 4.2, 7.1, 5.7])        # no lost orbit: the reference compares with allclose (NaN != NaN)
    with open(tmp_path / "test.pickle", "wb") as f:
        pickle.dump((2, Ntest, hyp[:3], hypp[:3], Q0, P0, xtp, ztp, Kyinvp, xt, zt, Kyinv), f)
    d05 = os.path.join(REF, "05_tokamak", "SympGPR")
    for name in ("func", "func_old", "calc_fieldlines", "common"):
        sys.modules.pop(name, None)
    sys.modules["func_old"] = _func_old_from_oracle()
    cwd = os.getcwd()
    sys.path.insert(0, d05)
    os.chdir(tmp_path)                                  # open('test.pickle') is relative to the working directory
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(d05, "test_sympgpr.py"), run_name="ref_test_sympgpr")
    finally:
        os.chdir(cwd)
        sys.path.remove(d05)
        for name in ("func", "func_old"):
            sys.modules.pop(name, None)
    out = buf.getvalue()
    assert out.count("applymap_tok matches") == 2 and out.count("calcP matches") == 2 and out.count("build_K matches") == 2, out


# ------------------------------------------------------------------------------- GPU: the example scripts at their own sizes
@needs_ref
@pytest.mark.gpu
def test_script_02_pert_pendulum_runs_unchanged_on_gpu():
    """python/02_pert_pendulum/main.py: L-BFGS-B on nll_grad / nll_grad_reg (sympgpr.build_k / buildkreg fills on the GPU, the
    script's own inv / build_dK), scipy.linalg.inv, then applymap = 2 x 30 x 99 sympgpr.calcp / calcq calls."""
    from sympgpr_b200 import _lib
    h0, m0 = _alpha_stats(_lib)
    ns, out, secs = _run_script("02_pert_pendulum/main.py")
    assert "training error" in out
    assert np.isfinite(float(ns["outtrain"])) and float(ns["outtrain"]) < 1e-6, out[-400:]
    qmap, pmap = np.asarray(ns["qmap"]), np.asarray(ns["pmap"])
    assert qmap.shape == (int(ns["nm"]), int(ns["Ntest"]))
    assert np.isfinite(qmap).all() and np.isfinite(pmap).all()
    # the learned map follows the Runge-Kutta reference orbits of the script for the first steps (it is a model of them)
    yint = np.asarray(ns["yinttest"])
    assert np.abs(pmap[1] - yint[1, :, 1]).max() < 5e-2
    h1, m1 = _alpha_stats(_lib)
    calls = 2 * (int(ns["nm"]) - 1) * int(ns["Ntest"])
    assert h1 - h0 >= calls and m1 - m0 <= 4, (h0, m0, h1, m1)          # Kyinv uploaded once, not per call
    print(f"\n02_pert_pendulum/main.py: {secs:.1f} s on the GPU boundary; training error {float(ns['outtrain']):.1e}; "
          f"alpha cache {h1 - h0} hits / {m1 - m0} misses")


def _alpha_stats(_lib):
    import ctypes
    h, m = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    _lib.check(_lib.lib().sgp_alpha_cache_stats(_lib.context().handle, ctypes.byref(h), ctypes.byref(m)), "stats")
    return h.value, m.value


@needs_ref
@pytest.mark.gpu
def test_script_05_tokamak_runs_unchanged_on_gpu():
    """python/05_tokamak/SympGPR/main.py (with its calc_fieldlines.py): field lines traced by the fieldlines shim's host
    integrator, nll_chol(_reg) fits through sympgpr.build_k / buildkreg, applymap_tok = 30 orbits x 999 steps of
    sympgpr.calcp / calcq + fieldlines.compute_r."""
    ns, out, secs = _run_script("05_tokamak/SympGPR/main.py")
    assert "training error" in out
    assert np.isfinite(float(ns["outtrain"])) and float(ns["outtrain"]) < 1e-6, out[-400:]
    qmap, pmap = np.asarray(ns["qmap"]), np.asarray(ns["pmap"])
    assert qmap.shape == (int(ns["nm"]), int(ns["Ntest"]))
    assert np.isfinite(pmap[0]).all() and np.isfinite(pmap[1]).mean() > 0.5
    print(f"\n05_tokamak/SympGPR/main.py: {secs:.1f} s; training error {float(ns['outtrain']):.1e}; "
          f"orbits alive after {int(ns['nm'])} steps: {int(np.isfinite(pmap[-1]).sum())} of {int(ns['Ntest'])}")


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("rel", ["01_pendulum/explicit/main.py", "01_pendulum/implicit_period_unknown/main.py",
                                 "03_henon_heiles/main.py"])
def test_remaining_scripts_run_unchanged(rel):
    """The other example scripts (host-layer GP maths over the scalar `kernels*` functions; 03 with the test-side stand-in
    for the VODE tracer `henon`): they finish and print a finite training error."""
    ns, out, secs = _run_script(rel, zero_empty=rel.startswith("03_"))
    assert "training error" in out, out[-400:]
    assert np.isfinite(float(ns["outtrain"])), out[-400:]
    print(f"\n{rel}: {secs:.1f} s, training error {float(ns['outtrain']):.1e}")
