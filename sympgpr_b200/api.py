"""Host-side mirror of the reference's GP layer (python/*/func.py) on top of the C ABI.

Function names, argument order and error behaviour follow the reference so its scripts and
tests read the same; everything numerical happens in libsympgpr_b200.so on the GPU.

  build_K, buildKreg                 python/02_pert_pendulum/func.py:32-50
  nll_chol, nll_chol_reg             python/05_tokamak/SympGPR/func.py:134-150
  nll_grad, nll_grad_reg             python/02_pert_pendulum/func.py:132-162
  guessP, calcQ, calcP               python/02_pert_pendulum/func.py:207-223
  applymap, applymap_henon           python/functions/func.py:216-260
  applymap_standard                  python/04_standard_map/func.py:218-254
  applymap_tok                       python/05_tokamak/SympGPR/func.py:182-211

Additive (not in the reference): fit(); the `family` / `per` / `solver` keyword arguments ("newton_delta": Newton
started at p + guess for guess GPs trained on P - p); applymap_quality() / quality() (the reference's `quality`
metrics accumulated inside the map kernel, no histories); StandardMapIterate() on the device; the 2-DOF 4x4-block
kernel (build_k4, nll_chol4, nll_grad4, fit(reg=4), applymap4); explicit and split maps (applymap_expl,
applymap_tok_split).  Ensembles that stay on the device go through the C ABI's sgp_model_* entry points with device
pointers (bench.py, ensemble.py hold them as torch tensors).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (ENERGIES, FAMILIES, MAP_KINDS, RES_A, RES_B, RES_DLX, RES_INFO, RES_LEN, RES_LOGD, RES_NLL, RES_QUAD, SOLVERS,
                   as_f64, as_f64_fortran, check, dptr)

_NULL = ctypes.POINTER(ctypes.c_double)()


def _fam(family):
    if isinstance(family, str):
        return FAMILIES[family]
    return int(family)


def _solver(s):
    return SOLVERS[s] if isinstance(s, str) else int(s)


def _kind(k):
    return MAP_KINDS[k] if isinstance(k, str) else int(k)


def _hyp3(h, name="hyp"):
    h = as_f64(h).ravel()
    if h.size != 3:
        # f2py: "0-th dimension must be fixed to 3 but got n"
        raise ValueError(f"{name}: 0-th dimension must be fixed to 3 but got {h.size}")
    return h


def _inout_matrix(K, name="k"):
    """f2py intent(inout): float64 and Fortran-contiguous, else ValueError.
    A C-contiguous array is accepted when it is square-symmetric use (x == x0 fills), as
    python/functions/func.py:133,149 allocates -- handled by the caller through a transpose."""
    if not isinstance(K, np.ndarray) or K.dtype != np.float64 or K.ndim != 2:
        raise ValueError(f"failed to initialize intent(inout) array -- expected a 2-d float64 array for `{name}`")
    return K


# --------------------------------------------------------------------------- fills
def _fill(fn, x, y, x0, y0, hyp, K, per, family, half):
    K = _inout_matrix(K)
    x, y, x0, y0 = as_f64(x).ravel(), as_f64(y).ravel(), as_f64(x0).ravel(), as_f64(y0).ravel()
    hyp = _hyp3(hyp)
    rows, cols = K.shape
    N = rows // 2 if half else rows
    N0 = cols // 2 if half else cols
    if x.size < N or y.size < N or x0.size < N0 or y0.size < N0:
        raise ValueError("build_k: point arrays shorter than the matrix dimensions require")
    ctx = _lib.context()
    if K.flags.f_contiguous:
        tgt, ld, swap = K, max(rows, 1), False
    elif K.flags.c_contiguous:
        # a C-ordered (rows, cols) array is a Fortran-ordered (cols, rows) one: fill the transpose
        tgt, ld, swap = K, max(cols, 1), True
    else:
        raise ValueError("failed to initialize intent(inout) array -- input not contiguous")
    if swap:
        st = fn(ctx.handle, _fam(family), per, dptr(x0), dptr(y0), N0, dptr(x), dptr(y), N, dptr(hyp), dptr(tgt), ld)
    else:
        st = fn(ctx.handle, _fam(family), per, dptr(x), dptr(y), N, dptr(x0), dptr(y0), N0, dptr(hyp), dptr(tgt), ld)
    check(st, "build_k")
    if half:
        # sympgpr.f90:21-22,37: with odd dimensions the last row/column is not written by the
        # loop but still scaled by hyp(3)
        if rows % 2:
            K[rows - 1, :] *= hyp[2]
        if cols % 2:
            K[: rows - (rows % 2), cols - 1] *= hyp[2]


def build_k(x, y, x0, y0, hyp, K, family="product", per=0.5):
    """sympgpr.build_k(x, y, x0, y0, hyp, k) -- sympgpr.f90:12-38."""
    _fill(_lib.lib().sgp_build_k, x, y, x0, y0, hyp, K, per, family, True)


def buildkreg(x, y, x0, y0, hyp, K, family="product", per=0.5):
    """sympgpr.buildkreg(x, y, x0, y0, hyp, k) -- sympgpr.f90:40-60."""
    _fill(_lib.lib().sgp_buildkreg, x, y, x0, y0, hyp, K, per, family, False)


def build_K(xin, x0in, hyp, K, family="product", per=0.5):
    """python/02_pert_pendulum/func.py:32-40."""
    N = K.shape[0] // 2
    N0 = K.shape[1] // 2
    xin, x0in = np.asarray(xin), np.asarray(x0in)
    build_k(xin[0:N], xin[N:2 * N], x0in[0:N0], x0in[N0:2 * N0], hyp, K, family, per)


def buildKreg(xin, x0in, hyp, K, family="product", per=0.5):
    """python/02_pert_pendulum/func.py:42-50."""
    N, N0 = K.shape
    xin, x0in = np.asarray(xin), np.asarray(x0in)
    buildkreg(xin[0:N], xin[N:2 * N], x0in[0:N0], x0in[N0:2 * N0], hyp, K, family, per)


# --------------------------------------------------------------------------- NLL
def _nll(hyp, x, y, n, reg, ngrad, family, per):
    hyp = as_f64(hyp).ravel()
    if hyp.size != 4:
        raise ValueError("hyp must be [lx, ly, sig, sig2n]")
    n = int(n)
    nx = n if int(reg) in (0, 4) else 2 * n          # coordinates: [x; y] (2N values) or [q1; q2; P1; P2] (reg = 4)
    x = as_f64(x).ravel()
    y = as_f64(y).ravel()
    if x.size < nx or y.size < n:
        raise ValueError("nll: x must hold the coordinates of all points and y n values")
    res = np.zeros(RES_LEN)
    st = _lib.lib().sgp_nll(_lib.context().handle, _fam(family), per, int(reg), dptr(hyp), dptr(x), dptr(y), n, ngrad,
                            dptr(res))
    check(st, "cholesky")
    return res


def nll_chol(hyp, x, y, N, family="product", per=0.5):
    """Negative log marginal likelihood of the derivative-kernel GP; N is the matrix order 2N."""
    return float(_nll(hyp, x, y, N, False, 0, family, per)[RES_NLL])


def nll_chol_reg(hyp, x, y, N, family="product", per=0.5):
    return float(_nll(hyp, x, y, N, True, 0, family, per)[RES_NLL])


def nll_grad(hyp, x, y, N, family="product", per=0.5, with_sig=False, reference_third_component=False):
    """(value, gradient w.r.t. [lx, ly]) as python/02_pert_pendulum/func.py:148-162.

    with_sig=True appends d/dsig (python/05_tokamak/SympGPR/func.py:152-168);
    reference_third_component=True reproduces that function's third entry as written there
    (quadratic term taken from dK[1], trace from dK[2]).
    """
    r = _nll(hyp, x, y, N, False, 3 if with_sig else 2, family, per)
    g = [r[RES_DLX], r[RES_DLX + 1]]
    if with_sig:
        if reference_third_component:
            g.append(-0.5 * r[RES_A + 1] + 0.5 * r[RES_B + 2])
        else:
            g.append(r[RES_DLX + 2])
    return float(r[RES_NLL]), np.array(g)


def nll_grad_reg(hyp, x, y, N, family="product", per=0.5):
    """python/02_pert_pendulum/func.py:132-146."""
    r = _nll(hyp, x, y, N, True, 2, family, per)
    return float(r[RES_NLL]), np.array([r[RES_DLX], r[RES_DLX + 1]])


# --------------------------------------------------------------------------- 2-DOF 4x4-block kernel (not in the reference)
def build_k4(x, x0, hyp, K):
    """2-DOF generalisation of build_K (BASELINE config 3; csrc/dof2.cu): x = [q1; q2; P1; P2]
    (4N values), x0 likewise, hyp = [lq, lP, sig]; fills the Fortran-ordered (4N, 4N0) array K in place."""
    K = _inout_matrix(K)
    if not K.flags.f_contiguous:
        raise ValueError("failed to initialize intent(inout) array -- input not Fortran contiguous")
    x, x0, hyp = as_f64(x).ravel(), as_f64(x0).ravel(), _hyp3(hyp)
    N, N0 = K.shape[0] // 4, K.shape[1] // 4
    if x.size < 4 * N or x0.size < 4 * N0:
        raise ValueError("build_k4: point arrays shorter than the matrix dimensions require")
    check(_lib.lib().sgp_build_k4(_lib.context().handle, dptr(x), N, dptr(x0), N0, dptr(hyp), dptr(K), max(K.shape[0], 1)), "build_k4")


def nll_chol4(hyp, x, y, n):
    """NLL of the 2-DOF derivative-kernel GP: hyp = [lq, lP, sig, sig2n], x = [q1; q2; P1; P2], y the n = 4N observations
    [p1 - P1; p2 - P2; Q1 - q1; Q2 - q2]."""
    return float(_nll(hyp, x, y, n, 4, 0, "sq", 0.5)[RES_NLL])


def nll_grad4(hyp, x, y, n, with_sig=False):
    """(value, gradient w.r.t. [lq, lP(, sig)]) of the 2-DOF derivative-kernel NLL."""
    r = _nll(hyp, x, y, n, 4, 3 if with_sig else 2, "sq", 0.5)
    return float(r[RES_NLL]), np.array(r[RES_DLX:RES_DLX + (3 if with_sig else 2)])


def nll_expl(hyp, x, y, N, ind, family="sum", per=0.5):
    """nll_expl(hyp=[l, sig, sig2n], x, y, N, ind) -- python/04_standard_map/func.py:126-141.

    N is the order 2*len(y) of the full derivative-kernel matrix, x = [q; p] its N coordinates, y the N/2
    observations of the block: ind = 0 fits lx on the (q,q) block, ind = 1 fits ly on the (P,P) block."""
    hyp = as_f64(hyp).ravel()
    if hyp.size != 3:
        raise ValueError("nll_expl: hyp must be [l, sig, sig2n]")
    h4 = np.array([hyp[0], hyp[0], hyp[1], hyp[2]])
    return float(_nll(h4, x, y, int(N) // 2, 2 if int(ind) == 0 else 3, 0, family, per)[RES_NLL])


def fit(hyp, x, z, n, reg=False, want_inverse=False, want_factor=False, family="product", per=0.5):
    """Model finalisation: alpha = (K + |sig2n| I)^-1 z [, Kyinv, L]; returns a dict."""
    hyp = as_f64(hyp).ravel()
    n = int(n)
    N = n if reg else n // 2
    x = as_f64(x).ravel()
    z = as_f64(z).ravel()
    if int(reg) == 4:                  # 2-DOF 4 x 4-block kernel: x = [q1; q2; P1; P2], n = 4 N coordinates in all
        N = n // 2
    if hyp.size != 4 or x.size < 2 * N or z.size < n:
        raise ValueError("fit: bad argument sizes")
    alpha = np.zeros(n)
    kyinv = np.zeros((n, n), order="F") if want_inverse else None
    L = np.zeros((n, n), order="F") if want_factor else None
    res = np.zeros(RES_LEN)
    st = _lib.lib().sgp_fit(_lib.context().handle, _fam(family), per, int(reg), dptr(hyp), dptr(x), dptr(z), n, dptr(alpha),
                            dptr(kyinv) if want_inverse else _NULL, dptr(L) if want_factor else _NULL, dptr(res))
    check(st, "cholesky")
    return dict(alpha=alpha, Kyinv=kyinv, L=L, nll=float(res[RES_NLL]), quad=float(res[RES_QUAD]),
                logdet_half=float(res[RES_LOGD]))


# --------------------------------------------------------------------------- prediction
def _scalar(v, name):
    a = as_f64(v).ravel()
    if a.size != 1:
        raise ValueError(f"{name}: expected a scalar / length-1 array")
    return float(a[0])


def guessp(x, y, hypp, xtrainp, ytrainp, ztrainp, kyinvp, family="product", per=0.5):
    """sympgpr.guessp -- sympgpr.f90:62-73."""
    hypp = _hyp3(hypp, "hypp")
    xtp, ytp, ztp = as_f64(xtrainp).ravel(), as_f64(ytrainp).ravel(), as_f64(ztrainp).ravel()
    kyi = as_f64_fortran(kyinvp)
    np_ = xtp.size
    if ytp.size != np_ or ztp.size != np_ or kyi.shape != (np_, np_):
        raise ValueError("guessp: inconsistent training-set shapes")
    out = ctypes.c_double(0.0)
    st = _lib.lib().sgp_guessp(_lib.context().handle, _fam(family), per, _scalar(x, "x"), _scalar(y, "y"), dptr(hypp),
                               dptr(xtp), dptr(ytp), dptr(ztp), dptr(kyi), np_, ctypes.byref(out))
    check(st, "guessp")
    return out.value


def calcq(x, y, xtrain, ytrain, hyp, kyinv, ztrain, family="product", per=0.5):
    """sympgpr.calcq -- sympgpr.f90:75-86."""
    hyp = _hyp3(hyp)
    xt, yt, zt = as_f64(xtrain).ravel(), as_f64(ytrain).ravel(), as_f64(ztrain).ravel()
    kyi = as_f64_fortran(kyinv)
    nt = xt.size
    if yt.size != nt or zt.size != 2 * nt or kyi.shape != (2 * nt, 2 * nt):
        raise ValueError("calcq: inconsistent training-set shapes")
    out = ctypes.c_double(0.0)
    st = _lib.lib().sgp_calcq(_lib.context().handle, _fam(family), per, _scalar(x, "x"), _scalar(y, "y"), dptr(xt), dptr(yt),
                              dptr(hyp), dptr(kyi), dptr(zt), nt, ctypes.byref(out))
    check(st, "calcq")
    return out.value


def calcp(x, y, hyp, hypp, xtrainp, ytrainp, ztrainp, kyinvp, xtrain, ytrain, ztrain, kyinv, family="product", per=0.5,
          solver="hybrd"):
    """sympgpr.calcp -- sympgpr.f90:88-125 (hybrd1, tol 1e-13, started at guessP)."""
    hyp, hypp = _hyp3(hyp), _hyp3(hypp, "hypp")
    xtp, ytp, ztp = as_f64(xtrainp).ravel(), as_f64(ytrainp).ravel(), as_f64(ztrainp).ravel()
    xt, yt, zt = as_f64(xtrain).ravel(), as_f64(ytrain).ravel(), as_f64(ztrain).ravel()
    kyip, kyi = as_f64_fortran(kyinvp), as_f64_fortran(kyinv)
    np_, nt = xtp.size, xt.size
    if kyip.shape != (np_, np_) or kyi.shape != (2 * nt, 2 * nt) or zt.size != 2 * nt or ztp.size != np_:
        raise ValueError("calcp: inconsistent training-set shapes")
    out = ctypes.c_double(0.0)
    st = _lib.lib().sgp_calcp(_lib.context().handle, _fam(family), per, _solver(solver), _scalar(x, "x"), _scalar(y, "y"),
                              dptr(hyp), dptr(hypp), dptr(xtp), dptr(ytp), dptr(ztp), dptr(kyip), np_, dptr(xt), dptr(yt),
                              dptr(zt), dptr(kyi), nt, ctypes.byref(out))
    check(st, "calcp")
    return out.value


def guessP(x, y, hypp, xtrainp, ztrainp, Kyinvp, family="product", per=0.5):
    """python/02_pert_pendulum/func.py:207-210."""
    xtrainp = np.asarray(xtrainp)
    Ntrain = len(xtrainp) // 2
    return guessp(x, y, hypp, xtrainp[0:Ntrain], xtrainp[Ntrain:], ztrainp, Kyinvp, family, per)


def calcQ(x, y, xtrain, l, Kyinv, ztrain, family="product", per=0.5):
    """python/02_pert_pendulum/func.py:213-216."""
    xtrain = np.asarray(xtrain)
    Ntrain = len(xtrain) // 2
    return calcq(x, y, xtrain[:Ntrain], xtrain[Ntrain:], l, Kyinv, ztrain, family, per)


def calcP(x, y, l, hypp, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family="product", per=0.5, solver="hybrd"):
    """python/02_pert_pendulum/func.py:218-222."""
    xtrain, xtrainp = np.asarray(xtrain), np.asarray(xtrainp)
    Ntrain = len(xtrain) // 2
    Ntrainp = len(xtrainp) // 2
    return calcp(x, y, l, hypp, xtrainp[:Ntrainp], xtrainp[Ntrainp:], ztrainp, Kyinvp, xtrain[:Ntrain], xtrain[Ntrain:],
                 ztrain, Kyinv, family, per, solver)


def applymap_tok_f2py(hyp, hypp, q0map, p0map, xtrainp, ytrainp, ztrainp, kyinvp, xtrain, ytrain, ztrain, kyinv, qmap, pmap,
                      nm=None, ntest=None, family="product", per=0.5, solver="hybrd", kind="tokamak"):
    """sympgpr.applymap_tok(hyp, hypp, q0map, p0map, ..., qmap, pmap[, nm, ntest]) -- in place on
    (nm, ntest, 1) Fortran-ordered arrays."""
    hyp, hypp = _hyp3(hyp), _hyp3(hypp, "hypp")
    q0, p0 = as_f64(q0map).ravel(), as_f64(p0map).ravel()
    for a, nme in ((qmap, "qmap"), (pmap, "pmap")):
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.ndim != 3 or not a.flags.f_contiguous:
            raise ValueError(f"failed to initialize intent(inout) array -- `{nme}` must be float64, rank 3, Fortran order")
    nm_ = qmap.shape[0] if nm is None else int(nm)
    nt_ = q0.size if ntest is None else int(ntest)
    if qmap.shape != (nm_, nt_, 1) or pmap.shape != (nm_, nt_, 1) or p0.size != nt_ or q0.size != nt_:
        raise ValueError("applymap_tok: shape(qmap, 0) == nm and shape(q0map, 0) == ntest must hold")
    xtp, ytp, ztp = as_f64(xtrainp).ravel(), as_f64(ytrainp).ravel(), as_f64(ztrainp).ravel()
    xt, yt, zt = as_f64(xtrain).ravel(), as_f64(ytrain).ravel(), as_f64(ztrain).ravel()
    kyip, kyi = as_f64_fortran(kyinvp), as_f64_fortran(kyinv)
    st = _lib.lib().sgp_applymap_tok(_lib.context().handle, _fam(family), per, _solver(solver), _kind(kind), nm_, nt_,
                                     dptr(hyp), dptr(hypp), dptr(q0), dptr(p0), dptr(xtp), dptr(ytp), dptr(ztp), dptr(kyip),
                                     xtp.size, dptr(xt), dptr(yt), dptr(zt), dptr(kyi), xt.size, dptr(qmap), dptr(pmap))
    check(st, "applymap_tok")


def _applymap(kind, nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family, per, solver,
              alphap=None, alpha=None, out_every=1, want_pdiff=False, return_stats=False):
    if _solver(solver) == SOLVERS["explicit"]:
        # no guess GP in the explicit maps: an empty ordinary-GP model
        hypp, xtrainp, alphap = [1.0, 1.0, 1.0], np.zeros(0), np.zeros(0)
    hyp, hypp = _hyp3(l), _hyp3(hypp, "hypp")
    q0, p0 = as_f64(Q0map).ravel(), as_f64(P0map).ravel()
    E = int(Ntest)
    if q0.size < E or p0.size < E:
        raise ValueError("applymap: fewer initial conditions than Ntest")
    xtrainp, xtrain = as_f64(xtrainp).ravel(), as_f64(xtrain).ravel()
    np_, nt = xtrainp.size // 2, xtrain.size // 2
    if alphap is None:
        alphap = as_f64(Kyinvp).dot(as_f64(ztrainp).ravel())     # the constant vector the reference
    if alpha is None:                                            # recomputes per call (sympgpr.f90:72,85,121)
        alpha = as_f64(Kyinv).dot(as_f64(ztrain).ravel())
    alphap, alpha = as_f64(alphap).ravel(), as_f64(alpha).ravel()
    if alphap.size != np_ or alpha.size != 2 * nt:
        raise ValueError("applymap: alpha vectors do not match the training sets")
    nm = int(nm)
    rows = 1 + (nm - 1) // out_every if out_every > 0 else 0
    qmap = np.zeros((rows, E))
    pmap = np.zeros((rows, E))
    pdiff = np.zeros((rows, E)) if want_pdiff else None
    qf, pf = np.zeros(E), np.zeros(E)
    stats = (ctypes.c_ulonglong * 2)()
    st = _lib.lib().sgp_applymap(
        _lib.context().handle, _kind(kind), _fam(family), per, _solver(solver), nm, E, dptr(q0), dptr(p0), dptr(hyp),
        dptr(hypp), dptr(xtrainp[:np_]), dptr(np.ascontiguousarray(xtrainp[np_:2 * np_])), dptr(alphap), np_,
        dptr(xtrain[:nt]), dptr(np.ascontiguousarray(xtrain[nt:2 * nt])), dptr(alpha), nt,
        dptr(qmap) if rows else _NULL, dptr(pmap) if rows else _NULL, dptr(pdiff) if want_pdiff and rows else _NULL,
        out_every, dptr(qf), dptr(pf), stats)
    check(st, "applymap")
    out = [qmap, pmap] if rows else [qf, pf]
    if want_pdiff:
        out.append(pdiff)
    if return_stats:
        out.append(dict(evaluations=int(stats[0]), unconverged=int(stats[1]), qfinal=qf, pfinal=pf))
    return tuple(out)


def applymap(nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family="product", per=0.5,
             solver="hybrd", **kw):
    """python/functions/func.py:216-237 -> (qmap, pmap), each (nm, Ntest)."""
    return _applymap("pendulum", nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family,
                     per, solver, **kw)


def applymap_henon(nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family="sq", per=0.5,
                   solver="hybrd", **kw):
    """python/functions/func.py:239-260 (no wrap)."""
    return _applymap("henon", nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family, per,
                     solver, **kw)


def applymap_standard(nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family="product",
                      per=0.5, solver="hybrd", **kw):
    """python/04_standard_map/func.py:218-254 -> (qmap, pmap, pdiff)."""
    kw.setdefault("want_pdiff", True)
    return _applymap("standard", nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family,
                     per, solver, **kw)


def applymap_tok(nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family="product",
                 per=0.5, solver="hybrd", **kw):
    """python/05_tokamak/SympGPR/func.py:182-211 (NaN marks lost orbits)."""
    return _applymap("tokamak", nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv, family,
                     per, solver, **kw)


def applymap_tok_split(nphmap, nm, Ntest, Q0map, P0map, xtrainp, ztrainp, Kyinvp, hypp, xtrain, ztrain, Kyinv, hyp,
                       family="product", per=0.5, solver="hybrd", alphap=None, alpha=None, return_stats=False):
    """applymap_tok(nphmap, nm, Ntest, Q0map, P0map, xtrainp, ztrainp, Kyinvp, hypp, xtrain, ztrain, Kyinv, hyp)
    -> (qmap, pmap) -- python/05_tokamak/Split_SympGPR/func.py:184-219: nphmap learned maps applied in turn.

    Array layout as in python/05_tokamak/Split_SympGPR/main.py:91-112: xtrainp (2N, nphmap), ztrainp (N, nphmap),
    Kyinvp (nphmap, N, N), hypp (nphmap, 3), xtrain (2N, nphmap), ztrain (2N, nphmap), Kyinv (nphmap, 2N, 2N),
    hyp (nphmap, 3).  Like the reference loop (`while i < nm - nphmap`, whole turns only) the last rows of the
    (nm, Ntest) result stay zero when nm - 1 is not reached by whole turns."""
    nph, nm, E = int(nphmap), int(nm), int(Ntest)
    xtrainp, xtrain = np.asarray(xtrainp, float), np.asarray(xtrain, float)
    ztrainp, ztrain = np.asarray(ztrainp, float), np.asarray(ztrain, float)
    hyp, hypp = np.asarray(hyp, float).reshape(nph, 3), np.asarray(hypp, float).reshape(nph, 3)
    np_, nt = xtrainp.shape[0] // 2, xtrain.shape[0] // 2
    if alphap is None:
        alphap = np.stack([np.asarray(Kyinvp[m], float).dot(ztrainp[:, m]) for m in range(nph)])
    if alpha is None:
        alpha = np.stack([np.asarray(Kyinv[m], float).dot(ztrain[:, m]) for m in range(nph)])
    alphap, alpha = as_f64(alphap).reshape(nph, np_), as_f64(alpha).reshape(nph, 2 * nt)
    q0, p0 = as_f64(Q0map).ravel()[:E], as_f64(P0map).ravel()[:E]
    turns = -(-(nm - nph) // nph) if nm > nph else 0
    nsteps = turns * nph
    xp, yp = as_f64(xtrainp[:np_, :].T), as_f64(xtrainp[np_:2 * np_, :].T)          # (nph, np): sub-maps one after the other
    xs, ys = as_f64(xtrain[:nt, :].T), as_f64(xtrain[nt:2 * nt, :].T)
    qh, ph = np.zeros((nsteps + 1, E)), np.zeros((nsteps + 1, E))
    stats = (ctypes.c_ulonglong * 2)()
    st = _lib.lib().sgp_applymap_split(_lib.context().handle, _fam(family), per, _solver(solver), nph, nsteps, E, dptr(q0), dptr(p0),
                                       dptr(as_f64(hyp)), dptr(as_f64(hypp)), dptr(xp), dptr(yp), dptr(alphap), np_, dptr(xs),
                                       dptr(ys), dptr(alpha), nt, dptr(qh), dptr(ph), stats)
    check(st, "applymap_tok_split")
    qmap, pmap = np.zeros((nm, E)), np.zeros((nm, E))
    qmap[0], pmap[0] = q0, p0
    qmap[:nsteps + 1], pmap[:nsteps + 1] = qh, ph
    if return_stats:
        return qmap, pmap, dict(evaluations=int(stats[0]), unconverged=int(stats[1]))
    return qmap, pmap


def applymap_quality(kind, nm, Ntest, l, hypp, Q0map, P0map, xtrainp, ztrainp, Kyinvp, xtrain, ztrain, Kyinv,
                     energy="pendulum", energy_par=(1.0,), e_every=1, family="product", per=0.5, solver="hybrd",
                     alphap=None, alpha=None):
    """Map an ensemble nm - 1 steps and return the ingredients of the reference's `quality`
    (python/functions/func.py:262-272; python/05_tokamak/Split_SympGPR/func.py:221-232) without any history:
    dict(q1, p1, qfinal, pfinal, Eosc, Hmean, evaluations, unconverged) with Eosc = std(H)/mean(H) per orbit over
    the rows 0, e_every, ... of the history the reference would hold, accumulated inside the map kernel.
    energy="pendulum": H = p^2/2 + U0 (1 - cos(q + pi)), energy_par = (U0,)  (01_pendulum/implicit/func.py:116-117);
    energy="tokamak":  H = -Aph(compute_r([p 1e-2, q, 0], 0.3), q, 0), energy_par = (eps, m, phase)
    (Split_SympGPR/func.py:234-246)."""
    hyp, hypp = _hyp3(l), _hyp3(hypp, "hypp")
    E = int(Ntest)
    q0, p0 = as_f64(Q0map).ravel()[:E], as_f64(P0map).ravel()[:E]
    if q0.size < E or p0.size < E:
        raise ValueError("applymap_quality: fewer initial conditions than Ntest")
    xtrainp, xtrain = as_f64(xtrainp).ravel(), as_f64(xtrain).ravel()
    np_, nt = xtrainp.size // 2, xtrain.size // 2
    if alphap is None:
        alphap = as_f64(Kyinvp).dot(as_f64(ztrainp).ravel())
    if alpha is None:
        alpha = as_f64(Kyinv).dot(as_f64(ztrain).ravel())
    alphap, alpha = as_f64(alphap).ravel(), as_f64(alpha).ravel()
    if alphap.size != np_ or alpha.size != 2 * nt:
        raise ValueError("applymap_quality: alpha vectors do not match the training sets")
    epar = np.zeros(4)
    ep = as_f64(energy_par).ravel()
    epar[:ep.size] = ep
    ek = ENERGIES[energy] if isinstance(energy, str) else int(energy)
    q1, p1, qf, pf, eosc, hmean = (np.zeros(E) for _ in range(6))
    stats = (ctypes.c_ulonglong * 2)()
    st = _lib.lib().sgp_applymap_quality(
        _lib.context().handle, _kind(kind), _fam(family), per, _solver(solver), int(nm), E, dptr(q0), dptr(p0), dptr(hyp),
        dptr(hypp), dptr(xtrainp[:np_]), dptr(np.ascontiguousarray(xtrainp[np_:2 * np_])), dptr(alphap), np_,
        dptr(xtrain[:nt]), dptr(np.ascontiguousarray(xtrain[nt:2 * nt])), dptr(alpha), nt, ek, dptr(epar), int(e_every),
        dptr(q1), dptr(p1), dptr(qf), dptr(pf), dptr(eosc), dptr(hmean), stats)
    check(st, "applymap_quality")
    return dict(q1=q1, p1=p1, qfinal=qf, pfinal=pf, Eosc=eosc, Hmean=hmean, evaluations=int(stats[0]),
                unconverged=int(stats[1]))


def quality(q1, p1, Eosc, ysint, Nm, order="qp"):
    """(Eosc, gd, stdgd) as `quality` returns them (python/functions/func.py:262-272): gd[k] = mean squared distance of
    the first mapped state from the reference orbit ysint[Nm, :, k]; order="pq" for the tokamak variant, which compares
    (p, q) with ysint[Nm, 0:2, k] after wrapping the angle (Split_SympGPR/func.py:221-232)."""
    ys = np.asarray(ysint, float)
    if order == "pq":
        ref = np.array([ys[Nm, 0], np.mod(ys[Nm, 1], 2 * np.pi)])
        mine = np.array([p1, q1])
    else:
        ref = np.array([ys[Nm, 0], ys[Nm, 1]])
        mine = np.array([q1, p1])
    gd = np.mean((mine - ref)**2, axis=0)
    return np.asarray(Eosc), gd, float(np.std(gd))


def StandardMapIterate(k, nm, N, X0):
    """StandardMapIterate(k, nm, N, X0) -> f (2, N, nm) -- python/04_standard_map/main.py:32-39, on the device."""
    X0 = as_f64(X0)
    if X0.shape != (2, int(N)):
        raise ValueError("StandardMapIterate: X0 must have shape (2, N)")
    f = np.zeros((2, int(N), int(nm)))
    check(_lib.lib().sgp_standard_map_iterate(_lib.context().handle, float(k), int(nm), int(N), dptr(X0), dptr(f)),
          "StandardMapIterate")
    return f


def applymap4(nm, Ntest, hyp, Q0map, P0map, xtrain, alpha, out_every=1, return_stats=False):
    """2-DOF map prediction with the 4 x 4-block kernel (BASELINE config 3; not in the reference): Q0map, P0map (2, Ntest),
    xtrain = [q1; q2; P1; P2], alpha from fit(hyp4, xtrain, z, 4 N, reg=4)["alpha"], hyp = [lq, lP, sig].
    Returns (qmap, pmap) of shape (rows, 2, Ntest) -- or the final states (2, Ntest) with out_every=0."""
    hyp = _hyp3(hyp)
    E, nm = int(Ntest), int(nm)
    q0, p0 = as_f64(Q0map), as_f64(P0map)
    if q0.shape != (2, E) or p0.shape != (2, E):
        raise ValueError("applymap4: Q0map and P0map must have shape (2, Ntest)")
    xtrain, alpha = as_f64(xtrain).ravel(), as_f64(alpha).ravel()
    N = xtrain.size // 4
    if alpha.size != 4 * N:
        raise ValueError("applymap4: alpha does not match the training set")
    rows = 1 + (nm - 1) // out_every if out_every > 0 else 0
    qmap, pmap = np.zeros((rows, 2, E)), np.zeros((rows, 2, E))
    qf, pf = np.zeros((2, E)), np.zeros((2, E))
    stats = (ctypes.c_ulonglong * 2)()
    st = _lib.lib().sgp_applymap4(_lib.context().handle, nm, E, dptr(q0), dptr(p0), dptr(hyp), dptr(xtrain), dptr(alpha), N,
                                  dptr(qmap) if rows else _NULL, dptr(pmap) if rows else _NULL, out_every, dptr(qf), dptr(pf), stats)
    check(st, "applymap4")
    out = [qmap, pmap] if rows else [qf, pf]
    if return_stats:
        out.append(dict(evaluations=int(stats[0]), unconverged=int(stats[1]), qfinal=qf, pfinal=pf))
    return tuple(out)


def applymap_expl(nm, Ntest, l, Q0map, P0map, xtrain, ztrain, Kyinv, family="sum", per=0.5, **kw):
    """applymap_expl(nm, Ntest, l, Q0map, P0map, xtrain, ztrain, Kyinv) -> (qmap, pmap, pdiff) --
    python/04_standard_map/func.py:256-285: P = p - F_q(q, p) (calcP_expl :174-179), pdiff, p = mod(P, 2pi),
    q = q + dq (not wrapped)."""
    kw.setdefault("want_pdiff", True)
    return _applymap("standard_expl", nm, Ntest, l, None, Q0map, P0map, None, None, None, xtrain, ztrain, Kyinv, family, per,
                     "explicit", **kw)


def applymap_expl_pendulum(l, Q0map, P0map, xtrain, ztrain, Kyinv, Ntest, nm, family="sum", per=0.5, **kw):
    """applymap(l, Q0map, P0map, xtrain, ztrain, Kyinv, Ntest, nm) -> (qmap, pmap) --
    python/01_pendulum/explicit/func_expl.py:114-128: p = p - F_q(q, p), q = mod(q + dq, 2pi)."""
    return _applymap("pendulum", nm, Ntest, l, None, Q0map, P0map, None, None, None, xtrain, ztrain, Kyinv, family, per,
                     "explicit", **kw)


def calcP_expl(x, y, l, xtrain, ztrain, Kyinv, family="sum", per=0.5):
    """calcP_expl(x, y, l, xtrain, ztrain, Kyinv) -- python/04_standard_map/func.py:174-179."""
    out = _applymap("henon", 2, 1, l, None, [x], [y], None, None, None, xtrain, ztrain, Kyinv, family, per, "explicit",
                    out_every=0)
    return float(out[1][0])


# --------------------------------------------------------------------------- scalar forms / fieldlines
SCALAR_NAMES = (
    "kern_num", "dkdx_num", "dkdy_num", "dkdx0_num", "dkdy0_num", "d2kdxdx0_num", "d2kdydy0_num", "d2kdxdy0_num",
    "d3kdxdx0dy0_num", "d3kdydy0dy0_num", "d3kdxdy0dy0_num", "dkdlx_num", "dkdly_num", "d3kdxdx0dlx_num",
    "d3kdydy0dlx_num", "d3kdxdy0dlx_num", "d3kdxdx0dly_num", "d3kdydy0dly_num", "d3kdxdy0dly_num")


def scalar_function(family, name, period_arg=False):
    """One of the 19 `*_num` functions of a kernels module: f(x_a, y_a, x_b, y_b, lx, ly[, p])."""
    fam = _fam(family)
    which = SCALAR_NAMES.index(name)
    fn = _lib.lib().sgp_kernel_scalar

    if period_arg:
        def f(x_a, y_a, x_b, y_b, lx, ly, p):
            return fn(fam, which, x_a, y_a, x_b, y_b, lx, ly, p)
    else:
        def f(x_a, y_a, x_b, y_b, lx, ly):
            return fn(fam, which, x_a, y_a, x_b, y_b, lx, ly, 0.5)
    f.__name__ = name
    return f


def compute_r(z, rstart):
    """fieldlines.compute_r(z(3), rstart) -- fieldlines.f90:94-107."""
    z = as_f64(z).ravel()
    if z.size != 3:
        raise ValueError("compute_r: 0-th dimension must be fixed to 3")
    return _lib.lib().sgp_compute_r(z[0], z[1], z[2], float(rstart))


def ath(r, th, ph):
    """fieldlines.ath -- fieldlines.f90:34-39."""
    return _lib.lib().sgp_ath(float(r), float(th), float(ph))


# --------------------------------------------------------------------------- dense blocks (tests / profiling)
def spd_factor(A, want_factor=True, want_inverse=False):
    A = as_f64_fortran(A)
    n = A.shape[0]
    L = np.zeros((n, n), order="F") if want_factor else None
    Ai = np.zeros((n, n), order="F") if want_inverse else None
    ld = ctypes.c_double(0.0)
    st = _lib.lib().sgp_spd_factor(_lib.context().handle, dptr(A), n, dptr(L) if want_factor else _NULL,
                                   dptr(Ai) if want_inverse else _NULL, ctypes.byref(ld))
    check(st, "cholesky")
    return L, Ai, ld.value


def selftest_gemm(al, bl, mode, Mt, Nt, K):
    err = ctypes.c_double(0.0)
    check(_lib.lib().sgp_selftest_gemm(_lib.context().handle, al, bl, mode, Mt, Nt, K, ctypes.byref(err)), "selftest_gemm")
    return err.value


def gemm_host(al, bl, mode, A, B, C, alpha=1.0, beta=0.0):
    """The library's DMMA GEMM on host operands, in place on the Fortran-ordered C (M x N, multiples of 128):
    C = beta C + alpha A Bt restricted to the tile set / k-ranges of `mode` (csrc/dmma_gemm.cuh).  A is (M, K) Fortran
    order for al = 0 and (K, M) Fortran order for al = 1 (i.e. element (m, k) at [k, m]); B likewise with N."""
    A, B = as_f64_fortran(A), as_f64_fortran(B)
    if not (isinstance(C, np.ndarray) and C.dtype == np.float64 and C.flags.f_contiguous):
        raise ValueError("gemm_host: C must be a Fortran-ordered float64 array")
    M, N = C.shape
    K = A.shape[1] if al == 0 else A.shape[0]
    check(_lib.lib().sgp_gemm_host(_lib.context().handle, int(al), int(bl), int(mode), M // 128, N // 128, K, float(alpha),
                                   float(beta), dptr(A), A.shape[0], dptr(B), B.shape[0], dptr(C), C.shape[0]), "gemm_host")
    return C


def ozaki_gemm(A, B, C=None, alpha=1.0, beta=0.0, slices=7):
    """OPT-IN: C = alpha A B^T + beta C with the products formed on the INT8 tensor pipe (tcgen05.mma kind::i8) from `slices`
    signed 7-bit slices per operand (csrc/ozaki.cu).  A (M, K), B (N, K); returns the Fortran-ordered (M, N) result."""
    A, B = as_f64_fortran(A), as_f64_fortran(B)
    M, K = A.shape
    N = B.shape[0]
    if B.shape[1] != K:
        raise ValueError("ozaki_gemm: A is (M, K) and B must be (N, K)")
    out = np.zeros((M, N), order="F") if C is None else np.array(C, dtype=np.float64, order="F")
    check(_lib.lib().sgp_ozaki_gemm_host(_lib.context().handle, int(slices), M, N, K, float(alpha), dptr(A), M, dptr(B), N, float(beta),
                                         dptr(out), M), "ozaki_gemm")
    return out


def ozaki_gemm_ex(A, B, M, N, K, la=0, ta=0, lb=0, tb=0, C=None, alpha=1.0, beta=0.0, kmode=0, lower=0, slices=7):
    """Test hook of the sliced product the INT8 recursion is made of (sgp_ozaki_gemm_host_ex): A, B are Fortran-ordered 2-D
    arrays holding the operands in storage order la / lb (0: A[r, k]; 1: A[k, r]) with validity ta / tb (1: k <= r, 2: k >= r;
    the other half may hold anything) -- returns the Fortran-ordered (M, N) result."""
    A, B = as_f64_fortran(A), as_f64_fortran(B)
    out = np.zeros((M, N), order="F") if C is None else np.array(C, dtype=np.float64, order="F")
    check(_lib.lib().sgp_ozaki_gemm_host_ex(_lib.context().handle, int(slices), M, N, K, float(alpha), dptr(A), A.shape[0], int(la), int(ta),
                                            dptr(B), B.shape[0], int(lb), int(tb), float(beta), dptr(out), M, int(kmode), int(lower)),
          "ozaki_gemm_ex")
    return out
