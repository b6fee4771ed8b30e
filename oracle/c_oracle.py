"""TEST INFRASTRUCTURE -- ctypes loader for the C oracle (oracle/csrc/sympgpr_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

from .kernel_forms import FAMILY_IDS

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    src = os.path.join(_HERE, "csrc", "sympgpr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "_build/liboracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_guessp.restype = ctypes.c_double
        L.oracle_calcq.restype = ctypes.c_double
        L.oracle_calcp_alpha.restype = ctypes.c_double
        L.oracle_target_alpha.restype = ctypes.c_double
        L.oracle_compute_r.restype = ctypes.c_double
        L.oracle_applymap_alpha.restype = ctypes.c_long
        L.oracle_applymap_alpha_ex.restype = ctypes.c_long
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(c_dp)


def _fam(family):
    return FAMILY_IDS[family]


def build_k(x, y, x0, y0, hyp, rows=None, cols=None, family="product", p=0.5):
    x, y, x0, y0, hyp = map(_d, (x, y, x0, y0, hyp))
    rows = 2 * len(x) if rows is None else rows
    cols = 2 * len(x0) if cols is None else cols
    K = np.zeros((rows, cols), order="F")
    lib().oracle_build_k(_fam(family), ctypes.c_double(p), _p(x), _p(y), _p(x0), _p(y0), _p(hyp), _p(K),
                         ctypes.c_long(rows), ctypes.c_long(cols), ctypes.c_long(rows))
    return K


def buildkreg(x, y, x0, y0, hyp, family="product", p=0.5):
    x, y, x0, y0, hyp = map(_d, (x, y, x0, y0, hyp))
    K = np.zeros((len(x), len(x0)), order="F")
    lib().oracle_buildkreg(_fam(family), ctypes.c_double(p), _p(x), _p(y), _p(x0), _p(y0), _p(hyp), _p(K),
                           ctypes.c_long(len(x)), ctypes.c_long(len(x0)), ctypes.c_long(len(x)))
    return K


def guessp(x, y, hypp, xtp, ytp, ztp, kyinvp, family="product", p=0.5):
    hypp, xtp, ytp, ztp = map(_d, (hypp, xtp, ytp, ztp))
    kyinvp = np.asfortranarray(kyinvp, dtype=np.float64)
    return lib().oracle_guessp(_fam(family), ctypes.c_double(p), ctypes.c_double(x), ctypes.c_double(y), _p(hypp),
                               _p(xtp), _p(ytp), _p(ztp), _p(kyinvp), ctypes.c_long(len(xtp)))


def calcq(x, y, xt, yt, hyp, kyinv, zt, family="product", p=0.5):
    hyp, xt, yt, zt = map(_d, (hyp, xt, yt, zt))
    kyinv = np.asfortranarray(kyinv, dtype=np.float64)
    return lib().oracle_calcq(_fam(family), ctypes.c_double(p), ctypes.c_double(x), ctypes.c_double(y), _p(xt),
                              _p(yt), _p(hyp), _p(kyinv), _p(zt), ctypes.c_long(len(xt)))


def calcp_alpha(x, y, hyp, hypp, xtp, ytp, alphap, xt, yt, alpha, family="product", p=0.5):
    hyp, hypp, xtp, ytp, alphap, xt, yt, alpha = map(_d, (hyp, hypp, xtp, ytp, alphap, xt, yt, alpha))
    info = ctypes.c_int(0)
    nfev = ctypes.c_int(0)
    P = lib().oracle_calcp_alpha(_fam(family), ctypes.c_double(p), ctypes.c_double(x), ctypes.c_double(y),
                                 _p(hyp), _p(hypp), _p(xtp), _p(ytp), _p(alphap), ctypes.c_long(len(xtp)),
                                 _p(xt), _p(yt), _p(alpha), ctypes.c_long(len(xt)),
                                 ctypes.byref(info), ctypes.byref(nfev))
    return P, info.value, nfev.value


def target_alpha(q, pp, P, hyp, xt, yt, alpha, family="product", p=0.5):
    hyp, xt, yt, alpha = map(_d, (hyp, xt, yt, alpha))
    return lib().oracle_target_alpha(_fam(family), ctypes.c_double(p), ctypes.c_double(q), ctypes.c_double(pp),
                                     ctypes.c_double(P), _p(hyp), _p(xt), _p(yt), _p(alpha),
                                     ctypes.c_long(len(xt)))


def compute_r(pth, th, rstart=0.3):
    return lib().oracle_compute_r(ctypes.c_double(pth), ctypes.c_double(th), ctypes.c_double(rstart))


def applymap_alpha(kind, nm, q0, p0, hyp, hypp, xtp, ytp, alphap, xt, yt, alpha, family="product", p=0.5,
                   want_pdiff=False, want_notconv=False, start_delta=False, reverse_sum=False, out_every=1):
    """Returns qmap, pmap (nm, E) [, pdiff], mean function evaluations per orbit-step
    [, per-orbit count of steps where hybrd1 returned info != 1, per-orbit largest
    residual |f(P)| of an accepted root].

    start_delta: hybrd1 is started at p + guess instead of the bare guess (sympgpr.f90:103-107 with a guess GP
    trained on P - p, python/04_standard_map/main.py:89-90) -- the oracle twin of the library's "newton_delta" start.
    reverse_sum: the training-set sums run backwards (rounding-sensitivity probe).  out_every = k > 1: only the
    rows 0, k, 2k, ... of the history are returned ((nm - 1) // k + 1 rows)."""
    q0, p0, hyp, hypp, xtp, ytp, alphap, xt, yt, alpha = map(
        _d, (q0, p0, hyp, hypp, xtp, ytp, alphap, xt, yt, alpha))
    E = len(q0)
    ex = bool(start_delta or reverse_sum or out_every != 1)
    rows = (nm - 1) // out_every + 1 if ex else nm
    qmap = np.zeros((rows, E)); pmap = np.zeros((rows, E))
    pdiff = np.zeros((rows, E)) if want_pdiff else None
    notconv = np.zeros(max(E, 1), dtype=np.int32)
    maxres = np.zeros(max(E, 1))
    args = (kind, _fam(family), ctypes.c_double(p), ctypes.c_long(nm), ctypes.c_long(E), _p(q0), _p(p0),
            _p(hyp), _p(hypp), _p(xtp), _p(ytp), _p(alphap), ctypes.c_long(len(xtp)),
            _p(xt), _p(yt), _p(alpha), ctypes.c_long(len(xt)), _p(qmap), _p(pmap),
            _p(pdiff) if want_pdiff else None, notconv.ctypes.data_as(c_ip), _p(maxres))
    if ex:
        fev = lib().oracle_applymap_alpha_ex(*args, ctypes.c_int((1 if start_delta else 0) | (2 if reverse_sum else 0)),
                                             ctypes.c_long(out_every))
    else:
        fev = lib().oracle_applymap_alpha(*args)
    nev = fev / max(1, E * (nm - 1))
    out = [qmap, pmap]
    if want_pdiff:
        out.append(pdiff)
    out.append(nev)
    if want_notconv:
        out.append(notconv[:E])
        out.append(maxres[:E])
    return tuple(out)


def num_threads():
    return lib().oracle_num_threads()
