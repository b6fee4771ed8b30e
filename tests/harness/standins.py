"""TEST HARNESS: stand-ins for the third-party packages the reference's example scripts import but this image does not
have ("Scripts cannot import here"): ghalton, tkinter, matplotlib (+ pyplot, mpl_toolkits.mplot3d), cma,
and the reference's own VODE-based `henon` extension (python/03_henon_heiles/henon.f90, a training-data generator
outside the hot path).  They are registered in sys.modules from the TEST side only; the product never ships or imports
them (sympgpr_b200/runner.py expects the real packages).

    from tests.harness import standins; standins.install()
"""
import sys
import types

import numpy as np


# ------------------------------------------------------------------------------------------ ghalton
def _radical_inverse(i, base):
    f, r = 1.0, 0.0
    while i > 0:
        f /= base
        r += f * (i % base)
        i //= base
    return r


_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29)


class _Halton:
    """ghalton.Halton(d): .get(n) -> the next n points of the unscrambled Halton sequence (bases 2, 3, 5, ...),
    starting at index 1, as a list of lists (the scripts multiply it by an ndarray)."""

    def __init__(self, dim):
        self.dim, self.next = int(dim), 1

    def get(self, n):
        out = [[_radical_inverse(i, _PRIMES[k]) for k in range(self.dim)] for i in range(self.next, self.next + int(n))]
        self.next += int(n)
        return out


class _GeneralizedHalton(_Halton):
    def __init__(self, dim, seed=0):
        super().__init__(dim)


# ------------------------------------------------------------------------------------------ matplotlib / tkinter
class _Anything:
    """Absorbs any attribute access / call / indexing / iteration the plotting code makes."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _AnyModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


def _subplots(nrows=1, ncols=1, *a, **k):
    n = int(nrows) * int(ncols)
    if n == 1:
        return _Anything(), _Anything()
    axes = np.empty(n, dtype=object)
    for i in range(n):
        axes[i] = _Anything()
    return _Anything(), (axes.reshape(int(nrows), int(ncols)) if int(nrows) > 1 and int(ncols) > 1 else axes)


# ------------------------------------------------------------------------------------------ cma
def _cma_fmin(objective, x0, sigma0, options=None, args=(), restarts=0, **kw):
    """cma.fmin(...) -> tuple whose element 0 is xbest (what Split_SympGPR/main.py reads): a derivative-free SciPy
    minimiser stands in for CMA-ES (the optimiser is outside the hot path; only its use of the NLL closure matters)."""
    import scipy.optimize
    opts = options or {}
    res = scipy.optimize.minimize(lambda x: float(objective(np.asarray(x), *args)), np.asarray(x0, float), method="Nelder-Mead",
                                  options={"maxfev": int(opts.get("maxfevals", 400)), "xatol": 1e-6, "fatol": 1e-9})
    return (res.x, res.fun, res.nfev, res.nfev, res.nit, res.x, np.zeros_like(res.x), None, None, None)


# ------------------------------------------------------------------------------------------ henon (VODE Poincare tracer)
class _HenonState:
    lam = 1.0
    tmax = 1000.0


_HENON_JIT = None


def _henon_kernel():
    """Fixed-step RK4 tracer with the cut located by a secant on the length of the last step; numba-compiled when numba is
    there (37 orbits x t = 5000 in about a second), plain Python otherwise."""
    global _HENON_JIT
    if _HENON_JIT is not None:
        return _HENON_JIT

    def rk4(z0, z1, z2, z3, h, lam):
        def f(a0, a1, a2, a3):
            return a2, a3, -a0 - 2.0 * lam * a0 * a1, -a1 - lam * (a0 * a0 - a1 * a1)
        k10, k11, k12, k13 = f(z0, z1, z2, z3)
        k20, k21, k22, k23 = f(z0 + 0.5 * h * k10, z1 + 0.5 * h * k11, z2 + 0.5 * h * k12, z3 + 0.5 * h * k13)
        k30, k31, k32, k33 = f(z0 + 0.5 * h * k20, z1 + 0.5 * h * k21, z2 + 0.5 * h * k22, z3 + 0.5 * h * k23)
        k40, k41, k42, k43 = f(z0 + h * k30, z1 + h * k31, z2 + h * k32, z3 + h * k33)
        return (z0 + h / 6.0 * (k10 + 2 * k20 + 2 * k30 + k40), z1 + h / 6.0 * (k11 + 2 * k21 + 2 * k31 + k41),
                z2 + h / 6.0 * (k12 + 2 * k22 + 2 * k32 + k42), z3 + h / 6.0 * (k13 + 2 * k23 + 2 * k33 + k43))

    def trace(z, lam, tmax, dt, tcut, zcut):
        ncut_max = tcut.shape[0]
        a0, a1, a2, a3 = z[0], z[1], z[2], z[3]
        t, icut = 0.0, 0
        while t < tmax and icut < ncut_max:
            b0, b1, b2, b3 = rk4(a0, a1, a2, a3, dt, lam)
            if a0 < 0.0 and b0 >= 0.0:
                # cut q1 = 0 with p1 > 0 inside this step: secant on the step length h, f(h) = q1(t + h)
                h0, f0, h1, f1 = 0.0, a0, dt, b0
                c0, c1, c2, c3 = b0, b1, b2, b3
                for _ in range(8):
                    if f1 == f0:
                        break
                    h2 = h1 - f1 * (h1 - h0) / (f1 - f0)
                    c0, c1, c2, c3 = rk4(a0, a1, a2, a3, h2, lam)
                    h0, f0, h1, f1 = h1, f1, h2, c0
                    if abs(c0) < 1e-15:
                        break
                tcut[icut] = t + h1
                zcut[0, icut] = 0.0
                zcut[1, icut] = c1
                zcut[2, icut] = c2
                zcut[3, icut] = c3
                icut += 1
            a0, a1, a2, a3 = b0, b1, b2, b3
            t += dt
        return icut
    try:
        import numba
        rk4 = numba.njit(cache=False)(rk4)
        _HENON_JIT = numba.njit(cache=False)(trace)
    except Exception:                                   # no numba: the plain-Python loop (slow but equivalent)
        _HENON_JIT = trace
    return _HENON_JIT


def _henon_integrate(z0):
    """henon.integrate(z0) -> (tcut, zcut, icut) as python/03_henon_heiles/henon.f90:34-87: Henon-Heiles orbit
    H = (p1^2 + p2^2)/2 + (q1^2 + q2^2)/2 + lam (q1^2 q2 - q2^3/3) traced from z0 = (q1, q2, p1, p2) up to tmax, cuts through
    q1 = 0 with p1 > 0 recorded: tcut (ncut), zcut (4, ncut), icut = number of cuts.  RK4 with dt = 0.005 (local error
    ~1e-13 per unit time at these energies) replaces DVODE (rtol 1e-12)."""
    tcut = np.zeros(1000)
    zcut = np.zeros((4, 1000), order="F")
    icut = _henon_kernel()(np.asarray(z0, float), float(_HenonState.lam), float(_HenonState.tmax), 0.005, tcut, zcut)
    return tcut, zcut, int(icut)


def install(with_henon=True):
    """Register the stand-ins (only for names that are not importable for real)."""
    import importlib.util

    def missing(name):
        if name in sys.modules:
            return False
        try:
            return importlib.util.find_spec(name) is None
        except (ImportError, ValueError):
            return True

    made = []
    if missing("ghalton"):
        m = types.ModuleType("ghalton")
        m.Halton, m.GeneralizedHalton = _Halton, _GeneralizedHalton
        sys.modules["ghalton"] = m
        made.append("ghalton")
    if missing("tkinter"):
        sys.modules["tkinter"] = _AnyModule("tkinter")
        made.append("tkinter")
    if missing("matplotlib"):
        mpl = _AnyModule("matplotlib")
        mpl.__path__ = []
        plt = _AnyModule("matplotlib.pyplot")
        plt.subplots = _subplots
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        for sub in ("cm", "colors", "ticker", "patches", "gridspec", "animation"):
            sm = _AnyModule("matplotlib." + sub)
            setattr(mpl, sub, sm)
            sys.modules["matplotlib." + sub] = sm
        tk = _AnyModule("mpl_toolkits")
        tk.__path__ = []
        m3 = _AnyModule("mpl_toolkits.mplot3d")
        tk.mplot3d = m3
        sys.modules["mpl_toolkits"] = tk
        sys.modules["mpl_toolkits.mplot3d"] = m3
        made.append("matplotlib")
    if missing("cma"):
        m = types.ModuleType("cma")
        m.fmin = _cma_fmin
        sys.modules["cma"] = m
        made.append("cma")
    if with_henon and "henon" not in sys.modules:
        m = types.ModuleType("henon")
        m.henon = _HenonState
        m.integrate = _henon_integrate
        sys.modules["henon"] = m
        made.append("henon")
    return made
