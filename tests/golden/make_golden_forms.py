#!/usr/bin/env python3
"""Generate tests/golden/kernel_forms_<family>.npz from the REFERENCE's own
SymPy derivation (run in the build container only; /root/reference does not
exist on the GPU box, the .npz fixtures are what travels).

For each kernel family the reference's ``init_func.py`` is executed verbatim
(in a scratch directory, because it writes ``kernels*.f90`` into the cwd) and
the SymPy expression objects it builds (``kern``, ``dkdxadxb`` ...) are taken
from its namespace.  They are evaluated with mpmath at 50 digits on seeded
random points and rounded to float64: that is the known-answer value of each
of the 19 ``*_num`` functions the generated Fortran computes.

    python tests/golden/make_golden_forms.py
"""
import os
import runpy
import sys
import tempfile

import numpy as np

REF = "/root/reference/python"
HERE = os.path.dirname(os.path.abspath(__file__))

# family -> (init_func.py, has free period argument)
SOURCES = {
    "product": (f"{REF}/04_standard_map/init_func.py", False),
    "sq": (f"{REF}/03_henon_heiles/init_func.py", False),
    "sum": (f"{REF}/01_pendulum/explicit/init_func.py", False),
    "period": (f"{REF}/01_pendulum/implicit_period_unknown/init_func.py", True),
}

# name in the generated Fortran -> variable name in init_func.py
# (python/04_standard_map/init_func.py:58-76)
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


def run_family(family, path, has_p, npts=48, seed=1234):
    import mpmath
    import sympy

    mpmath.mp.dps = 50
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            ns = runpy.run_path(path)
        finally:
            os.chdir(cwd)
    syms = [ns["xa"], ns["ya"], ns["xb"], ns["yb"], ns["lx"], ns["ly"]]
    if has_p:
        syms.append(ns["p"])
    rng = np.random.default_rng(seed)
    pts = np.empty((npts, len(syms)))
    pts[:, 0] = rng.uniform(-1.0, 7.0, npts)     # x_a
    pts[:, 1] = rng.uniform(-3.0, 3.0, npts)     # y_a
    pts[:, 2] = rng.uniform(-1.0, 7.0, npts)     # x_b
    pts[:, 3] = rng.uniform(-3.0, 3.0, npts)     # y_b
    pts[:, 4] = rng.uniform(0.2, 3.0, npts)      # lx
    pts[:, 5] = rng.uniform(0.2, 3.0, npts)      # ly
    if has_p:
        pts[:, 6] = rng.uniform(0.2, 1.5, npts)  # p
    # first points: literal inputs of test_sympgpr.py:7-10,19 (x0/y0 as a, x/y as b)
    lit = [(1.0, 0.0, 1.0, 0.0), (2.0, 3.0, 1.0, 0.0), (1.0, 0.0, 3.0, 2.0), (2.0, 3.0, 2.0, 3.0)]
    for k, (xa_, ya_, xb_, yb_) in enumerate(lit):
        pts[k, :6] = (xa_, ya_, xb_, yb_, 0.5, 2.0)
    out = {"points": pts}
    for fname, var in EXPRS:
        expr = sympy.nsimplify(ns[var], rational=True)   # 0.5 -> 1/2, exact
        f = sympy.lambdify(syms, expr, modules="mpmath")
        vals = np.empty(npts)
        for k in range(npts):
            args = [mpmath.mpf(float(v)) for v in pts[k]]
            vals[k] = float(f(*args))
        out[fname] = vals
        print(f"  {family}.{fname}: ok", flush=True)
    np.savez(os.path.join(HERE, f"kernel_forms_{family}.npz"), **out)


if __name__ == "__main__":
    fams = sys.argv[1:] or list(SOURCES)
    for fam in fams:
        print(f"[{fam}] running {SOURCES[fam][0]}", flush=True)
        run_family(fam, *SOURCES[fam])
