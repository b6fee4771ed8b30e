"""Run a reference example script unchanged against libsympgpr_b200.

    python -m sympgpr_b200.runner /path/to/python/02_pert_pendulum/main.py [--family sq] [--solver newton_delta] [--int8-digits 7]

Registers the shim modules under the names the reference imports (`sympgpr`, `fortran.sympgpr`,
`kernels`, `kernels_sq`, `kernels_sum`, `fieldlines`), puts the reference's `python/` directory on
sys.path (for `functions`), changes into the script's directory and executes it with runpy.

The f2py build of the reference selects the kernel family by which kernels*.f90 it compiles into the module named
`kernels`; here that is a run-time choice.  By default it follows the example directory:
    03_henon_heiles                        -> family "sq"       (kernels_sq.f90)
    01_pendulum/implicit_period_unknown    -> family "period"   (7-argument functions f(..., lx, ly, p): `kernels` is bound
                                              to the period_arg=True variant of the shim)
    everything else                        -> family "product"  (kernels.f90)
Third-party packages the scripts import (matplotlib, tkinter, ghalton, cma) must be installed; this module does not
fake them (the test suite registers stand-ins from its own side, tests/harness/standins.py).
"""
import argparse
import os
import runpy
import sys

FAMILY_CHOICES = ["auto", "product", "sq", "sum", "period"]
SOLVER_CHOICES = ["hybrd", "newton", "newton_delta"]


def family_of(script_dir):
    d = script_dir.replace("\\", "/")
    if d.endswith("implicit_period_unknown"):
        return "period"
    if "03_henon_heiles" in d:
        return "sq"
    return "product"


def run(script, family="auto", per=0.5, solver="hybrd", run_name="__main__", int8_digits=0):
    import sympgpr_b200
    if solver not in SOLVER_CHOICES:
        raise ValueError(f"solver must be one of {SOLVER_CHOICES}")
    if int8_digits:
        # opt-in INT8 route for the script's NLL / fit calls: factor + inverse + lauum from INT8 digit products
        from sympgpr_b200 import _lib
        _lib.context().set_ozaki_ex(int(int8_digits), 3, 0)
    script = os.path.abspath(script)
    ex_dir = os.path.dirname(script)
    if family == "auto":
        family = family_of(ex_dir)
    if family not in FAMILY_CHOICES:
        raise ValueError(f"family must be one of {FAMILY_CHOICES}")
    mods = sympgpr_b200.install_shims()
    sym = mods["sympgpr"].sympgpr
    sym.family, sym.per, sym.solver = ("product" if family == "period" else family), per, solver
    if family == "period":
        # python/01_pendulum/implicit_period_unknown/func.py does `from kernels import *` and calls f(..., lx, ly, p)
        sys.modules["kernels"] = mods["kernels_period"]
    py_root = ex_dir
    while py_root != "/" and os.path.basename(py_root) != "python":
        py_root = os.path.dirname(py_root)
    added = []
    for p in (ex_dir, py_root):
        if p and p != "/" and p not in sys.path:
            sys.path.insert(1, p)
            added.append(p)
    cwd = os.getcwd()
    os.chdir(ex_dir)
    try:
        return runpy.run_path(script, run_name=run_name)
    finally:
        os.chdir(cwd)
        for p in added:
            if p in sys.path:
                sys.path.remove(p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("script")
    ap.add_argument("--family", default="auto", choices=FAMILY_CHOICES)
    ap.add_argument("--per", type=float, default=0.5)
    ap.add_argument("--solver", default="hybrd", choices=SOLVER_CHOICES)
    ap.add_argument("--int8-digits", type=int, default=0, choices=[0, 4, 5, 6, 7, 8],
                    help="opt-in INT8 route for the training stages (0 = off, the default DMMA route; 6: 47 bits, 7: 55 bits)")
    a = ap.parse_args()
    run(a.script, a.family, a.per, a.solver, int8_digits=a.int8_digits)


if __name__ == "__main__":
    main()
