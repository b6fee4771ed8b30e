"""Run a reference example script unchanged against libsympgpr_b200.

    python -m sympgpr_b200.runner /path/to/python/02_pert_pendulum/main.py [--family sq] [--solver newton]

Registers the shim modules under the names the reference imports (`sympgpr`, `fortran.sympgpr`,
`kernels`, `kernels_sq`, `kernels_sum`, `fieldlines`), puts the reference's `python/` directory on
sys.path (for `functions`), changes into the script's directory and executes it with runpy.
Third-party packages the scripts import (matplotlib, tkinter, ghalton, cma) must be installed;
this module does not fake them.
"""
import argparse
import os
import runpy
import sys


def run(script, family="product", per=0.5, solver="hybrd", run_name="__main__"):
    import sympgpr_b200
    mods = sympgpr_b200.install_shims()
    sym = mods["sympgpr"].sympgpr
    sym.family, sym.per, sym.solver = family, per, solver
    script = os.path.abspath(script)
    ex_dir = os.path.dirname(script)
    py_root = ex_dir
    while py_root != "/" and os.path.basename(py_root) != "python":
        py_root = os.path.dirname(py_root)
    for p in (ex_dir, py_root):
        if p and p != "/" and p not in sys.path:
            sys.path.insert(1, p)
    cwd = os.getcwd()
    os.chdir(ex_dir)
    try:
        return runpy.run_path(script, run_name=run_name)
    finally:
        os.chdir(cwd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("script")
    ap.add_argument("--family", default="product", choices=["product", "sq", "sum"])
    ap.add_argument("--per", type=float, default=0.5)
    ap.add_argument("--solver", default="hybrd", choices=["hybrd", "newton"])
    a = ap.parse_args()
    run(a.script, a.family, a.per, a.solver)


if __name__ == "__main__":
    main()
