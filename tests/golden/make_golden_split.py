#!/usr/bin/env python3
"""Generate tests/golden/path_split.npz: the REFERENCE's own split-map loop (applymap_tok of
python/05_tokamak/Split_SympGPR/func.py:184-219, imported unmodified) run in the build container.

That file needs the compiled f2py modules `sympgpr` and `fieldlines`, which cannot be built here (no
Fortran compiler).  They are replaced by adapters with the f2py signatures whose
arithmetic is the CPU oracle's -- itself pinned to the reference's Python layer by path_product.npz --
so what this fixture pins is the LOOP: cycling through the nphmap learned maps, whole turns only
(`while i < nm - nphmap`), NaN propagation, the loss test at the new angle.

    python tests/golden/make_golden_split.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden_path import REF, load, reference_kernels_module   # noqa: E402
from oracle import oracle as O                                       # noqa: E402


def main():
    sys.modules["kernels"] = reference_kernels_module(f"{REF}/05_tokamak/Split_SympGPR/init_func.py", "kernels")

    class _Sym:          # f2py surface of sympgpr.f90, arithmetic by the oracle
        @staticmethod
        def build_k(x, y, x0, y0, hyp, K): O.build_k(x, y, x0, y0, hyp, K)
        @staticmethod
        def buildkreg(x, y, x0, y0, hyp, K): O.buildkreg(x, y, x0, y0, hyp, K)
        @staticmethod
        def guessp(x, y, hypp, xtp, ytp, ztp, kyinvp): return O.guessp(x, y, hypp, xtp, ytp, ztp, kyinvp)
        @staticmethod
        def calcq(x, y, xt, yt, hyp, kyinv, zt): return O.calcq(x, y, xt, yt, hyp, kyinv, zt)
        @staticmethod
        def calcp(x, y, hyp, hypp, xtp, ytp, ztp, kyinvp, xt, yt, zt, kyinv):
            return O.calcp(x, y, hyp, hypp, xtp, ytp, ztp, kyinvp, xt, yt, zt, kyinv)

    class _Fl:
        @staticmethod
        def compute_r(z, rstart): return O.compute_r(z, rstart)

    ms = types.ModuleType("sympgpr"); ms.sympgpr = _Sym
    mf = types.ModuleType("fieldlines"); mf.fieldlines = _Fl
    sys.modules["sympgpr"] = ms
    sys.modules["fieldlines"] = mf
    fs = load(f"{REF}/05_tokamak/Split_SympGPR/func.py", "ref_split_func")

    nph, N = 4, 24
    xtp = np.zeros((2 * N, nph)); ztp = np.zeros((N, nph)); xt = np.zeros((2 * N, nph)); zt = np.zeros((2 * N, nph))
    hyp = np.zeros((nph, 3)); hypp = np.zeros((nph, 3)); Kyinv = np.zeros((nph, 2 * N, 2 * N)); Kyinvp = np.zeros((nph, N, N))
    for m in range(nph):
        q = O.halton(N, 2, start=1 + 7 * m) * 2 * np.pi
        p = 0.5 + O.halton(N, 3, start=1 + 5 * m) * 5.0
        P = p + 0.1 * (1 + 0.2 * m) * np.sin(q)
        Q = q + 0.25 * P
        xt[:, m] = np.hstack((q, P)); zt[:, m] = np.concatenate((p - P, Q - q))
        xtp[:, m] = np.hstack((q, p)); ztp[:, m] = P
        l = 1.2 + 0.1 * m
        hyp[m] = [l, l, 2 * np.max(np.abs(zt[:, m]))**2]
        hypp[m] = [l, l, 2 * np.max(np.abs(ztp[:, m]))**2]
        Kyinv[m] = np.linalg.inv(O.build_k_vec(q, P, q, P, hyp[m]) + 1e-6 * np.eye(2 * N))
        Kyinvp[m] = np.linalg.inv(O.buildkreg_vec(q, p, q, p, hypp[m]) + 1e-6 * np.eye(N))
    E, nm = 9, 14                       # 14 - 4 = 10 -> three whole turns = 12 steps; row 13 stays zero
    q0 = O.halton(E, 5) * 2 * np.pi
    p0 = 0.3 + O.halton(E, 7) * 4.5
    qmap, pmap = fs.applymap_tok(nph, nm, E, q0, p0, xtp, ztp, Kyinvp, hypp, xt, zt, Kyinv, hyp)
    np.savez(os.path.join(HERE, "path_split.npz"), nph=np.array([nph]), N=np.array([N]), xtp=xtp, ztp=ztp, xt=xt, zt=zt,
             hyp=hyp, hypp=hypp, Kyinv=Kyinv, Kyinvp=Kyinvp, q0=q0, p0=p0, qmap=qmap, pmap=pmap)
    print("wrote path_split.npz; NaNs:", int(np.isnan(pmap).sum()), "zero rows:", [i for i in range(nm) if np.all(qmap[i] == 0)])


if __name__ == "__main__":
    main()
