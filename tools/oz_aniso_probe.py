import sys; sys.path.insert(0, ".")
import numpy as np
from oracle import oracle as O
from sympgpr_b200 import _lib, api
ctx = _lib.context()
N = 1024
d = O.standard_map_training(N)
base = O.timing_hyp(N, d["sig"], 1e-8)
for fx, fy in ((1, 1), (0.2, 5), (5, 0.2), (0.05, 5), (0.02, 10)):
    hyp = base.copy(); hyp[0] *= fx; hyp[1] *= fy
    try:
        ctx.set_ozaki_ex(0, 1, 0)
        v0, g0 = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    except Exception as e:
        print(fx, fy, "DMMA failed", type(e).__name__); continue
    row = [f"lx x{fx} ly x{fy}: nll {v0:.6e}"]
    for ns in (6, 7, 8):
        ctx.set_ozaki_ex(ns, 3, 256)
        try:
            v, g = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
            row.append(f"{ns}d: {abs(v - v0) / abs(v0):.1e} {np.max(np.abs(g - g0) / np.abs(g0)):.1e}")
        except Exception as e:
            row.append(f"{ns}d: {type(e).__name__}")
    print("  ".join(row), flush=True)
ctx.set_ozaki_ex(0, 1, 0)
