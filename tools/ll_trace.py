"""Scratch: per-task time stamps of potrf_ll_kernel.  SGP_LL_TRACE=1 python tools/ll_trace.py N"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ["SGP_LL_TRACE"] = "1"
os.environ.setdefault("SGP_LL_TRACE_FILE", "gpurun_out/potrf_ll_trace.txt")
from sympgpr_b200 import api, workloads as W
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = W.standard_map_training(N); hyp = W.timing_hyp(N, d["sig"], 1e-8)
for _ in range(2):
    v = api.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N)
t = np.loadtxt(os.environ["SGP_LL_TRACE_FILE"])
diag = t[t[:, 1] == t[:, 2]]
diag = diag[np.argsort(diag[:, 1])]
print("diag tasks: j, start, mainloop_end-start, Sbuilt-ml, factor, invert, rest, end   (us)")
for r in diag[:: max(1, len(diag) // 16)]:
    j = int(r[1]); s = r[3:9] / 1e3
    print(f"{j:4d} start {s[0]:9.1f}  ml {s[1]-s[0]:8.1f}  S {s[2]-s[1]:6.1f}  fac {s[3]-s[2]:6.1f}  inv {s[4]-s[3]:6.1f}  out {s[5]-s[4]:6.1f}  end {s[5]:9.1f}  warp32 {r[9]/1e3:6.1f}")
ends = diag[:, 8] / 1e3
print("diag-to-diag interval (us): mean %.1f  min %.1f  max %.1f" % (np.diff(ends).mean(), np.diff(ends).min(), np.diff(ends).max()))
off = t[t[:, 1] == t[:, 2] + 1]
off = off[np.argsort(off[:, 2])]
for r in off[:: max(1, len(off) // 8)]:
    s = r[3:9] / 1e3
    print(f"sub-diag ({int(r[1])},{int(r[2])}) start {s[0]:9.1f} ml {s[1]-s[0]:8.1f} cstore {s[2]-s[1]:6.1f} solve+store {s[5]-s[2]:6.1f} end {s[5]:9.1f}")
print("total", t[:, 3:9].max() / 1e3, "us")
