"""Device-resident NLL+gradient timings with stage times at small/mid sizes (potrf chain experiments).
   [SYMPGPR_B200_LIB=...] python tools/potrf_perf.py [N ...]"""
import ctypes, json, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from sympgpr_b200 import _lib, workloads as W
Ns = [int(x) for x in sys.argv[1:]] or [200, 1024, 2048, 4096, 8192]
L = _lib.lib(); ctx = _lib.context(0)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
_lib.check(L.sgp_set_profiling(ctx.handle, 1), "prof")
import os
OZ = int(os.environ.get("SGP_OZAKI", "0"))
ctx.set_ozaki(OZ)
for N in Ns:
    d = W.standard_map_training(N); hyp = W.timing_hyp(N, d["sig"], 1e-8)
    x = torch.from_numpy(d["xtrain"].copy()).to(dev); z = torch.from_numpy(d["ztrain"].copy()).to(dev)
    res = torch.zeros(16, dtype=torch.float64, device=dev); hc = (ctypes.c_double * 4)(*hyp)
    f = lambda: _lib.check(L.sgp_nll_dev(ctx.handle, 0, 0.5, 0, hc, x.data_ptr(), z.data_ptr(), 2 * N, 2, res.data_ptr()), "nll")
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if N <= 4096 else 5
    e0.record(st)
    for _ in range(reps): f()
    e1.record(st); e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sm = (ctypes.c_double * 7)(); _lib.check(L.sgp_stage_times(ctx.handle, sm), "st")
    n = 2.0 * N
    print(json.dumps({"lib": _lib.LIB_PATH.split("/")[-1], "ozaki": OZ, "N": N, "n": 2 * N, "ms": round(ms, 4), "TF": round(n**3 / ms / 1e9, 2),
                      "potrf": round(sm[1], 4), "trtri": round(sm[3], 4), "lauum": round(sm[4], 4), "us_per_tile_col": round(1e3 * sm[1] / max(1, (2 * N + 127) // 128), 1),
                      "nll": float(res[0].item()), "g": [float(res[1].item()), float(res[2].item())]}), flush=True)
