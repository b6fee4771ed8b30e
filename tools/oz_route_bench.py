"""INT8 (Ozaki) route of the NLL+gradient evaluation against the DMMA route: stage times and results, same inputs.
    python tools/oz_route_bench.py [N ...]     (env OZ_CONFIGS="ns:stages:leaf,..." default "0:0:0,8:1:0,7:3:4096,7:3:2048,8:3:4096")"""
import ctypes, json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
from sympgpr_b200 import _lib, api, workloads as W

L = _lib.lib(); ctx = _lib.context(0)
_lib.check(L.sgp_set_profiling(ctx.handle, 1), "sgp_set_profiling")
cfgs = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("OZ_CONFIGS", "0:0:0,8:1:0,7:3:4096,7:3:2048,8:3:4096").split(",")]
names = ["fill", "potrf", "potrs", "trtri", "lauum", "grad", "finalize"]
for N in [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384]:
    d = W.standard_map_training(N); hyp = W.timing_hyp(N, d["sig"], 1e-8)
    ref = None
    for ns, stages, leaf in cfgs:
        ctx.set_ozaki_ex(ns, stages if ns else 1, leaf)
        try:
            for _ in range(2):
                v, g = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
            t = time.perf_counter(); reps = 3
            for _ in range(reps):
                v, g = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
            ms = (time.perf_counter() - t) / reps * 1e3
            st = (ctypes.c_double * 7)()
            _lib.check(L.sgp_stage_times(ctx.handle, st), "sgp_stage_times")
            api.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N)
            t = time.perf_counter()
            for _ in range(reps):
                vv = api.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N)
            ms_v = (time.perf_counter() - t) / reps * 1e3
            row = {"N": N, "slices": ns, "stages": stages, "leaf": leaf, "ms_per_eval_e2e": round(ms, 3), "fp64_equiv_TFLOP/s": round((2.0 * N) ** 3 / ms / 1e9, 2),
                   "stages_ms": {k: round(float(x), 3) for k, x in zip(names, st)}, "nll": v, "grad": [float(g[0]), float(g[1])],
                   "value_only_ms_e2e": round(ms_v, 3), "value_only_nll": vv}
            if ref is None:
                ref = (v, np.asarray(g))
            else:
                row["nll_rel_diff_vs_dmma"] = abs(v - ref[0]) / abs(ref[0])
                row["grad_rel_diff_vs_dmma"] = float(np.max(np.abs(np.asarray(g) - ref[1]) / np.abs(ref[1])))
        except Exception as e:
            api.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N)
            t = time.perf_counter()
            for _ in range(reps):
                vv = api.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N)
            ms_v = (time.perf_counter() - t) / reps * 1e3
            row = {"N": N, "slices": ns, "stages": stages, "leaf": leaf, "error": str(e)[:300]}
        print(json.dumps(row), flush=True)
    ctx.set_ozaki_ex(0, 1, 0)
    ctx.release_workspace()
