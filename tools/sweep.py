"""BASELINE config 5: NLL+gradient size sweep N = 2k..32k (n up to 65 536) and a 10^7-orbit map launch.
    python tools/sweep.py > gpurun_out/sweep.jsonl"""
import ctypes, json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from sympgpr_b200 import _lib, api, workloads as W

L = _lib.lib(); ctx = _lib.context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
_lib.check(L.sgp_set_profiling(ctx.handle, 1), "prof")
MAXN = int(os.environ.get("SWEEP_MAX_N", "32768"))
MINN = int(os.environ.get("SWEEP_MIN_N", "0"))
for N in [v for v in (1024, 2048, 4096, 8192, 16384, 32768) if MINN <= v <= MAXN]:
    n = 2 * N
    d = W.standard_map_training(N)
    hyp = W.timing_hyp(N, d["sig"], 1e-8)
    hyp_c = (ctypes.c_double * 4)(*hyp)
    x_d = torch.from_numpy(d["xtrain"].copy()).to(dev); z_d = torch.from_numpy(d["ztrain"].copy()).to(dev)
    res_d = torch.zeros(16, dtype=torch.float64, device=dev)
    def step(ng):
        _lib.check(L.sgp_nll_dev(ctx.handle, 0, 0.5, 0, hyp_c, x_d.data_ptr(), z_d.data_ptr(), n, ng, res_d.data_ptr()), "nll")
    out = {"N": N, "n": n}
    for name, ng in (("nll_grad", 2), ("nll", 0)):
        step(ng); torch.cuda.synchronize()
        reps = 3 if N <= 16384 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps): step(ng)
        e1.record(stream); e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = float(n)**3 * (1.0 if ng else 1.0 / 3.0)
        out[name + "_ms"] = ms; out[name + "_tflops"] = fl / ms / 1e9
        if ng:
            st = (ctypes.c_double * 7)(); _lib.check(L.sgp_stage_times(ctx.handle, st), "stages")
            out["stages_ms"] = dict(zip(["fill", "potrf", "potrs", "trtri", "lauum", "grad", "finalize"], [round(v, 3) for v in st]))
    r = res_d.cpu().numpy(); out["nll_value"] = float(r[0]); out["info"] = float(r[4])
    print(json.dumps(out), flush=True)
    ctx.release_workspace(); torch.cuda.empty_cache()

# 10^7 orbits, 16 steps, Nt = 4096
if os.environ.get("SWEEP_NO_MAP"):
    sys.exit(0)
Nt, E, steps = 4096, 10_000_000, 16
dm = W.standard_map_training(Nt)
hm = W.timing_hyp(Nt, dm["sig"], 1e-8, factor=1.0); hpm = W.timing_hyp(Nt, dm["sigp"], 1e-8, factor=1.0)
fm = api.fit(hm, dm["xtrain"], dm["ztrain"], 2 * Nt); fpm = api.fit(hpm, dm["xtrainp"], dm["ztrainp"], Nt, reg=True)
q0_all, p0_all = W.ensemble(E)
q0 = torch.from_numpy(q0_all).to(dev); p0 = torch.from_numpy(p0_all).to(dev)
qf, pf = torch.empty_like(q0), torch.empty_like(p0)
stats = torch.zeros(2, dtype=torch.int64, device=dev)
model = ctypes.c_void_p(); dp = _lib.dptr
xtp, xt = dm["xtrainp"], dm["xtrain"]
_lib.check(L.sgp_model_create(ctx.handle, 0, 0.5, dp(np.ascontiguousarray(hm[:3])), dp(np.ascontiguousarray(hpm[:3])),
                              dp(np.ascontiguousarray(xtp[:Nt])), dp(np.ascontiguousarray(xtp[Nt:])), dp(fpm["alpha"]), Nt,
                              dp(np.ascontiguousarray(xt[:Nt])), dp(np.ascontiguousarray(xt[Nt:])), dp(fm["alpha"]), Nt,
                              ctypes.byref(model)), "model_create")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
_lib.check(L.sgp_model_applymap_dev(ctx.handle, model, 2, 3, steps, E, q0.data_ptr(), p0.data_ptr(), qf.data_ptr(), pf.data_ptr(),
                                    None, None, 0, stats.data_ptr()), "applymap")
e1.record(stream); e1.synchronize()
t = e0.elapsed_time(e1) * 1e-3
print(json.dumps({"map_E": E, "steps": steps, "Nt": Nt, "seconds": t, "orbit_steps_per_s": E * steps / t,
                  "sweeps_per_orbit_step": 1 + int(stats[0].item()) / (E * steps), "unconverged": int(stats[1].item()),
                  "nan_final": int(torch.isnan(qf).sum().item())}), flush=True)
