"""Scratch: one map launch for ncu.  python tools/prof_map.py [Nt] [E] [steps] [solver]"""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from sympgpr_b200 import _lib, api, workloads as W
Nt = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
E = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
solver = int(sys.argv[4]) if len(sys.argv) > 4 else 1
L = _lib.lib(); ctx = _lib.context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
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
for rep in range(2):
    stats.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    _lib.check(L.sgp_model_applymap_dev(ctx.handle, model, 2, solver, steps, E, q0.data_ptr(), p0.data_ptr(), qf.data_ptr(),
                                        pf.data_ptr(), None, None, 0, stats.data_ptr()), "applymap_dev")
    e1.record(stream); e1.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    ev = int(stats[0].item())
    print(f"Nt={Nt} E={E} steps={steps} solver={solver}: {t*1e3:.2f} ms  {E*steps/t:.4e} orbit-steps/s  sweeps/orbit-step {1+ev/(E*steps):.3f} "
          f"pair-evals/s {(E*steps*Nt + ev*Nt)/t:.4e} unconverged {int(stats[1].item())}", flush=True)
