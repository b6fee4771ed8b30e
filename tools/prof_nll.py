"""Scratch: one or more NLL+gradient evaluations for ncu launch lists.  python tools/prof_nll.py N [reps]"""
import sys
sys.path.insert(0, ".")
from sympgpr_b200 import api, workloads as W
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
import os
if os.environ.get("SGP_OZAKI"):                   # "ns:stages:leaf", e.g. 7:3:4096 = all three stages on the INT8 pipe
    from sympgpr_b200 import _lib
    ns, stages, leaf = (int(v) for v in os.environ["SGP_OZAKI"].split(":"))
    _lib.context().set_ozaki_ex(ns, stages, leaf)
d = W.standard_map_training(N); hyp = W.timing_hyp(N, d["sig"], 1e-8)
for _ in range(reps):
    v, g = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
print(N, v, g)
