"""Scratch: panel-width scan of the blocked potrf / trtri on a GPU box."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")  # run from the repo root
from sympgpr_b200 import _lib, api, workloads as W
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = W.standard_map_training(N); hyp = W.timing_hyp(N, d["sig"], 1e-8); n = 2 * N
api.nll_chol(hyp, d["xtrain"], d["ztrain"], n)
for nb in (2, 4, 8, 16):
    os.environ["SGP_POTRF_NB"] = str(nb)
    ts = []
    for _ in range(2):
        t = time.time(); v = api.nll_chol(hyp, d["xtrain"], d["ztrain"], n); ts.append(time.time() - t)
    print(f"potrf NB={nb*128}: nll {min(ts)*1e3:.1f} ms ({n**3/3/min(ts)/1e12:.2f} TF) val={v:.12g}", flush=True)
os.environ["SGP_POTRF_NB"] = "4"
for nb in (2, 4, 8, 16):
    os.environ["SGP_TRTRI_NB"] = str(nb)
    ts = []
    for _ in range(2):
        t = time.time(); v, g = api.nll_grad(hyp, d["xtrain"], d["ztrain"], n); ts.append(time.time() - t)
    print(f"trtri NB={nb*128}: nll+grad {min(ts)*1e3:.1f} ms ({n**3/min(ts)/1e12:.2f} TF) val={v:.12g} g={g}", flush=True)
