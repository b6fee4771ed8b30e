#!/usr/bin/env python3
"""Build a variant of libsympgpr_b200.so with extra nvcc defines for kernel experiments on the GPU box:

    python tools/build_variant.py unr4 -DLL_UNR=4      ->  sympgpr_b200/_variants/libsympgpr_b200_unr4.so
    SYMPGPR_B200_LIB=sympgpr_b200/_variants/libsympgpr_b200_unr4.so python tools/ll_trace.py 4096

(*.so is git-ignored but travels with gpurun.)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sympgpr_b200 import build as B   # noqa: E402

tag, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "sympgpr_b200", "_variants")
obj_dir = os.path.join(out_dir, "obj_" + tag)
os.makedirs(obj_dir, exist_ok=True)
objs, procs = [], []
for src in B.SOURCES:
    o = os.path.join(obj_dir, src.replace(".cu", ".o"))
    objs.append(o)
    procs.append(subprocess.Popen([B.nvcc()] + B.NVCC_FLAGS + defs + ["-c", os.path.join(B.CSRC, src), "-o", o]))
if any(p.wait() for p in procs):
    sys.exit("nvcc failed")
lib = os.path.join(out_dir, f"libsympgpr_b200_{tag}.so")
subprocess.check_call([B.nvcc()] + B.ARCH + ["-shared", "-o", lib] + objs)
print(lib)
