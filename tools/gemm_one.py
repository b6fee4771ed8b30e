"""Scratch: a single DMMA GEMM launch for ncu.  python tools/gemm_one.py al bl mode Mt Nt K"""
import ctypes, sys
sys.path.insert(0, ".")
from sympgpr_b200 import _lib
L = _lib.lib(); ctx = _lib.context()
al, bl, mode, Mt, Nt, K = [int(v) for v in sys.argv[1:7]]
ms = ctypes.c_double(0.0)
_lib.check(L.sgp_bench_gemm(ctx.handle, al, bl, mode, Mt, Nt, K, 1, ctypes.byref(ms)), "bench_gemm")
print(al, bl, mode, Mt, Nt, K, ms.value, "ms")
