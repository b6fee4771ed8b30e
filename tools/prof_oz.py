"""Scratch: one opt-in Ozaki GEMM (INT8 tensor pipe) for ncu.  python tools/prof_oz.py [n] [slices]"""
import ctypes, sys
sys.path.insert(0, ".")
from sympgpr_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 7
ms = (ctypes.c_double * 2)()
_lib.check(_lib.lib().sgp_ozaki_bench(_lib.context(0).handle, ns, n, n, n, 1, ms), "ozaki_bench")
print(n, ns, ms[0], ms[1])
