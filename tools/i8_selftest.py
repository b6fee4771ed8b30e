"""tcgen05 kind::i8 self test (csrc/ozaki.cu): python tools/i8_selftest.py [K ...]"""
import ctypes, sys
sys.path.insert(0, ".")
from sympgpr_b200 import _lib
L = _lib.lib(); ctx = _lib.context(0)
for K in [int(x) for x in sys.argv[1:]] or [128, 256, 1024]:
    bad, r, g = ctypes.c_int(-1), ctypes.c_int(0), ctypes.c_int(0)
    st = L.sgp_i8mma_selftest(ctx.handle, K, ctypes.byref(bad), ctypes.byref(r), ctypes.byref(g))
    print("K", K, "status", st, "mismatches", bad.value, "of 8192; probe ref", r.value, "got", g.value, _lib.last_error() if st else "", flush=True)
