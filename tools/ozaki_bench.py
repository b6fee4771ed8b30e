"""Opt-in Ozaki GEMM on the INT8 tensor pipe vs the DMMA GEMM: python tools/ozaki_bench.py [n ...]"""
import ctypes, json, sys
sys.path.insert(0, ".")
from sympgpr_b200 import _lib
L = _lib.lib(); ctx = _lib.context(0)
for n in [int(x) for x in sys.argv[1:]] or [2048, 4096, 8192]:
    row = {"M=N=K": n}
    for ns in (6, 7):
        ms = (ctypes.c_double * 2)()
        _lib.check(L.sgp_ozaki_bench(ctx.handle, ns, n, n, n, 3, ms), "ozaki_bench")
        row[f"ozaki{ns}_total_ms"] = round(ms[0], 3); row[f"ozaki{ns}_gemm_ms"] = round(ms[1], 3)
        row[f"ozaki{ns}_TF_equiv_total"] = round(2.0 * n**3 / ms[0] / 1e9, 1); row[f"ozaki{ns}_TF_equiv_gemm"] = round(2.0 * n**3 / ms[1] / 1e9, 1)
    msd = ctypes.c_double(0.0)
    _lib.check(L.sgp_bench_gemm(ctx.handle, 0, 0, 0, n // 128, n // 128, n, 3, ctypes.byref(msd)), "bench_gemm")
    row["dmma_ms"] = round(msd.value, 3); row["dmma_TF"] = round(2.0 * n**3 / msd.value / 1e9, 1)
    print(json.dumps(row), flush=True)
