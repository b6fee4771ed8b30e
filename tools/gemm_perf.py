"""Scratch: DMMA GEMM kernel timings on a GPU box.  python tools/gemm_perf.py [--one]"""
import ctypes, sys
sys.path.insert(0, ".")
from sympgpr_b200 import _lib
L = _lib.lib(); ctx = _lib.context()
MODES = {0: "full", 1: "lower", 2: "lower_kge", 3: "b_lower", 4: "a_lower"}


def flops(mode, Mt, Nt, K):
    T = 128
    if mode == 0: return 2.0 * Mt * T * Nt * T * K
    if mode == 1: return 2.0 * (Mt * (Mt + 1) / 2) * T * T * K
    if mode == 2: return sum(2.0 * (tm + 1) * T * T * (K - tm * T) for tm in range(Mt))
    if mode == 3: return sum(2.0 * Mt * T * T * (K - tn * T) for tn in range(Nt))
    if mode == 4: return sum(2.0 * Nt * T * T * min(K, (tm + 1) * T) for tm in range(Mt))


def run(al, bl, mode, Mt, Nt, K, reps=3):
    ms = ctypes.c_double(0.0)
    _lib.check(L.sgp_bench_gemm(ctx.handle, al, bl, mode, Mt, Nt, K, reps, ctypes.byref(ms)), "bench_gemm")
    f = flops(mode, Mt, Nt, K)
    print(f"al={al} bl={bl} mode={MODES[mode]:9s} Mt={Mt:4d} Nt={Nt:4d} K={K:6d}: {ms.value:9.3f} ms  {f / ms.value / 1e9:7.2f} TF", flush=True)


if "--one" in sys.argv:
    run(0, 0, 0, 74, 64, 8192, reps=1)      # 4736 tiles = 32 full waves of 148
    sys.exit(0)
# full square-ish GEMMs: waves of 148
for K in (128, 256, 512, 1024, 2048, 4096, 8192):
    run(0, 0, 0, 74, 64, K)
run(0, 0, 0, 64, 64, 8192)
run(0, 1, 0, 74, 64, 4096)
run(1, 1, 0, 74, 64, 4096)
# syrk (lower) as in the potrf trailing updates
for Mt, K in ((128, 128 * 128), (192, 8192), (64, 8192), (32, 4096), (16, 2048), (8, 1024)):
    run(0, 0, 1, Mt, Mt, K)
# lauum shape
run(1, 1, 2, 128, 128, 128 * 128, reps=1)
# trtri shapes
run(0, 1, 3, 64, 64, 64 * 128)
run(0, 1, 4, 64, 64, 64 * 128)
# trsm leaf / thin updates
run(0, 0, 0, 148, 1, 128, reps=10)
run(0, 0, 0, 148, 2, 256, reps=10)
run(0, 0, 0, 148, 4, 512, reps=10)
