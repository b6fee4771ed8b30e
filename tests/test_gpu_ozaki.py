"""GPU tests of the OPT-IN FP64 GEMM on the INT8 tensor pipe (tcgen05.mma kind::i8, Ozaki splitting; csrc/ozaki.cu):
the tcgen05 plumbing against a plain integer kernel, the sliced GEMM against NumPy and against the library's DMMA GEMM."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from sympgpr_b200 import _lib, api as a
    if _lib.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests need a B200 (the product has no CPU fallback)")
    return a


@pytest.mark.parametrize("K", [128, 384, 4096])
def test_int8_umma_matches_plain_integer_kernel(api, K):
    """One 128 x 64 x K INT8 product through tcgen05.mma kind::i8 (TMEM accumulators, K-major 128-byte-swizzled shared-memory
    descriptors, tcgen05.commit -> mbarrier, tcgen05.ld) against a thread-per-element integer kernel: bit exact."""
    from sympgpr_b200 import _lib
    bad, r, g = ctypes.c_int(-1), ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.lib().sgp_i8mma_selftest(_lib.context().handle, K, ctypes.byref(bad), ctypes.byref(r), ctypes.byref(g)), "i8mma_selftest")
    assert bad.value == 0, (bad.value, r.value, g.value)
    assert r.value == g.value


def _rel_err(C, A, B):
    ref = A @ B.T
    scale = np.abs(A) @ np.abs(B).T
    return float(np.max(np.abs(C - ref) / np.maximum(scale, 1e-300)))


@pytest.mark.parametrize("M,N,K", [(128, 64, 128), (200, 100, 300), (384, 256, 1024), (1000, 520, 2048)])
@pytest.mark.parametrize("ns", [7, 8])
def test_ozaki_gemm_matches_numpy(api, M, N, K, ns):
    """C = A B^T from signed 7-bit slices on the INT8 tensor pipe against NumPy's FP64 product, error measured against
    sum_k |a_mk||b_nk| per element: 8 slices reproduce FP64 (1e-14); 7 slices keep 49 bits of each row's scale."""
    rng = np.random.default_rng(M + N + K + ns)
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((N, K))
    C = api.ozaki_gemm(A, B, slices=ns)
    e = _rel_err(C, A, B)
    print(f"\nozaki {M}x{N}x{K}, {ns} slices: max |C - AB^T| / (|A||B|^T) = {e:.2e}")
    assert e < (1e-14 if ns == 8 else 2e-13), e


def test_ozaki_gemm_alpha_beta_and_row_scales(api):
    """alpha / beta, and operands whose rows differ by 30 orders of magnitude (per-row exponents: the error stays relative to
    each row's own scale); a zero row and a zero matrix."""
    rng = np.random.default_rng(5)
    M, N, K = 256, 128, 512
    A = rng.standard_normal((M, K)) * 10.0 ** rng.integers(-15, 15, size=(M, 1))
    B = rng.standard_normal((N, K)) * 10.0 ** rng.integers(-15, 15, size=(N, 1))
    A[7] = 0.0
    C0 = rng.standard_normal((M, N))
    C = api.ozaki_gemm(A, B, C=C0, alpha=-1.25, beta=0.75, slices=8)
    ref = 0.75 * C0 - 1.25 * (A @ B.T)
    scale = 0.75 * np.abs(C0) + 1.25 * (np.abs(A) @ np.abs(B).T)
    assert np.max(np.abs(C - ref) / scale) < 1e-14
    assert np.array_equal(C[7], 0.75 * C0[7])
    Z = api.ozaki_gemm(np.zeros((130, 140)), B[:, :140], slices=7)
    assert Z.shape == (130, 128) and not Z.any()


def test_ozaki_gemm_against_the_dmma_gemm(api):
    """Same operands through the library's DMMA kernel (gemm_f64_ws_kernel) and through the INT8 path: both within 1e-14 of
    sum |a||b| of NumPy, and of each other."""
    rng = np.random.default_rng(9)
    M, N, K = 512, 256, 1024
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((N, K))
    Cd = np.zeros((M, N), order="F")
    api.gemm_host(0, 0, 0, A, B, Cd, 1.0, 0.0)
    Co = api.ozaki_gemm(A, B, slices=8)
    scale = np.abs(A) @ np.abs(B).T
    assert np.max(np.abs(Cd - A @ B.T) / scale) < 1e-14
    assert np.max(np.abs(Co - Cd) / scale) < 1e-14


@pytest.fixture()
def ozaki_ctx():
    """Switch the default context's lauum stage to the INT8 pipe for one test, and back."""
    from sympgpr_b200 import _lib
    ctx = _lib.context()
    yield ctx
    ctx.set_ozaki(0)


@pytest.mark.parametrize("ns", [7, 8])
@pytest.mark.parametrize("N", [100, 1024])
def test_nll_gradient_with_lauum_on_the_int8_pipe(api, ozaki_ctx, N, ns):
    """Opt-in route past the DMMA ceiling on the real path: K^-1 = X^T X (a third of an NLL+gradient evaluation) from INT8
    slice products, judged on the SAME oracle values and tolerance (1e-9) as the DMMA path; the inverse itself against SciPy."""
    import scipy.linalg
    from oracle import oracle as O
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    vr, grr = O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    v0, g0 = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    ozaki_ctx.set_ozaki(ns)
    v, gr = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    assert np.isclose(v, vr, rtol=1e-9) and v == v0                      # the value does not depend on the inverse
    assert np.allclose(gr, grr, rtol=1e-9, atol=1e-9 * np.abs(grr).max()), (gr, grr, g0)
    if N <= 200:
        f = api.fit(hyp, d["xtrain"], d["ztrain"], 2 * N, want_inverse=True)
        xt = d["xtrain"]
        K = O.build_k_vec(xt[:N], xt[N:], xt[:N], xt[N:], hyp[:3]) + hyp[3] * np.eye(2 * N)
        Kir = scipy.linalg.inv(K)
        assert np.allclose(f["Kyinv"], Kir, rtol=1e-8, atol=1e-9 * np.abs(Kir).max())


def test_full_size_gradient_with_lauum_on_the_int8_pipe(api, ozaki_ctx):
    """BASELINE's headline size (N = 16 384, n = 32 768) with the lauum stage on the INT8 pipe (8 slices) against the CPU
    golden (tests/golden/fullsize_nll_N16384.json), 1e-9 as for the DMMA path."""
    import json
    import os
    from oracle import oracle as O
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fullsize_nll_N16384.json")))
    N = g["N"]
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    ozaki_ctx.set_ozaki(8)
    v, gr = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    assert abs(v - g["nll"]) <= 1e-9 * abs(g["nll"])
    assert np.allclose(gr, g["grad"], rtol=1e-9, atol=1e-9 * np.abs(g["grad"]).max()), (gr, g["grad"])
