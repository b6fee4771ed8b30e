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


# ---------------------------------------------------------------------------------------------------------------------
# the sliced products the INT8 factor-and-inverse recursion is made of (csrc/ozaki_chol.cu), one by one against NumPy
# ---------------------------------------------------------------------------------------------------------------------
def _lower_with_garbage(rng, n, scale_rows=False):
    """(valid lower-triangular matrix, the same with unrelated numbers above the diagonal): the Cholesky works in the lower
    triangle of a matrix whose upper triangle was never initialised."""
    X = np.tril(rng.standard_normal((n, n)))
    if scale_rows:
        X *= 10.0 ** rng.integers(-6, 6, size=(n, 1))
    G = X + np.triu(1e30 * rng.standard_normal((n, n)), 1)
    return X, np.asfortranarray(G)


def _bound(A, B, ns):
    """A-priori error bound of the sliced product, elementwise: the digit pairs with s + t >= ns are dropped, each at most
    128 x 128 x 2^(-8 (s + t)) x 2^(ea + eb - 14) per k with 2^ea <= 2.01 max_k |a|, plus the final FP64 roundings of the ns-term
    sum: it is relative to the ROW MAXIMA of the operands (the Ozaki scheme scales rows), not to sum |a||b|."""
    K = A.shape[1]
    dropped = sum((2 * ns - 1 - d) * 2.0 ** (-8 * d) for d in range(ns, 2 * ns - 1))
    return (4.1 * dropped * K) * np.outer(np.abs(A).max(axis=1), np.abs(B).max(axis=1)) + 5e-15 * (np.abs(A) @ np.abs(B).T)


def _check(C, A, B, sign=1.0, mask=None):
    """C against sign * A B^T (A, B = the valid parts of the operands) within _bound; returns the worst error / bound."""
    ns = _check.ns
    r = np.abs(C - sign * (A @ B.T)) / np.maximum(_bound(A, B, ns), 1e-300)
    if mask is not None:
        r = r[mask]
    assert r.max() < 1.0, r.max()
    return float(r.max())


@pytest.mark.parametrize("ns", [5, 6, 7, 8])
@pytest.mark.parametrize("h1,h2", [(128, 128), (384, 256), (256, 640)])
def test_sliced_products_of_the_recursion(api, ns, h1, h2):
    """Steps 2, 3, 4 and 6 of ozaki_chol.cu's node and the lauum product, with the k-ranges, storage orders and triangular
    validity flags the recursion passes, on operands whose unused triangle holds 1e30 and whose rows differ by 12 orders of
    magnitude: each against the FP64 product of the valid parts within the a-priori bound of the scheme."""
    rng = np.random.default_rng(100 * ns + h1 + h2)
    _check.ns = ns
    A21 = np.asfortranarray(rng.standard_normal((h2, h1)))
    X11, X11g = _lower_with_garbage(rng, h1, scale_rows=True)
    X22, X22g = _lower_with_garbage(rng, h2, scale_rows=True)
    # 2. L21 = A21 X11^T : B = X11 (rows r, k <= r valid), k < 64 (tn + 1)
    C = api.ozaki_gemm_ex(A21, X11g, h2, h1, h1, la=0, ta=0, lb=0, tb=1, kmode=8, slices=ns)
    w = [_check(C, A21, X11)]
    # 3. A22 - L21 L21^T on the lower tiles; tiles entirely above the diagonal keep their contents
    S0 = np.asfortranarray(rng.standard_normal((h2, h2)))
    C = api.ozaki_gemm_ex(A21, A21, h2, h2, h1, C=S0, alpha=-1.0, beta=1.0, lower=1, slices=ns)
    low = np.tril(np.ones((h2, h2), dtype=bool))
    w.append(_check(C - S0, A21, A21, sign=-1.0, mask=low))
    tm, tn = np.arange(h2)[:, None] // 128, np.arange(h2)[None, :] // 64
    untouched = tn > 2 * tm + 1
    assert np.array_equal(C[untouched], S0[untouched])
    # 4. L21 X11 : B = X11^T stored as X11 (element (n, k) at X11[k, n], k >= n valid), k >= 64 tn
    C = api.ozaki_gemm_ex(A21, X11g, h2, h1, h1, la=0, ta=0, lb=1, tb=2, kmode=2, slices=ns)
    w.append(_check(C, A21, X11.T))
    # 6. -X22 T2 : A = X22 (k <= r valid), k < 128 (tm + 1); B = T2^T stored as T2 (h2 x h1)
    T2 = np.asfortranarray(rng.standard_normal((h2, h1)))
    C = api.ozaki_gemm_ex(X22g, T2, h2, h1, h2, la=0, ta=1, lb=1, tb=0, alpha=-1.0, kmode=4, slices=ns)
    w.append(_check(C, X22, T2.T, sign=-1.0))
    # lauum: X^T X, A = B = X^T stored as X (k >= r valid), k >= 128 tm, lower tiles
    C = api.ozaki_gemm_ex(X22g, X22g, h2, h2, h2, la=1, ta=2, lb=1, tb=2, kmode=1, lower=1, slices=ns)
    w.append(_check(C, X22.T, X22.T, mask=low))
    print(f"\nns={ns} h1={h1} h2={h2}: worst error / bound per product {[f'{x:.2e}' for x in w]}")


@pytest.fixture()
def ozaki_all():
    """All three n^3/3 stages on the INT8 pipe for one test, and back to the DMMA default."""
    from sympgpr_b200 import _lib
    ctx = _lib.context()
    yield ctx
    ctx.set_ozaki_ex(0, 1, 0)


@pytest.mark.parametrize("ns,leaf", [(6, 128), (7, 256), (8, 512)])
@pytest.mark.parametrize("n", [700, 1536])
def test_spd_inverse_and_logdet_on_the_int8_route(api, ozaki_all, n, ns, leaf):
    """sgp_spd_factor with factor + triangular inverse + lauum all on the INT8 pipe (leaves of `leaf` rows on DMMA): inverse
    and log-determinant of a random SPD matrix (condition ~1e6) against SciPy."""
    import scipy.linalg
    rng = np.random.default_rng(n + ns + leaf)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (Q * np.logspace(0, -6, n)) @ Q.T
    A = 0.5 * (A + A.T)
    _, Ai0, ld0 = api.spd_factor(A, want_factor=False, want_inverse=True)
    ozaki_all.set_ozaki_ex(ns, 3, leaf)
    _, Ai, ld = api.spd_factor(A, want_factor=False, want_inverse=True)
    ref = scipy.linalg.inv(A)
    ldr = 0.5 * np.linalg.slogdet(A)[1]
    assert abs(ld - ldr) < 1e-10 * max(1.0, abs(ldr)) and abs(ld0 - ldr) < 1e-10 * max(1.0, abs(ldr))
    e = np.abs(Ai - ref).max() / np.abs(ref).max()
    e0 = np.abs(Ai0 - ref).max() / np.abs(ref).max()
    r = np.abs(A @ Ai - np.eye(n)).max()             # the residual does not depend on the comparison inverse
    r0 = np.abs(A @ Ai0 - np.eye(n)).max()
    print(f"\nn={n} ns={ns} leaf={leaf}: max|Ainv - ref| / max|ref| = {e:.2e} (DMMA route {e0:.2e}); |A Ainv - I| = {r:.2e} ({r0:.2e})")
    # 6 digits keep 47 bits of every row's scale against the 53 of FP64; 7 and 8 keep 55 and 63 and must match the DMMA route
    slack = 3000.0 if ns == 6 else 8.0
    assert e < slack * e0, (e, e0)
    assert r < slack * r0, (r, r0)


@pytest.mark.parametrize("ns,leaf", [(6, 256), (7, 256), (8, 512)])
@pytest.mark.parametrize("N", [600, 1024])
def test_nll_gradient_with_all_stages_on_the_int8_pipe(api, ozaki_all, N, ns, leaf):
    """NLL + gradient with potrf, trtri and lauum replaced by the INT8 recursion + INT8 lauum, judged on the SAME oracle
    values and tolerance (1e-9) as the DMMA path."""
    from oracle import oracle as O
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    vr, grr = O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    ozaki_all.set_ozaki_ex(ns, 3, leaf)
    v, gr = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    assert np.isclose(v, vr, rtol=1e-9), (v, vr)
    assert np.allclose(gr, grr, rtol=1e-9, atol=1e-9 * np.abs(grr).max()), (gr, grr)
    if N == 600:
        # model finalisation (sgp_fit: alpha and Kyinv) on the same route against the DMMA route
        f = api.fit(hyp, d["xtrain"], d["ztrain"], 2 * N, want_inverse=True)
        ozaki_all.set_ozaki_ex(0, 1, 0)
        f0 = api.fit(hyp, d["xtrain"], d["ztrain"], 2 * N, want_inverse=True)
        tol = 1e-6 if ns == 6 else 1e-9
        assert np.allclose(f["alpha"], f0["alpha"], rtol=tol, atol=tol * np.abs(f0["alpha"]).max())
        assert np.allclose(f["Kyinv"], f0["Kyinv"], rtol=tol, atol=tol * np.abs(f0["Kyinv"]).max())


@pytest.mark.parametrize("ns,leaf", [(6, 128), (7, 256), (8, 384)])
@pytest.mark.parametrize("N", [600, 1024, 1600])
def test_nll_value_on_the_factor_only_int8_recursion(api, ozaki_all, N, ns, leaf):
    """nll_chol (value only: the objective of the scripts' L-BFGS loops) through the factor-only variant of the INT8 recursion --
    inverse factors on the diagonal blocks, L below them, forward substitution riding along -- against the oracle at 1e-9, and
    the regression-kernel variant (nll_chol_reg) likewise."""
    from oracle import oracle as O
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    vr = O.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N)
    v0 = api.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N)
    ozaki_all.set_ozaki_ex(ns, 3, leaf)
    v = api.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N)
    print(f"\nN={N} ns={ns} leaf={leaf}: nll {v!r} (DMMA {v0!r}, oracle {vr!r})")
    assert np.isclose(v, vr, rtol=1e-9), (v, vr)
    assert np.isclose(v, v0, rtol=1e-11), (v, v0)
    hypp = O.timing_hyp(N, d["sigp"], 1e-8)
    vr = O.nll_chol_reg(hypp, d["xtrainp"], d["ztrainp"], N)
    v = api.nll_chol_reg(hypp, d["xtrainp"], d["ztrainp"], N)
    assert np.isclose(v, vr, rtol=1e-9), (v, vr)


def test_int8_route_reports_a_matrix_that_is_not_positive_definite(api, ozaki_all):
    """The leaf factorisations carry the info word: a negative pivot in the SECOND half of the recursion (after sliced updates)
    still raises LinAlgError, as the reference's bare `except` around scipy.linalg.cholesky expects."""
    n = 1024
    A = np.eye(n) * 2.0
    A[900, 900] = -1.0
    ozaki_all.set_ozaki_ex(7, 3, 256)
    with pytest.raises(np.linalg.LinAlgError):
        api.spd_factor(A, want_factor=False, want_inverse=True)


@pytest.mark.parametrize("ns", [6, 7])
def test_full_size_gradient_with_all_stages_on_the_int8_pipe(api, ozaki_all, ns):
    """BASELINE's headline size (N = 16 384, n = 32 768): all three stages on the INT8 pipe (6 digits = what bench.py times, and 7;
    leaves of 4096) against the CPU golden (tests/golden/fullsize_nll_N16384.json), 1e-9 as for the DMMA path; the value-only
    evaluation (factor-only recursion) likewise."""
    import json
    import os
    from oracle import oracle as O
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fullsize_nll_N16384.json")))
    N = g["N"]
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    ozaki_all.set_ozaki_ex(ns, 3, 4096)
    v, gr = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    v1 = api.nll_chol(hyp, d["xtrain"], d["ztrain"], 2 * N)
    print(f"\nfull size, INT8 route, {ns} digits: nll rel err {abs(v - g['nll']) / abs(g['nll']):.2e} (value-only {abs(v1 - g['nll']) / abs(g['nll']):.2e}), "
          f"grad rel err {np.max(np.abs(np.asarray(gr) - np.asarray(g['grad'])) / np.abs(g['grad'])):.2e}")
    assert abs(v - g["nll"]) <= 1e-9 * abs(g["nll"])
    assert abs(v1 - g["nll"]) <= 1e-9 * abs(g["nll"])
    assert np.allclose(gr, g["grad"], rtol=1e-9, atol=1e-9 * np.abs(g["grad"]).max()), (gr, g["grad"])


def test_ill_conditioned_map_model_on_the_int8_route(api, ozaki_all):
    """The map models of scripts 03/04/05 are fitted with length scales twice those of the timing workload: cond(K) ~ 1e10, where
    FP64 itself only holds ~1e-7 of the gradient.  8 digits (63 bits of every row's scale) must stay within a small factor of
    the DMMA route's own distance to the oracle; the distances of 6 and 7 digits are printed for the record."""
    from oracle import oracle as O
    N = 768
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    hyp[:2] *= 2.0
    vr, gr = O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    v0, g0 = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    dist = lambda v, g: (abs(v - vr) / abs(vr), float(np.max(np.abs(np.asarray(g) - gr) / np.abs(gr))))
    e0 = dist(v0, g0)
    rows = {}
    for ns in (6, 7, 8):
        ozaki_all.set_ozaki_ex(ns, 3, 256)
        v, g = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
        rows[ns] = dist(v, g)
    print(f"\ncond 1e10 model, N={N}: distance to the oracle (nll, grad): DMMA {e0[0]:.1e} {e0[1]:.1e}; "
          + "; ".join(f"{ns} digits {e[0]:.1e} {e[1]:.1e}" for ns, e in rows.items()))
    assert rows[8][0] <= max(5.0 * e0[0], 1e-9) and rows[8][1] <= max(5.0 * e0[1], 1e-9), (rows, e0)
    assert rows[7][0] <= max(50.0 * e0[0], 1e-9) and rows[7][1] <= max(50.0 * e0[1], 1e-9), (rows, e0)


def test_env_switch_selects_the_int8_route_for_unchanged_scripts(api):
    """SYMPGPR_B200_OZAKI=digits:stages:leaf in the environment of a process that never mentions the route (what a reference script
    run through sympgpr_b200.runner is): the default context takes the INT8 route -- the potrf stage disappears from the stage
    times (factor and inverse run as one recursion) -- and the results agree with the DMMA route of this process."""
    import json
    import os
    import subprocess
    import sys
    from oracle import oracle as O
    N = 600
    code = (
        "import ctypes, json, sys\n"
        "sys.path.insert(0, '.')\n"
        "from oracle import oracle as O\n"
        "from sympgpr_b200 import _lib, api\n"
        f"d = O.standard_map_training({N}); hyp = O.timing_hyp({N}, d['sig'], 1e-8)\n"
        "ctx = _lib.context(); _lib.check(_lib.lib().sgp_set_profiling(ctx.handle, 1), 'prof')\n"
        f"v, g = api.nll_grad(hyp, d['xtrain'], d['ztrain'], {2 * N})\n"
        "st = (ctypes.c_double * 7)(); _lib.check(_lib.lib().sgp_stage_times(ctx.handle, st), 'st')\n"
        "print(json.dumps({'v': v, 'g': [float(g[0]), float(g[1])], 'potrf_ms': st[1], 'trtri_ms': st[3]}))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SYMPGPR_B200_OZAKI="7:3:256")
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    v0, g0 = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    assert r["potrf_ms"] < 0.02 < r["trtri_ms"], r                      # potrf + trtri ran as one recursion, timed under trtri
    assert np.isclose(r["v"], v0, rtol=1e-11) and np.allclose(r["g"], g0, rtol=1e-9)


def test_other_kernel_variants_on_the_int8_route(api, ozaki_all):
    """The regression-kernel gradient (nll_grad_reg) and the 2-DOF 4 x 4-block kernel (reg = 4, BASELINE config 3) take the same
    recursion: against the oracle at the tolerances of their DMMA tests."""
    from oracle import oracle as O
    N = 1024
    d = O.standard_map_training(N)
    hypp = O.timing_hyp(N, d["sigp"], 1e-8)
    vr, gr = O.nll_grad_reg(hypp, d["xtrainp"], d["ztrainp"], N)
    ozaki_all.set_ozaki_ex(7, 3, 256)
    v, g = api.nll_grad_reg(hypp, d["xtrainp"], d["ztrainp"], N)
    assert np.isclose(v, vr, rtol=1e-9), (v, vr)
    assert np.allclose(g, gr, rtol=1e-9, atol=1e-9 * np.abs(gr).max()), (g, gr)
    N2 = 200
    x, z = O.henon_like_training(N2)
    hyp = np.array([0.35, 0.4, 2 * np.max(np.abs(z))**2, 1e-4])          # cond(Ky) ~ 4e5
    vr, gr = O.nll_grad4(hyp, x, z, 4 * N2, with_sig=True)
    v, g = api.nll_grad4(hyp, x, z, 4 * N2, with_sig=True)
    assert np.isclose(v, vr, rtol=1e-9), (v, vr)
    assert np.allclose(g, gr, rtol=1e-8, atol=1e-8 * np.abs(gr).max()), (g, gr)


def test_int8_route_on_a_diagonally_scaled_matrix(api, ozaki_all):
    """A = D B D with B well conditioned and d spanning 16 orders of magnitude.  The digits of the INT8 route are fixed-point numbers
    relative to the maximum of each operand ROW: a spread of scales ALONG k (the columns of A21, the columns of X = L^-1) would eat
    them, so sgp_spd_factor equilibrates an arbitrary matrix by powers of two first (exact, undone exactly in the outputs):
    D Ainv D = B^-1 and the log-determinant as accurately as on the scale-invariant DMMA route."""
    import scipy.linalg
    rng = np.random.default_rng(11)
    n = 1280
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    B = (Q * np.linspace(1.0, 50.0, n)) @ Q.T
    B = 0.5 * (B + B.T)
    Bi = scipy.linalg.inv(B)
    dscale = 10.0 ** rng.uniform(-8, 8, n)
    A = B * dscale[:, None] * dscale[None, :]
    A = 0.5 * (A + A.T)
    ldr = 0.5 * np.linalg.slogdet(B)[1] + np.sum(np.log(dscale))
    out = {}
    for name, ns in (("dmma", 0), ("int8", 7)):
        ozaki_all.set_ozaki_ex(ns, 3 if ns else 1, 256)
        _, Ai, ld = api.spd_factor(A, want_factor=False, want_inverse=True)
        out[name] = (np.abs(Ai * dscale[:, None] * dscale[None, :] - Bi).max() / np.abs(Bi).max(), abs(ld - ldr) / abs(ldr))
    print(f"\nD B D, 16 decades: DMMA {out['dmma'][0]:.1e} (logdet {out['dmma'][1]:.1e}), INT8 {out['int8'][0]:.1e} ({out['int8'][1]:.1e})")
    assert out["dmma"][0] < 1e-11 and out["dmma"][1] < 1e-12, out
    assert out["int8"][0] < 1e-11 and out["int8"][1] < 1e-12, out
    # the factor comes back unscaled as well (lauum-only mode keeps L): L L^T = A
    ozaki_all.set_ozaki_ex(7, 1, 256)
    L, Ai, _ = api.spd_factor(A, want_factor=True, want_inverse=True)
    assert np.abs((L @ L.T) / A - 1.0).max() < 1e-9
    assert np.abs(Ai * dscale[:, None] * dscale[None, :] - Bi).max() / np.abs(Bi).max() < 1e-11
