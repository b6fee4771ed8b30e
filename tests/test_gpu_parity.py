"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the oracle
on the same inputs and against the committed golden fixtures.

Tolerances are the ones BASELINE.json states: kernel entries 1e-12 (allclose-style, as
python/05_tokamak/SympGPR/test_sympgpr.py:26-74), NLL and gradient 1e-9 relative, orbits 1e-8."""
import os

import numpy as np
import pytest
import scipy.linalg

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def api():
    from sympgpr_b200 import _lib, api as a
    if _lib.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests need a B200 (the product has no CPU fallback)")
    return a


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def C():
    from oracle import c_oracle
    return c_oracle


# ------------------------------------------------------------------------------- dense blocks
LAY = {"MN": 0, "K": 1}


@pytest.mark.parametrize("al,bl,mode,Mt,Nt,K", [
    (0, 0, 0, 1, 1, 16), (0, 0, 0, 2, 3, 128), (0, 0, 0, 3, 1, 384), (0, 0, 1, 3, 3, 256),
    (0, 1, 0, 2, 2, 128), (0, 1, 3, 3, 2, 256), (0, 1, 4, 3, 2, 384), (0, 1, 4, 1, 1, 128),
    (1, 1, 0, 2, 2, 64), (1, 1, 2, 3, 3, 384), (1, 1, 2, 1, 1, 128), (1, 0, 0, 2, 1, 48),
    (0, 0, 0, 9, 5, 2048),
])
def test_dmma_gemm_matches_plain_fp64(api, al, bl, mode, Mt, Nt, K):
    err = api.selftest_gemm(al, bl, mode, Mt, Nt, K)
    assert err < 1e-11 * max(1, K / 64), err


@pytest.mark.parametrize("al,bl,mode,Mt,Nt,K", [
    (0, 0, 0, 2, 3, 128), (0, 0, 1, 3, 3, 256), (0, 1, 3, 3, 2, 256), (0, 1, 4, 3, 2, 384), (1, 1, 2, 3, 3, 384),
    (1, 0, 0, 2, 1, 48), (0, 0, 0, 5, 4, 1024),
])
def test_dmma_gemm_matches_numpy(api, al, bl, mode, Mt, Nt, K):
    """The DMMA kernel on host operands against NumPy (not against a sibling kernel of the library): every operand
    layout and tile mode (full / lower / the three triangular k-ranges) the Cholesky drivers use."""
    rng = np.random.default_rng(1000 * al + 100 * bl + 10 * mode + Mt)
    M, N = 128 * Mt, 128 * Nt
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K))
    C0 = np.asfortranarray(rng.standard_normal((M, N)))
    alpha, beta = -1.25, 0.75
    C = C0.copy(order="F")
    api.gemm_host(al, bl, mode, A if al == 0 else A.T, B if bl == 0 else B.T, C, alpha, beta)
    ref = C0.copy()
    for tm in range(Mt):
        for tn in range(Nt):
            if mode in (1, 2) and tn > tm:
                continue                                          # lower modes leave the upper tiles alone
            k0, k1 = 0, K
            if mode == 2:
                k0 = tm * 128
            if mode == 3:
                k0 = tn * 128
            if mode == 4:
                k1 = min((tm + 1) * 128, K)
            r, c = slice(tm * 128, (tm + 1) * 128), slice(tn * 128, (tn + 1) * 128)
            ref[r, c] = beta * C0[r, c] + alpha * A[r, k0:k1] @ B[c, k0:k1].T
    assert np.allclose(C, ref, rtol=0, atol=1e-12 * max(1, K / 16)), np.abs(C - ref).max()


@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 300, 640, 1500])
def test_spd_factor_and_inverse(api, n):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n + 3))
    A = B @ B.T / n + 0.5 * np.eye(n)
    L, Ai, ld = api.spd_factor(A, want_factor=True, want_inverse=True)
    Lr = scipy.linalg.cholesky(A, lower=True)
    assert np.allclose(L, Lr, rtol=1e-11, atol=1e-12)
    assert np.isclose(ld, np.sum(np.log(np.diag(Lr))), rtol=1e-12, atol=1e-12)
    Air = np.linalg.inv(A)
    assert np.allclose(Ai, Air, rtol=1e-9, atol=1e-10 * np.abs(Air).max())
    assert np.allclose(Ai, Ai.T, rtol=0, atol=0)


def test_not_positive_definite_raises_linalgerror(api):
    A = np.eye(200)
    A[150, 150] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        api.spd_factor(A)
    # nll with a negative signal variance is not PD either; the reference's bare except relies on the raise
    x = np.linspace(0, 3, 8)
    with pytest.raises(np.linalg.LinAlgError):
        api.nll_chol([1.0, 1.0, -1.0, 1e-10], x, np.ones(8), 8)


# ------------------------------------------------------------------------------- fills
LIT = dict(x=np.array([1.0, 2.0, 3.0]), y=np.array([0.0, 3.0, 2.0]), x0=np.array([1.0, 2.0]), y0=np.array([0.0, 3.0]),
           hyp=np.array([0.5, 2.0, 0.4]), hypp=np.array([0.6, 1.9, 0.3]))


@pytest.mark.parametrize("family", ["product", "sq"])
def test_literal_inputs_of_test_sympgpr(api, family):
    """python/05_tokamak/SympGPR/test_sympgpr.py:17-74 against the golden values."""
    g = np.load(os.path.join(G, f"path_{family}.npz"))
    x, y, x0, y0, hyp, hypp = (LIT[k] for k in ("x", "y", "x0", "y0", "hyp", "hypp"))
    K = np.empty((3, 2), order="F")
    api.buildkreg(x, y, x0, y0, hyp, K, family)
    assert np.allclose(K, g["lit_buildkreg"], rtol=1e-12, atol=1e-12)
    K = np.zeros((1, 2), order="F")
    api.buildkreg(x[:1], y[:1], x0, y0, hyp, K, family)
    assert np.allclose(K, g["lit_buildkreg_1"], rtol=1e-12, atol=1e-12)
    K = np.empty((6, 4), order="F")
    api.build_k(x, y, x0, y0, hyp, K, family)
    assert np.allclose(K, g["lit_build_k"], rtol=1e-12, atol=1e-12)
    # same through the func.py-level wrappers
    K2 = np.empty((6, 4), order="F")
    api.build_K(np.hstack((x, y)), np.hstack((x0, y0)), hyp, K2, family)
    assert np.array_equal(K, K2)
    Kyinvp = np.array([[0.9, -0.3], [0.3, 0.9]], order="F")
    ztp = np.cos(x0 + y0)
    assert np.allclose(api.guessp(x[0], y[0], hypp, x0, y0, ztp, Kyinvp, family), g["lit_guessp"], rtol=1e-12, atol=1e-12)
    assert np.allclose(api.guessP(x[0], y[0], hypp, np.hstack((x0, y0)), ztp, Kyinvp, family), g["lit_guessp"],
                       rtol=1e-12, atol=1e-12)
    Kyinv = np.reshape(np.arange(16), (4, 4), order="F")          # int64, as in the reference test
    zt = np.hstack((np.cos(x0 + y0), np.sin(x0 + y0)))
    assert np.allclose(api.calcq(x[0], y[0], x0, y0, hyp, Kyinv, zt, family), g["lit_calcq"], rtol=1e-12, atol=1e-12)
    for solver in ("hybrd", "newton"):
        p = api.calcp(x[0], y[0], hyp, hypp, x0, y0, ztp, Kyinvp, x0, y0, zt, Kyinv, family, solver=solver)
        assert np.allclose(p, g["lit_calcp"], rtol=1e-12, atol=1e-12), (solver, p)
    p = api.calcP(x[0], y[0], hyp, hypp, np.hstack((x0, y0)), ztp, Kyinvp, np.hstack((x0, y0)), zt, Kyinv, family)
    assert np.allclose(p, g["lit_calcp"], rtol=1e-12, atol=1e-12)
    if family == "product":
        assert abs(p - 1.08172922) < 5e-9          # sympgpr.f90:106 "! pgss = 1.08172922d0"


@pytest.mark.parametrize("family,per", [("product", 0.5), ("sq", 0.5), ("sum", 0.5), ("period", 0.8)])
@pytest.mark.parametrize("N,N0", [(1, 1), (2, 7), (33, 65), (257, 130), (301, 299)])
def test_fill_matches_oracle(api, O, family, per, N, N0):
    rng = np.random.default_rng(N * 1000 + N0)
    x, y = rng.uniform(0, 6.3, N), rng.uniform(-2, 2, N)
    x0, y0 = rng.uniform(0, 6.3, N0), rng.uniform(-2, 2, N0)
    hyp = np.array([0.7, 1.3, 2.5])
    K = np.full((2 * N, 2 * N0), np.nan, order="F")
    api.build_k(x, y, x0, y0, hyp, K, family, per)
    Kr = O.build_k_vec(x, y, x0, y0, hyp, family, per)
    assert np.allclose(K, Kr, rtol=1e-12, atol=1e-12 * hyp[2])
    Kg = np.full((N, N0), np.nan, order="F")
    api.buildkreg(x, y, x0, y0, hyp, Kg, family, per)
    assert np.allclose(Kg, O.buildkreg_vec(x, y, x0, y0, hyp, family, per), rtol=1e-12, atol=1e-12)


def test_fill_edge_cases(api, O):
    x = np.array([0.3, 1.1]); y = np.array([0.2, -0.4]); hyp = np.array([0.8, 0.9, 2.0])
    # odd dimensions: sympgpr.f90:21-22,37 -- last row/col untouched by the loop, then scaled
    K = np.full((5, 4), 7.0, order="F")
    api.build_k(x, y, x, y, hyp, K)
    Kr = np.full((5, 4), 7.0, order="F")
    O.build_k(x, y, x, y, hyp, Kr)
    assert np.allclose(K, Kr, rtol=1e-12, atol=1e-12)
    # C-ordered square matrix (python/functions/func.py:133,149 allocate it that way)
    Kc = np.empty((4, 4))
    api.build_k(x, y, x, y, hyp, Kc)
    assert np.allclose(Kc, O.build_k_vec(x, y, x, y, hyp), rtol=1e-12, atol=1e-12)
    # empty
    api.build_k(np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), hyp, np.zeros((0, 0), order="F"))
    # int input arrays are converted like f2py intent(in)
    Ki = np.empty((4, 4), order="F")
    api.build_k(np.array([0, 1]), np.array([1, 2]), np.array([0, 1]), np.array([1, 2]), hyp, Ki)
    assert np.allclose(Ki, O.build_k_vec([0.0, 1.0], [1.0, 2.0], [0.0, 1.0], [1.0, 2.0], hyp), rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------------------- NLL / gradient
@pytest.mark.parametrize("family", ["product", "sq"])
def test_nll_and_gradient_match_reference_python_layer(api, family):
    g = np.load(os.path.join(G, f"path_{family}.npz"))
    N = int(g["N"][0])
    xt, zt, xtp, ztp = g["xtrain"], g["ztrain"], g["xtrainp"], g["ztrainp"]
    for k, h in enumerate(g["hyps"]):
        assert np.isclose(api.nll_chol(h, xt, zt, 2 * N, family), g["nll_chol"][k], rtol=1e-9)
        v, gr = api.nll_grad(h, xt, zt, 2 * N, family)
        assert np.isclose(v, g["nll_grad_val"][k], rtol=1e-9)
        assert np.allclose(gr, g["nll_grad_grad"][k], rtol=1e-6, atol=1e-6)     # reference = LU + full GEMM traces
    for k, h in enumerate(g["hypps"]):
        assert np.isclose(api.nll_chol_reg(h, xtp, ztp, N, family), g["nll_chol_reg"][k], rtol=1e-9)
        v, gr = api.nll_grad_reg(h, xtp, ztp, N, family)
        assert np.isclose(v, g["nll_grad_reg_val"][k], rtol=1e-9)
        assert np.allclose(gr, g["nll_grad_reg_grad"][k], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("family,per", [("product", 0.5), ("sq", 0.5), ("sum", 0.5), ("period", 0.7)])
@pytest.mark.parametrize("N", [3, 64, 100, 200, 333])
def test_nll_grad_matches_oracle(api, O, family, per, N):
    """N = 200 is the size of BASELINE config 01_pendulum (n = 400)."""
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    if family == "sum":
        hyp[3] = 0.1           # block-diagonal 1-D blocks: cond(Ky) ~ 3e5 with this noise (3e10 at 1e-6)
    xt, zt = d["xtrain"], d["ztrain"]
    v, gr = api.nll_grad(hyp, xt, zt, 2 * N, family, per, with_sig=True)
    vr, grr = O.nll_grad(hyp, xt, zt, 2 * N, family, per, with_sig=True)
    assert np.isclose(v, vr, rtol=1e-9), (v, vr)
    assert np.allclose(gr, grr, rtol=1e-9, atol=1e-9 * np.abs(grr).max()), (gr, grr)
    assert np.isclose(api.nll_chol(hyp, xt, zt, 2 * N, family, per), vr, rtol=1e-9)
    hypp = O.timing_hyp(N, d["sigp"], 0.1 if family == "sum" else 1e-8)
    v, gr = api.nll_grad_reg(hypp, d["xtrainp"], d["ztrainp"], N, family, per)
    vr, grr = O.nll_grad_reg(hypp, d["xtrainp"], d["ztrainp"], N, family, per)
    assert np.isclose(v, vr, rtol=1e-9)
    assert np.allclose(gr, grr, rtol=1e-9, atol=1e-9 * np.abs(grr).max())


def test_reference_third_gradient_component(api, O):
    """python/05_tokamak/SympGPR/func.py:163-167 mixes dK[1] and dK[2] in its third entry; the
    mirror reproduces it on request and otherwise returns the consistent derivative."""
    N = 40
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    vr, gref = O.nll_grad3_reference(hyp, d["xtrain"], d["ztrain"], 2 * N)
    v, gq = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N, with_sig=True, reference_third_component=True)
    assert np.isclose(v, vr, rtol=1e-9)
    assert np.allclose(gq, gref, rtol=1e-7, atol=1e-7 * np.abs(gref).max())
    # consistent d/dsig against a central difference of the oracle's NLL
    _, gc = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N, with_sig=True)
    e = 1e-4 * hyp[2]
    hp_, hm_ = hyp.copy(), hyp.copy()
    hp_[2] += e; hm_[2] -= e
    fd = (O.nll_chol(hp_, d["xtrain"], d["ztrain"], 2 * N) - O.nll_chol(hm_, d["xtrain"], d["ztrain"], 2 * N)) / (2 * e)
    assert np.isclose(gc[2], fd, rtol=1e-5)


def test_fit_returns_alpha_inverse_and_factor(api, O):
    N = 90
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    xt, zt = d["xtrain"], d["ztrain"]
    f = api.fit(hyp, xt, zt, 2 * N, want_inverse=True, want_factor=True)
    K = O.build_k_vec(xt[:N], xt[N:], xt[:N], xt[N:], hyp[:3]) + hyp[3] * np.eye(2 * N)
    Lr = scipy.linalg.cholesky(K, lower=True)
    assert np.allclose(f["L"], Lr, rtol=1e-10, atol=1e-12)
    ar = O.fit_alpha(hyp, xt, zt, 2 * N)
    assert np.allclose(f["alpha"], ar, rtol=1e-9, atol=1e-9 * np.abs(ar).max())
    Kir = scipy.linalg.inv(K)
    assert np.allclose(f["Kyinv"], Kir, rtol=1e-8, atol=1e-9 * np.abs(Kir).max())
    assert np.isclose(f["nll"], O.nll_chol(hyp, xt, zt, 2 * N), rtol=1e-9)
    fr = api.fit(O.timing_hyp(N, d["sigp"]), d["xtrainp"], d["ztrainp"], N, reg=True)
    arr = O.fit_alpha(O.timing_hyp(N, d["sigp"]), d["xtrainp"], d["ztrainp"], N, reg=True)
    assert np.allclose(fr["alpha"], arr, rtol=1e-9, atol=1e-9 * np.abs(arr).max())


def test_nll_grad_medium_size(api, O):
    """n = 2048 (8 x 8 tiles... 16 tiles): exercises the recursive potrf/trsm/trtri and lauum."""
    N = 1024
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    v, gr = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    vr, grr = O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    assert np.isclose(v, vr, rtol=1e-9), (v, vr)
    assert np.allclose(gr, grr, rtol=1e-9, atol=1e-9 * np.abs(grr).max()), (gr, grr)


def test_nll_grad_full_size_against_cpu_golden(api, O):
    """BASELINE's headline size, N = 16384 training pairs (n = 32768): NLL and gradient against the CPU oracle's value
    (tests/golden/fullsize_nll_N16384.json, generated by tests/golden/make_golden_fullsize.py: the oracle's formulas
    evaluated block by block with SciPy dpotrf/dpotri on the host cores), tolerance 1e-9 relative as north_star states;
    plus a size-independent property: the gradient agrees with central differences of the NLL itself."""
    import json, os
    path = os.path.join(os.path.dirname(__file__), "golden", "fullsize_nll_N16384.json")
    g = json.load(open(path))
    N = g["N"]
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    assert np.allclose(hyp, g["hyp"], rtol=0, atol=0)
    v, gr = api.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * N)
    assert abs(v - g["nll"]) <= 1e-9 * abs(g["nll"]), (v, g["nll"])
    assert np.allclose(gr, g["grad"], rtol=1e-9, atol=1e-9 * np.abs(g["grad"]).max()), (gr, g["grad"])
    for t in range(2):
        h = 1e-5 * hyp[t]
        hp, hm = hyp.copy(), hyp.copy()
        hp[t] += h
        hm[t] -= h
        fd = (api.nll_chol(hp, d["xtrain"], d["ztrain"], 2 * N) - api.nll_chol(hm, d["xtrain"], d["ztrain"], 2 * N)) / (2 * h)
        assert abs(fd - gr[t]) <= 2e-5 * abs(gr[t]), (t, fd, gr[t])


# ------------------------------------------------------------------------------- map
def _model(O, N, family="product", lfac=2.0, kch=0.9, guess="dP"):
    """Standard-map model.  guess="dP": ordinary GP trained on P - p as 03/04/05 do (the solver then
    starts at Delta P); guess="P": trained on P as the pendulum scripts 01/02 do."""
    d = O.standard_map_training(N, kch)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    hypp = O.timing_hyp(N, d["sigp"], 1e-8)
    hyp[:2] *= lfac
    hypp[:2] *= lfac
    xt, zt, xtp = d["xtrain"], d["ztrain"], d["xtrainp"]
    ztp = d["ztrainp"] if guess == "dP" else d["P"].copy()
    if guess == "P":
        hypp[2] = 2 * np.amax(np.abs(ztp))**2
    Kyinv = np.linalg.inv(O.build_k_vec(xt[:N], xt[N:], xt[:N], xt[N:], hyp[:3], family) + hyp[3] * np.eye(2 * N))
    Kyinvp = np.linalg.inv(O.buildkreg_vec(xtp[:N], xtp[N:], xtp[:N], xtp[N:], hypp[:3], family) + hypp[3] * np.eye(N))
    return dict(N=N, hyp=hyp[:3].copy(), hypp=hypp[:3].copy(), xt=xt, zt=zt, xtp=xtp, ztp=ztp, Kyinv=Kyinv, Kyinvp=Kyinvp,
                alpha=Kyinv @ zt, alphap=Kyinvp @ ztp)


def _wrapdiff(a, b, wrap):
    d = np.abs(a - b)
    return np.minimum(d, np.abs(d - 2 * np.pi)) if wrap else d


def _check_orbits(C, m, q, p, qr, pr, good, wrapq, wrapp, max_other_root=0.0, family="product"):
    """Trajectories (nm, E) against the oracle's on the `good` orbits.  Up to a fraction
    max_other_root of them may leave the oracle's trajectory, but only by settling on ANOTHER genuine
    root of the reference residual at the step where they part (multi-root residuals)."""
    N = m["N"]
    dq = _wrapdiff(q, qr, wrapq)
    dp = _wrapdiff(p, pr, wrapp)
    bad_step = ((dq > 1e-8) | (dp > 1e-8) | (np.isnan(p) != np.isnan(pr))) & good[None, :]
    bad = np.nonzero(bad_step.any(axis=0))[0]
    assert len(bad) <= max_other_root * max(1, good.sum()), (len(bad), good.sum(), dq[:, good].max(), dp[:, good].max())
    for k in bad:
        i = int(np.argmax(bad_step[:, k]))
        r = C.target_alpha(q[i - 1, k], p[i - 1, k], p[i, k], m["hyp"], m["xt"][:N], m["xt"][N:], m["alpha"], family)
        assert abs(r) < 1e-9, ("left the oracle trajectory on a non-root", k, i, r)
    return len(bad)


def _oracle_map(C, kind, nm, q0, p0, m, want_pdiff=False, start_delta=False):
    """Oracle trajectories + mask of orbits whose every accepted root is a real root (|f| < 1e-10).
    Where the learned map has no root the reference's hybrd1 stops on a non-root (info is ignored,
    sympgpr.f90:107); such orbits are garbage in any implementation and are not compared.
    start_delta: hybrd1 started at p + guess (the oracle twin of the "newton_delta" start)."""
    N = m["N"]
    out = C.applymap_alpha(kind, nm, q0, p0, m["hyp"], m["hypp"], m["xtp"][:N], m["xtp"][N:], m["alphap"], m["xt"][:N],
                           m["xt"][N:], m["alpha"], want_pdiff=want_pdiff, want_notconv=True, start_delta=start_delta)
    good = out[-1] < 1e-10
    return out, good


@pytest.mark.parametrize("kind,kname", [(0, "pendulum"), (1, "henon"), (2, "standard"), (3, "tokamak")])
@pytest.mark.parametrize("solver", ["hybrd", "newton", "newton_delta"])
def test_applymap_matches_oracle(api, O, C, kind, kname, solver):
    """Every map-loop variant x every solver against the oracle started at the SAME point:
       hybrd         = the reference's solver and start (guess as it is)            -> oracle as the reference runs it
       newton        = analytic Newton from the reference's start                   -> same oracle
       newton_delta  = analytic Newton from p + guess (guess GP trained on P - p)   -> oracle with hybrd1 started at p + guess
    1e-8 on every comparable orbit.  Only Newton from the FAR start of the P - p models (kinds 2, 3; neither a default nor
    benchmarked) may settle on another genuine root of a multi-root residual for a few orbits."""
    guess = "P" if kind in (0, 1) else "dP"          # what the respective reference scripts do
    if solver == "newton_delta" and guess == "P":
        pytest.skip("newton_delta is the start for guess GPs trained on P - p (scripts 03/04/05)")
    m = _model(O, 100, guess=guess)
    N = m["N"]
    E, nm = 37, 12
    q0 = O.halton(E, 5) * 2 * np.pi
    p0 = 1.0 + O.halton(E, 7) * 4.0
    if kind == 3:
        p0 = p0 * 0.6          # some orbits fall below 0 / outside r < 0.5 and are lost
    want_pd = kind == 2
    ref, good = _oracle_map(C, kind, nm, q0, p0, m, want_pd, start_delta=(solver == "newton_delta"))
    assert good.sum() >= 0.7 * E
    fn = {0: api.applymap, 1: api.applymap_henon, 2: api.applymap_standard, 3: api.applymap_tok}[kind]
    out = fn(nm, E, m["hyp"], m["hypp"], q0, p0, m["xtp"], m["ztp"], m["Kyinvp"], m["xt"], m["zt"], m["Kyinv"],
             family="product", solver=solver, alphap=m["alphap"], alpha=m["alpha"], return_stats=True)
    q, p = out[0], out[1]
    qr, pr = ref[0], ref[1]
    assert q.shape == (nm, E) and p.shape == (nm, E)
    far_newton = solver == "newton" and guess == "dP"
    _check_orbits(C, m, q, p, qr, pr, good, kind != 1, kind == 2, max_other_root=0.1 if far_newton else 0.0)
    if want_pd and not far_newton:
        assert np.allclose(out[2][:, good], ref[2][:, good], rtol=1e-8, atol=1e-8)
    st = out[-1]
    assert st["evaluations"] > 0
    if kind == 3:
        assert np.isnan(pr).any() and not np.isnan(pr[:, :]).all()


def test_newton_delta_start(api, O, C):
    """SGP_SOLVER_NEWTON_DELTA: Newton started at p + guess for a guess GP trained on P - p (scripts 03/04/05), against the
    oracle's hybrd1 started at the same point: the same trajectories (1e-8) on EVERY orbit whose roots are roots, no
    unconverged step, and fewer residual evaluations than Newton from the reference's far start."""
    m = _model(O, 150, guess="dP")
    E, nm = 64, 12
    q0 = O.halton(E, 5) * 2 * np.pi
    p0 = 1.0 + O.halton(E, 7) * 4.0
    ref, good = _oracle_map(C, 2, nm, q0, p0, m, True, start_delta=True)
    assert good.sum() >= 0.8 * E
    outs = {}
    for solver in ("newton", "newton_delta"):
        outs[solver] = api.applymap_standard(nm, E, m["hyp"], m["hypp"], q0, p0, m["xtp"], m["ztp"], m["Kyinvp"], m["xt"], m["zt"],
                                             m["Kyinv"], solver=solver, alphap=m["alphap"], alpha=m["alpha"], return_stats=True)
    q, p, pd, st = outs["newton_delta"]
    _check_orbits(C, m, q, p, ref[0], ref[1], good, True, True, max_other_root=0.0)
    assert np.allclose(pd[:, good], ref[2][:, good], rtol=1e-8, atol=1e-8)
    assert st["unconverged"] == 0
    assert st["evaluations"] < outs["newton"][-1]["evaluations"]


def test_applymap_first_step_matches_reference_python_layer(api):
    """One map step against tests/golden/path_product.npz (reference calcQ / Pnewton root)."""
    g = np.load(os.path.join(G, "path_product.npz"))
    N = int(g["N"][0])
    for solver in ("hybrd", "newton"):
        q, p, pd = api.applymap_standard(g["map_q"].shape[0], g["map_q"].shape[1], g["hyps"][0, :3], g["hypps"][0, :3],
                                         g["map_q"][0], g["map_p"][0], g["xtrainp"], g["ztrainp"], g["Kyinvp"], g["xtrain"],
                                         g["ztrain"], g["Kyinv"], solver=solver)
        assert np.allclose(q[1], g["map_q"][1], rtol=1e-9, atol=1e-9), solver
        assert np.allclose(p[1], g["map_p"][1], rtol=1e-9, atol=1e-9), solver
        assert np.allclose(q, g["map_q"], rtol=1e-6, atol=1e-6) and np.allclose(p, g["map_p"], rtol=1e-6, atol=1e-6)


def test_applymap_tok_f2py_layout_and_strides(api, O, C):
    m = _model(O, 64)
    N = m["N"]
    E, nm = 9, 5
    q0 = O.halton(E, 5) * 2 * np.pi
    p0 = 1.0 + O.halton(E, 7) * 2.0
    qmap = np.zeros((nm, E, 1), order="F")
    pmap = np.zeros((nm, E, 1), order="F")
    api.applymap_tok_f2py(m["hyp"], m["hypp"], q0, p0, m["xtp"][:N], m["xtp"][N:], m["ztp"], m["Kyinvp"], m["xt"][:N],
                          m["xt"][N:], m["zt"], m["Kyinv"], qmap, pmap)
    (qr, pr, _, _, _), good = _oracle_map(C, 3, nm, q0, p0, m)
    assert np.allclose(qmap[:, good, 0], qr[:, good], rtol=1e-8, atol=1e-8, equal_nan=True)
    assert np.allclose(pmap[:, good, 0], pr[:, good], rtol=1e-8, atol=1e-8, equal_nan=True)
    # strided history and final-only output
    q2, p2, st = api.applymap(nm, E, m["hyp"], m["hypp"], q0, p0, m["xtp"], m["ztp"], m["Kyinvp"], m["xt"], m["zt"],
                              m["Kyinv"], out_every=2, return_stats=True)
    (qf, pf, _, _, _), good = _oracle_map(C, 0, nm, q0, p0, m)
    assert q2.shape == (3, E)
    assert np.allclose(q2[:, good], qf[::2][:, good], rtol=1e-8, atol=1e-8)
    assert np.allclose(p2[:, good], pf[::2][:, good], rtol=1e-8, atol=1e-8)
    assert np.allclose(st["qfinal"][good], qf[-1][good], rtol=1e-8, atol=1e-8)


def test_applymap_multi_chunk_training_set(api, O, C):
    """Nt = 700 > 512: the training set is streamed through shared memory in two chunks."""
    m = _model(O, 700, lfac=2.0, guess="P")
    E, nm = 130, 4          # 130 orbits: two thread blocks, second one partially filled
    q0 = O.halton(E, 5) * 2 * np.pi
    p0 = 1.0 + O.halton(E, 7) * 4.0
    ref, good = _oracle_map(C, 0, nm, q0, p0, m)
    qr, pr = ref[0], ref[1]
    assert good.sum() > 0.8 * E
    for solver in ("hybrd", "newton"):
        q, p = api.applymap(nm, E, m["hyp"], m["hypp"], q0, p0, m["xtp"], m["ztp"], m["Kyinvp"], m["xt"], m["zt"],
                            m["Kyinv"], solver=solver, alphap=m["alphap"], alpha=m["alpha"])
        _check_orbits(C, m, q, p, qr, pr, good, True, False, max_other_root=0.0)


def _rotation_model(O, N=100, a=0.5, l=1.0, noise=1e-8):
    """Isochronous test map: rotation by the angle a, learned with the SE x SE family.  No shear, so
    rounding-level differences between implementations are not amplified over many steps."""
    q = -1 + 2 * O.halton(N, 2)
    p = -1 + 2 * O.halton(N, 3)
    Q = q * np.cos(a) + p * np.sin(a)
    P = -q * np.sin(a) + p * np.cos(a)
    xt = np.hstack((q, P)); zt = np.concatenate((p - P, Q - q)); xtp = np.hstack((q, p)); ztp = P.copy()
    hyp = np.array([l, l, 2 * np.max(np.abs(zt))**2, noise])
    hypp = np.array([l, l, 2 * np.max(np.abs(ztp))**2, noise])
    Kyinv = np.linalg.inv(O.build_k_vec(xt[:N], xt[N:], xt[:N], xt[N:], hyp[:3], "sq") + noise * np.eye(2 * N))
    Kyinvp = np.linalg.inv(O.buildkreg_vec(xtp[:N], xtp[N:], xtp[:N], xtp[N:], hypp[:3], "sq") + noise * np.eye(N))
    return dict(N=N, hyp=hyp[:3].copy(), hypp=hypp[:3].copy(), xt=xt, zt=zt, xtp=xtp, ztp=ztp, Kyinv=Kyinv, Kyinvp=Kyinvp,
                alpha=Kyinv @ zt, alphap=Kyinvp @ ztp)


def test_thousand_steps_within_1e8(api, O, C):
    """BASELINE tolerance: predicted orbits within 1e-8 after 1000 map steps.  Checked on a learned
    rotation (SE x SE kernel, Henon-style loop without wrap): all orbits are regular and there is no
    twist, so the bound is meaningful for every orbit."""
    m = _rotation_model(O)
    N = m["N"]
    E, nm = 24, 1001
    r0 = 0.2 + 0.5 * O.halton(E, 5)
    th = 2 * np.pi * O.halton(E, 7)
    q0, p0 = r0 * np.cos(th), r0 * np.sin(th)
    qa, pa, nev, notconv, maxres = C.applymap_alpha(1, nm, q0, p0, m["hyp"], m["hypp"], m["xtp"][:N], m["xtp"][N:],
                                                    m["alphap"], m["xt"][:N], m["xt"][N:], m["alpha"], "sq",
                                                    want_notconv=True)
    assert maxres.max() < 1e-10
    assert np.abs(np.hypot(qa[-1], pa[-1]) - r0).max() < 1e-3        # the learned map is a rotation (symplectic: no drift)
    for solver in ("hybrd", "newton"):
        q, p, st = api.applymap_henon(nm, E, m["hyp"], m["hypp"], q0, p0, m["xtp"], m["ztp"], m["Kyinvp"], m["xt"], m["zt"],
                                      m["Kyinv"], family="sq", solver=solver, alphap=m["alphap"], alpha=m["alpha"],
                                      out_every=100, return_stats=True)
        assert q.shape == (11, E)
        assert np.abs(q - qa[::100]).max() < 1e-8 and np.abs(p - pa[::100]).max() < 1e-8, (
            solver, np.abs(q - qa[::100]).max(), np.abs(p - pa[::100]).max())
        assert st["unconverged"] <= (0 if solver == "newton" else nm * E)


def test_thousand_steps_standard_map_scaled_by_sensitivity(api, O, C):
    """A twist map shears: an offset of 1e-12 in p0 grows to ~5e-9 over 1000 steps in the ORACLE
    itself (two CPU implementations of hybrd1 differ by ~2e-8 there).  So after 1000 steps the GPU
    must agree with the oracle to 1e-8 or to 100 x the oracle's own sensitivity, whichever is
    larger, and to 1e-8 after the first 100 steps."""
    m = _model(O, 200, lfac=2.0, kch=0.3, guess="P")
    N = m["N"]
    E, nm = 12, 1001
    q0 = O.halton(E, 5) * 2 * np.pi
    p0 = 1.5 + O.halton(E, 7) * 3.0
    (qa, pa, _, _, _), good = _oracle_map(C, 0, nm, q0, p0, m)
    qb, pb, _ = C.applymap_alpha(0, nm, q0, p0 + 1e-12, m["hyp"], m["hypp"], m["xtp"][:N], m["xtp"][N:], m["alphap"],
                                 m["xt"][:N], m["xt"][N:], m["alpha"])
    sens = np.maximum(_wrapdiff(qa[-1], qb[-1], True), np.abs(pa[-1] - pb[-1]))
    assert good.sum() >= 8
    for solver in ("hybrd", "newton"):
        q, p = api.applymap(nm, E, m["hyp"], m["hypp"], q0, p0, m["xtp"], m["ztp"], m["Kyinvp"], m["xt"], m["zt"],
                            m["Kyinv"], solver=solver, alphap=m["alphap"], alpha=m["alpha"], out_every=100)
        d100 = np.maximum(_wrapdiff(q[1], qa[100], True), np.abs(p[1] - pa[100]))
        d1000 = np.maximum(_wrapdiff(q[-1], qa[-1], True), np.abs(p[-1] - pa[-1]))
        assert d100[good].max() < 1e-8, (solver, d100[good].max())
        assert np.all(d1000[good] <= np.maximum(1e-8, 100 * sens[good])), (solver, d1000[good], sens[good])


def test_nan_and_empty_inputs(api, O):
    m = _model(O, 32)
    q0 = np.array([1.0, np.nan, 2.0]); p0 = np.array([2.0, 2.0, np.nan])
    q, p = api.applymap(3, 3, m["hyp"], m["hypp"], q0, p0, m["xtp"], m["ztp"], m["Kyinvp"], m["xt"], m["zt"], m["Kyinv"])
    assert np.isfinite(q[:, 0]).all() and np.isnan(q[1:, 1]).all() and np.isnan(p[1:, 2]).all()
    q, p = api.applymap(3, 0, m["hyp"], m["hypp"], np.zeros(0), np.zeros(0), m["xtp"], m["ztp"], m["Kyinvp"], m["xt"],
                        m["zt"], m["Kyinv"])
    assert q.shape == (3, 0)


# ------------------------------------------------------------------------------- explicit maps
def test_explicit_map_matches_reference_python_layer(api, O):
    """applymap_expl / calcP_expl / nll_expl (python/04_standard_map/func.py:126-141,174-179,256-285) and the
    explicit pendulum loop (python/01_pendulum/explicit/func_expl.py:114-128): the GPU path against the
    reference's own Python functions (tests/golden/path_expl.npz) and against the oracle at a larger size."""
    g = np.load(os.path.join(G, "path_expl.npz"))
    N = int(g["N"][0])
    xt, zt, hyp, Kyinv = g["xtrain"], g["ztrain"], g["hyp"], g["Kyinv"]
    sig2n = float(g["sig2n"][0])
    for k, l in enumerate((0.9, 0.5)):
        assert np.isclose(api.nll_expl([l, hyp[2], sig2n], xt, zt[:N], 2 * N, 0), g["nll_expl_0"][k], rtol=1e-9)
    for k, l in enumerate((0.6, 0.35)):
        assert np.isclose(api.nll_expl([l, hyp[2], sig2n], xt, zt[N:], 2 * N, 1), g["nll_expl_1"][k], rtol=1e-9)
    q0, p0 = g["q0"], g["p0"]
    for k in range(len(q0)):
        assert np.isclose(api.calcP_expl(q0[k], p0[k], hyp, xt, zt, Kyinv), g["calcp_expl"][k], rtol=1e-11, atol=1e-11)
    nm = g["std_q"].shape[0]
    q, p, pd = api.applymap_expl(nm, len(q0), hyp, q0, p0, xt, zt, Kyinv)
    assert np.allclose(q, g["std_q"], rtol=1e-9, atol=1e-9) and np.allclose(p, g["std_p"], rtol=1e-9, atol=1e-9)
    assert np.allclose(pd, g["std_pdiff"], rtol=1e-9, atol=1e-9)
    q, p = api.applymap_expl_pendulum(hyp, q0, p0, xt, zt, Kyinv, len(q0), nm)
    assert np.allclose(q, g["pen_q"], rtol=1e-9, atol=1e-9) and np.allclose(p, g["pen_p"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("family", ["sum", "product"])
def test_explicit_map_matches_oracle(api, O, family):
    """Explicit step with either kernel family, ensemble larger than one warp, training set of several chunks."""
    Nt = 150
    d = O.standard_map_training(Nt)
    xt = np.hstack((d["q"], d["p"]))              # explicit variant: generating function of the OLD momentum
    zt = d["ztrain"]
    hyp = np.array([0.4, 0.5, d["sig"]])
    noise = 1e-2              # cond(Ky) ~ 3e5, |alpha| < 40: a rounding-level change of the sums moves an orbit by
    #                           ~1e-13 in one step and ~20x more with every further step of this (rough) learned map
    Kyinv = np.linalg.inv(O.build_k_vec(xt[:Nt], xt[Nt:], xt[:Nt], xt[Nt:], hyp, family) + noise * np.eye(2 * Nt))
    E, nm = 45, 5
    q0 = O.halton(E, 5) * 2 * np.pi
    p0 = 1.0 + O.halton(E, 7) * 4.0
    qr, pr, pdr = O.applymap_expl("standard", nm, q0, p0, hyp, xt[:Nt], xt[Nt:], zt, Kyinv, family)
    q, p, pd = api.applymap_expl(nm, E, hyp, q0, p0, xt, zt, Kyinv, family=family)
    # one step = two sums of kernel entries weighted with alpha, with heavy cancellation: compared at
    # 1e-14 * sum|alpha_j| * max|K*| (1.4e-8 here; 100x below what 1e-12-accurate kernel entries would allow)
    tol = 1e-14 * np.abs(Kyinv @ zt).sum() * hyp[2] / min(hyp[0], hyp[1])**2
    assert np.allclose(q[1], qr[1], rtol=0, atol=tol) and np.allclose(pd[1], pdr[1], rtol=0, atol=tol)
    # every later step on its own, restarted from the oracle's state (no amplification of rounding differences)
    for i in range(1, nm - 1):
        qi, pi, _ = api.applymap_expl(2, E, hyp, qr[i], pr[i], xt, zt, Kyinv, family=family)
        assert np.allclose(qi[1], qr[i + 1], rtol=0, atol=tol) and np.max(_wrapdiff(pi[1], pr[i + 1], True)) < tol
    # whole trajectories: this rough learned map amplifies a rounding-level difference ~20x per step
    assert np.allclose(q, qr, rtol=1e-6, atol=1e-6) and np.allclose(pd, pdr, rtol=1e-6, atol=1e-6)
    assert np.max(_wrapdiff(p, pr, True)) < 1e-6
    # nll_expl at a size with several tiles
    for ind, l in ((0, 0.1), (1, 0.1)):           # cond ~ 2e5 / 4e5 with this noise (1e7 at l = 0.4, noise 1e-3)
        y = zt[:Nt] if ind == 0 else zt[Nt:]
        assert np.isclose(api.nll_expl([l, d["sig"], 0.1], xt, y, 2 * Nt, ind), O.nll_expl([l, d["sig"], 0.1], xt, y, 2 * Nt, ind),
                          rtol=1e-9)


# ------------------------------------------------------------------------------- split map
def test_split_map_matches_reference_loop(api):
    """The GPU split map against the reference's own applymap_tok loop (Split_SympGPR/func.py:184-219) run in the
    build container with oracle-backed f2py adapters (tests/golden/make_golden_split.py)."""
    g = np.load(os.path.join(G, "path_split.npz"))
    nm, E = g["qmap"].shape
    for solver in ("hybrd", "newton"):
        q, p = api.applymap_tok_split(int(g["nph"][0]), nm, E, g["q0"], g["p0"], g["xtp"], g["ztp"], g["Kyinvp"], g["hypp"],
                                      g["xt"], g["zt"], g["Kyinv"], g["hyp"], solver=solver)
        assert np.array_equal(np.isnan(p), np.isnan(g["pmap"])) and np.array_equal(np.isnan(q), np.isnan(g["qmap"]))
        assert np.nanmax(_wrapdiff(q, g["qmap"], True)) < 1e-8 and np.nanmax(np.abs(p - g["pmap"])) < 1e-8
        assert np.all(q[-1] == 0.0) and np.all(p[-1] == 0.0)


def test_split_map_matches_oracle(api, O):
    """applymap_tok of python/05_tokamak/Split_SympGPR/func.py:184-219: nphmap learned maps in turn, loss test at the
    new angle, whole turns only.  Four sub-maps = four quarter-strength standard-map kicks, so one turn is a map
    with regular orbits; training sets / hyper-parameters differ per sub-map."""
    nph, N = 4, 60
    xtp = np.zeros((2 * N, nph)); ztp = np.zeros((N, nph)); xt = np.zeros((2 * N, nph)); zt = np.zeros((2 * N, nph))
    hyp = np.zeros((nph, 3)); hypp = np.zeros((nph, 3)); Kyinv = np.zeros((nph, 2 * N, 2 * N)); Kyinvp = np.zeros((nph, N, N))
    for m in range(nph):
        q = O.halton(N, 2, start=1 + 7 * m) * 2 * np.pi
        p = 0.5 + O.halton(N, 3, start=1 + 5 * m) * 5.0
        P = p + 0.1 * (1 + 0.2 * m) * np.sin(q)
        Q = q + 0.25 * P
        xt[:, m] = np.hstack((q, P)); zt[:, m] = np.concatenate((p - P, Q - q))
        xtp[:, m] = np.hstack((q, p)); ztp[:, m] = P
        l = 1.2 + 0.1 * m
        hyp[m] = [l, l, 2 * np.max(np.abs(zt[:, m]))**2]
        hypp[m] = [l, l, 2 * np.max(np.abs(ztp[:, m]))**2]
        Kyinv[m] = np.linalg.inv(O.build_k_vec(q, P, q, P, hyp[m]) + 1e-6 * np.eye(2 * N))
        Kyinvp[m] = np.linalg.inv(O.buildkreg_vec(q, p, q, p, hypp[m]) + 1e-6 * np.eye(N))
    E, nm = 9, 14                       # 14 - 4 = 10 -> three whole turns = 12 steps, last row stays zero
    q0 = O.halton(E, 5) * 2 * np.pi
    p0 = 0.3 + O.halton(E, 7) * 4.5     # some start below the loss boundary region, some get lost
    qr, pr = O.applymap_tok_split(nph, nm, q0, p0, xtp, ztp, Kyinvp, hypp, xt, zt, Kyinv, hyp)
    q, p = api.applymap_tok_split(nph, nm, E, q0, p0, xtp, ztp, Kyinvp, hypp, xt, zt, Kyinv, hyp)
    assert q.shape == (nm, E)
    assert np.array_equal(np.isnan(p), np.isnan(pr)) and np.array_equal(np.isnan(q), np.isnan(qr))
    assert np.all(q[13] == 0.0) and np.all(p[13] == 0.0) and np.all(qr[13] == 0.0)
    assert np.nanmax(_wrapdiff(q, qr, True)) < 1e-8 and np.nanmax(np.abs(p - pr)) < 1e-8
    q2, p2 = api.applymap_tok_split(nph, nm, E, q0, p0, xtp, ztp, Kyinvp, hypp, xt, zt, Kyinv, hyp, solver="newton")
    assert np.nanmax(_wrapdiff(q2, qr, True)) < 1e-8 and np.nanmax(np.abs(p2 - pr)) < 1e-8


# ------------------------------------------------------------------------------- X1: 2-DOF 4x4-block kernel
@pytest.mark.parametrize("N,N0", [(1, 1), (5, 9), (130, 67)])
def test_dof2_fill_matches_oracle(api, O, N, N0):
    rng = np.random.default_rng(N + 10 * N0)
    x, x0 = rng.uniform(-1, 1, 4 * N), rng.uniform(-1, 1, 4 * N0)
    hyp = np.array([0.6, 0.8, 1.7])
    K = np.full((4 * N, 4 * N0), np.nan, order="F")
    api.build_k4(x, x0, hyp, K)
    assert np.allclose(K, O.build_k4(x, x0, hyp), rtol=1e-12, atol=1e-12 * hyp[2] / 0.36)


@pytest.mark.parametrize("N", [8, 40, 100, 200])
def test_dof2_nll_and_gradient_match_oracle(api, O, N):
    """BASELINE config 3's training side (2-DOF, 4x4 blocks), parity against the (self-validated) oracle."""
    x, z = O.henon_like_training(N)
    hyp = np.array([0.35, 0.4, 2 * np.max(np.abs(z))**2, 1e-4])          # cond(Ky) ~ 4e5 at N = 200
    v, g = api.nll_grad4(hyp, x, z, 4 * N, with_sig=True)
    vr, gr = O.nll_grad4(hyp, x, z, 4 * N, with_sig=True)
    assert np.isclose(v, vr, rtol=1e-9), (v, vr)
    assert np.allclose(g, gr, rtol=1e-8, atol=1e-8 * np.abs(gr).max()), (g, gr)
    assert np.isclose(api.nll_chol4(hyp, x, z, 4 * N), vr, rtol=1e-9)


# ------------------------------------------------------------------- quality metrics, generators
@pytest.mark.parametrize("energy", ["pendulum", "tokamak"])
def test_fused_quality_matches_history_based_quality(api, O, energy):
    """Eosc = std(H)/mean(H) accumulated inside the map kernel against `quality` evaluated on the full history the
    reference would hold (python/functions/func.py:262-272; energy: 01_pendulum/implicit/func.py:116-117 and
    Split_SympGPR/func.py:234-246)."""
    m = _model(O, 120, guess="P")
    E, nm = 70, 41
    q0 = O.halton(E, 5) * 2 * np.pi
    if energy == "pendulum":
        kind, fn, par, e_every = "pendulum", api.applymap, (0.7,), 1
        p0 = 1.0 + O.halton(E, 7) * 4.0
    else:
        kind, fn, par, e_every = "tokamak", api.applymap_tok, (0.001, 3.0, 0.2), 4
        p0 = (1.0 + O.halton(E, 7) * 4.0) * 0.8          # some orbits are lost: NaN energy, NaN Eosc
    args = (m["hyp"], m["hypp"], q0, p0, m["xtp"], m["ztp"], m["Kyinvp"], m["xt"], m["zt"], m["Kyinv"])
    q, p = fn(nm, E, *args, solver="newton", alphap=m["alphap"], alpha=m["alpha"])[:2]
    out = api.applymap_quality(kind, nm, E, *args, energy=energy, energy_par=par, e_every=e_every, solver="newton",
                               alphap=m["alphap"], alpha=m["alpha"])
    qs, ps = q[::e_every], p[::e_every]
    if energy == "pendulum":
        H = O.energy_pendulum(qs, ps, par[0])
    else:
        H = O.energy_tok(qs, ps, par[0], par[1], par[2]).T
    with np.errstate(invalid="ignore"):
        ref = O.quality_eosc(H)
        hm = H.mean(axis=0)
    assert np.array_equal(np.isnan(out["Eosc"]), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert ok.sum() > E // 2
    if energy == "tokamak":
        assert (~ok).any()
    assert np.allclose(out["Eosc"][ok], ref[ok], rtol=1e-7, atol=1e-13), np.abs(out["Eosc"][ok] / ref[ok] - 1).max()
    assert np.allclose(out["Hmean"][ok], hm[ok], rtol=1e-12)
    assert np.array_equal(out["q1"], qs[1], equal_nan=True) and np.array_equal(out["p1"], ps[1], equal_nan=True)
    assert np.array_equal(out["qfinal"][ok], q[-1][ok])
    # the reference's quality() on the same numbers: gd against a shifted "reference orbit"
    ys = np.zeros((3, 2, E)); ys[2, 0] = qs[1] + 1e-3; ys[2, 1] = ps[1] - 2e-3
    Eosc, gd, stdgd = api.quality(out["q1"], out["p1"], out["Eosc"], ys, 2)
    assert np.allclose(gd[ok], 0.5 * (1e-6 + 4e-6), rtol=1e-6)


def test_standard_map_iterate_matches_reference_loop(api, O):
    N, nm, k = 50, 25, 0.9
    X0 = np.vstack((O.halton(N, 2) * 2 * np.pi, O.halton(N, 3) * 2 * np.pi))
    f = api.StandardMapIterate(k, nm, N, X0)
    fr = O.standard_map_iterate(k, nm, N, X0)
    assert f.shape == (2, N, nm)
    assert np.array_equal(f[:, :, 0], X0)
    assert np.allclose(f, fr, rtol=1e-11, atol=1e-11)
    assert np.allclose(f[:, :, 1], fr[:, :, 1], rtol=1e-15, atol=1e-15)
    with pytest.raises(ValueError):
        api.StandardMapIterate(k, nm, N, X0.T)


def test_learned_map_is_symplectic_at_full_ensemble_size(api, O):
    """Size-independent property at BASELINE's map size (Nt = 4096 training pairs, 1e5 orbits): a map defined through a
    mixed-variable generating function, P = p - F_q(q, P), Q = q + F_P(q, P), is symplectic whatever F is -- the
    Jacobian of (q, p) -> (Q, P) has determinant 1 (the reason the reference solves the implicit equation at all).
    Checked by central differences through the whole GPU path (guess GP, root solve, dQ sweep): any error in the
    kernel sums, the analytic derivative or the root accuracy shows up as det != 1."""
    Nt, E, eps = 4096, 100000, 1e-4
    d = O.standard_map_training(Nt)
    hyp = O.timing_hyp(Nt, d["sig"], 1e-8)
    hypp = O.timing_hyp(Nt, d["sigp"], 1e-8)
    hyp[:2] *= 2.0
    hypp[:2] *= 2.0
    f = api.fit(hyp, d["xtrain"], d["ztrain"], 2 * Nt)
    fp = api.fit(hypp, d["xtrainp"], d["ztrainp"], Nt, reg=True)
    q0 = 0.5 + 5.0 * O.halton(E, 5)
    p0 = 1.0 + 4.0 * O.halton(E, 7)
    Q, P = [], []
    for dq, dp in ((eps, 0.0), (-eps, 0.0), (0.0, eps), (0.0, -eps)):
        out = api.applymap_henon(2, E, hyp[:3], hypp[:3], q0 + dq, p0 + dp, d["xtrainp"], None, None, d["xtrain"], None, None,
                                 family="product", solver="newton_delta", alphap=fp["alpha"], alpha=f["alpha"], out_every=0,
                                 return_stats=True)
        Q.append(out[0]); P.append(out[1])
        assert out[-1]["unconverged"] <= 1e-4 * E
    Qq, Qp = (Q[0] - Q[1]) / (2 * eps), (Q[2] - Q[3]) / (2 * eps)
    Pq, Pp = (P[0] - P[1]) / (2 * eps), (P[2] - P[3]) / (2 * eps)
    det = Qq * Pp - Qp * Pq
    err = np.abs(det - 1.0)
    # accuracy of the check: the residual of this model bottoms out near 4e-12 (rounding in the sums), so the difference
    # quotients carry ~4e-12 / eps of noise (the CPU oracle gives 7e-9 median); a wrong term in any sweep shows up at
    # the 1e-2 .. 1 level
    assert np.isfinite(det).mean() > 0.999
    assert np.nanmedian(err) < 1e-6, np.nanmedian(err)
    assert np.nanquantile(err, 0.99) < 1e-4, np.nanquantile(err, 0.99)
    # and the learned map is the standard map it was trained on (K = 0.9), to the accuracy of the regression
    Pc = 0.5 * (P[2] + P[3])
    assert np.nanmedian(np.abs(Pc - (p0 + 0.9 * np.sin(q0)))) < 1e-4


# ------------------------------------------------------------------- 2-DOF map (BASELINE config 3, not in the reference)
def _dof2_model(api, O, N):
    x, z = O.henon_like_training(N)
    hyp = np.array([0.35, 0.4, 2 * np.max(np.abs(z))**2, 1e-8])
    f = api.fit(hyp, x, z, 4 * N, reg=4)
    return x, z, hyp, f["alpha"]


def test_dof2_map_matches_oracle(api, O):
    """map4_kernel (2 x 2 Newton with the analytic Jacobian) against oracle.applymap4, whose grad F is the 4 x 4-block
    matrix of build_k4 times alpha and whose Jacobian is taken by differences: same roots, same orbits."""
    N = 150
    x, z, hyp, alpha = _dof2_model(api, O, N)
    import scipy.linalg
    Ky = O.build_k4(x, x, hyp[:3]) + hyp[3] * np.eye(4 * N)
    alpha_ref = scipy.linalg.cho_solve(scipy.linalg.cho_factor(Ky, lower=True), z)
    assert np.allclose(alpha, alpha_ref, rtol=1e-6, atol=1e-6 * np.abs(alpha_ref).max())
    E, nm = 45, 9
    q0 = np.vstack((-0.25 + 0.5 * O.halton(E, 2, start=50), -0.25 + 0.5 * O.halton(E, 3, start=50)))
    p0 = np.vstack((-0.25 + 0.5 * O.halton(E, 5, start=50), -0.25 + 0.5 * O.halton(E, 7, start=50)))
    qr, pr = O.applymap4(nm, q0, p0, hyp[:3], x, alpha)
    q, p, st = api.applymap4(nm, E, hyp[:3], q0, p0, x, alpha, return_stats=True)
    assert q.shape == (nm, 2, E)
    assert np.array_equal(q[0], q0) and np.array_equal(p[0], p0)
    assert np.abs(q - qr).max() < 1e-9 and np.abs(p - pr).max() < 1e-9, (np.abs(q - qr).max(), np.abs(p - pr).max())
    assert st["unconverged"] == 0 and st["evaluations"] >= 3 * E * (nm - 1)
    assert np.array_equal(st["qfinal"], q[-1]) and np.array_equal(st["pfinal"], p[-1])
    qs, ps = api.applymap4(nm, E, hyp[:3], q0, p0, x, alpha, out_every=4)
    assert qs.shape == (3, 2, E) and np.array_equal(qs[2], q[8]) and np.array_equal(ps[1], p[4])
    with pytest.raises(ValueError):
        api.applymap4(nm, E, hyp[:3], q0.T, p0, x, alpha)


def test_dof2_map_is_symplectic_and_follows_the_training_map(api, O):
    """1e5 orbits (BASELINE config 3's ensemble size): M^T Omega M = Omega for the Jacobian M of one learned step, and the
    step agrees with the kick-drift map the model was trained on."""
    N, E, eps = 400, 100000, 1e-5
    x, z, hyp, alpha = _dof2_model(api, O, N)
    q0 = np.vstack((-0.3 + 0.6 * O.halton(E, 2, start=7), -0.3 + 0.6 * O.halton(E, 3, start=7)))
    p0 = np.vstack((-0.3 + 0.6 * O.halton(E, 5, start=7), -0.3 + 0.6 * O.halton(E, 7, start=7)))
    M = np.zeros((4, 4, E))
    for c in range(4):
        dq = np.zeros((2, 1)); dp = np.zeros((2, 1))
        (dq if c < 2 else dp)[c % 2] = eps
        qa, pa = api.applymap4(2, E, hyp[:3], q0 + dq, p0 + dp, x, alpha, out_every=0)
        qb, pb = api.applymap4(2, E, hyp[:3], q0 - dq, p0 - dp, x, alpha, out_every=0)
        M[:, c] = np.vstack((qa - qb, pa - pb)) / (2 * eps)
    Om = np.block([[np.zeros((2, 2)), np.eye(2)], [-np.eye(2), np.zeros((2, 2))]])
    R = np.einsum("ijk,il,lmk->jmk", M, Om, M) - Om[:, :, None]
    err = np.abs(R).max(axis=(0, 1))
    assert np.isfinite(err).all()
    assert np.median(err) < 1e-6 and err.max() < 1e-4, (np.median(err), err.max())    # difference quotients: ~1e-12 / eps of noise
    qf, pf = api.applymap4(2, E, hyp[:3], q0, p0, x, alpha, out_every=0)
    dt = 0.3
    P = np.vstack((p0[0] - dt * (q0[0] + 2 * q0[0] * q0[1]), p0[1] - dt * (q0[1] + q0[0]**2 - q0[1]**2)))
    assert np.abs(pf - P).max() < 1e-3 and np.abs(qf - (q0 + dt * P)).max() < 1e-3
