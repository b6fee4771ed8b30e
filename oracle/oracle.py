"""TEST INFRASTRUCTURE -- CPU oracle, not product code.

NumPy/SciPy restatement of the SympGPR hot path.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (sympgpr_b200) never does.

Every function cites the reference file:line it follows.  Loops are written
the way the reference writes them ("literal" functions, for small sizes) with
a vectorised twin where the sizes of BASELINE.json need it; the twins are
checked against the literal versions in tests/test_oracle.py.

Parity pin (DESIGN.md "Oracle"): the scalar forms are pinned to the
reference's SymPy derivation (tests/golden/kernel_forms_*.npz); the matrix /
NLL / map functions are pinned to the reference's own pure-Python loop
implementations (python/04_standard_map/func.py, python/02_pert_pendulum/
func.py) executed in the build container with those scalar forms injected
(tests/golden/make_golden_path.py -> tests/golden/path_*.npz).
"""
import numpy as np
import scipy.linalg
import scipy.optimize

from .kernel_forms import FAMILIES

TWO_PI = 2.0 * np.pi


def _fam(family):
    return FAMILIES[family] if isinstance(family, str) else family


def _extra(family, p):
    return (p,) if _fam(family).nargs == 7 else ()


# ---------------------------------------------------------------- fills (F1, F2)
def build_k(x, y, x0, y0, hyp, K, family="product", p=0.5):
    """sympgpr.f90:12-38 -- fill K(2N,2N0) in place, literal double loop.

    N = size(K,1)/2, N0 = size(K,2)/2 (integer division); a=(x0_j,y0_j),
    b=(x_i,y_i); K *= hyp(3) at the end (also scales untouched odd rows/cols).
    """
    f = _fam(family)
    ex = _extra(family, p)
    N = K.shape[0] // 2
    N0 = K.shape[1] // 2
    for j in range(N0):
        for i in range(N):
            K[i, j] = f.d2kdxdx0_num(x0[j], y0[j], x[i], y[i], hyp[0], hyp[1], *ex)
            K[N + i, j] = f.d2kdxdy0_num(x0[j], y0[j], x[i], y[i], hyp[0], hyp[1], *ex)
            K[i, N0 + j] = f.d2kdxdy0_num(x0[j], y0[j], x[i], y[i], hyp[0], hyp[1], *ex)
            K[N + i, N0 + j] = f.d2kdydy0_num(x0[j], y0[j], x[i], y[i], hyp[0], hyp[1], *ex)
    K[...] = hyp[2] * K


def build_k_vec(x, y, x0, y0, hyp, family="product", p=0.5):
    """Vectorised twin of build_k; returns a new F-ordered (2N,2N0) array."""
    f = _fam(family)
    ex = _extra(family, p)
    x = np.asarray(x, float); y = np.asarray(y, float)
    x0 = np.asarray(x0, float); y0 = np.asarray(y0, float)
    N, N0 = len(x), len(x0)
    xa, ya = x0[None, :], y0[None, :]
    xb, yb = x[:, None], y[:, None]
    K = np.empty((2 * N, 2 * N0), order="F")
    K[:N, :N0] = f.d2kdxdx0_num(xa, ya, xb, yb, hyp[0], hyp[1], *ex)
    kxy = f.d2kdxdy0_num(xa, ya, xb, yb, hyp[0], hyp[1], *ex)
    K[N:, :N0] = kxy
    K[:N, N0:] = kxy
    K[N:, N0:] = f.d2kdydy0_num(xa, ya, xb, yb, hyp[0], hyp[1], *ex)
    K *= hyp[2]
    return K


def buildkreg(x, y, x0, y0, hyp, K, family="product", p=0.5):
    """sympgpr.f90:40-60 -- plain kernel matrix K(N,N0) in place."""
    f = _fam(family)
    ex = _extra(family, p)
    N, N0 = K.shape
    for j in range(N0):
        for i in range(N):
            K[i, j] = f.kern_num(x0[j], y0[j], x[i], y[i], hyp[0], hyp[1], *ex)
    K[...] = hyp[2] * K


def buildkreg_vec(x, y, x0, y0, hyp, family="product", p=0.5):
    f = _fam(family)
    ex = _extra(family, p)
    x = np.asarray(x, float); y = np.asarray(y, float)
    x0 = np.asarray(x0, float); y0 = np.asarray(y0, float)
    K = f.kern_num(x0[None, :], y0[None, :], x[:, None], y[:, None], hyp[0], hyp[1], *ex)
    return np.asfortranarray(hyp[2] * K)


# ------------------------------------------------- hyper-parameter derivatives (F3)
def build_dk(xin, x0in, hyp, family="product", p=0.5, with_sig=False):
    """python/02_pert_pendulum/func.py:80-129 -- dense dK/dlx, dK/dly.

    Loop index roles are transposed with respect to build_K exactly as in the
    reference (k over x0, lk over x); with_sig appends the third component of
    python/05_tokamak/SympGPR/func.py:101-116 (unscaled K, i.e. K/sig).
    """
    f = _fam(family)
    ex = _extra(family, p)
    xin = np.asarray(xin, float); x0in = np.asarray(x0in, float)
    N = len(xin) // 2
    N0 = len(x0in) // 2
    lx, ly, sig = hyp[0], hyp[1], hyp[2]
    x0, y0 = x0in[:N0], x0in[N0:2 * N0]
    x, y = xin[:N], xin[N:2 * N]
    xa, ya = x0[:, None], y0[:, None]      # k index (rows)
    xb, yb = x[None, :], y[None, :]        # lk index (cols)
    out = []
    for sfx in ("dlx", "dly"):
        k11 = sig * getattr(f, "d3kdxdx0" + sfx + "_num")(xa, ya, xb, yb, lx, ly, *ex)
        k21 = sig * getattr(f, "d3kdxdy0" + sfx + "_num")(xa, ya, xb, yb, lx, ly, *ex)
        k22 = sig * getattr(f, "d3kdydy0" + sfx + "_num")(xa, ya, xb, yb, lx, ly, *ex)
        out.append(np.vstack([np.hstack([k11, k21]), np.hstack([k21, k22])]))
    if with_sig:
        xa, ya = x0[None, :], y0[None, :]
        xb, yb = x[:, None], y[:, None]
        k11 = f.d2kdxdx0_num(xa, ya, xb, yb, lx, ly, *ex)
        k21 = f.d2kdxdy0_num(xa, ya, xb, yb, lx, ly, *ex)
        k22 = f.d2kdydy0_num(xa, ya, xb, yb, lx, ly, *ex)
        out.append(np.vstack([np.hstack([k11, k21]), np.hstack([k21, k22])]))
    return out


def build_dkreg(xin, x0in, hyp, family="product", p=0.5):
    """python/02_pert_pendulum/func.py:52-78."""
    f = _fam(family)
    ex = _extra(family, p)
    xin = np.asarray(xin, float); x0in = np.asarray(x0in, float)
    N = len(xin) // 2
    N0 = len(x0in) // 2
    lx, ly, sig = hyp[0], hyp[1], hyp[2]
    x0, y0 = x0in[:N0], x0in[N0:2 * N0]
    x, y = xin[:N], xin[N:2 * N]
    xa, ya = x0[None, :], y0[None, :]
    xb, yb = x[:, None], y[:, None]
    return [sig * f.dkdlx_num(xa, ya, xb, yb, lx, ly, *ex),
            sig * f.dkdly_num(xa, ya, xb, yb, lx, ly, *ex)]


# ------------------------------------------------------------ NLL (L1, L2, G1)
def _split(xin, n_half):
    xin = np.asarray(xin, float)
    return xin[:n_half], xin[n_half:2 * n_half]


def solve_cholesky(L, b):
    """python/02_pert_pendulum/func.py:173-177."""
    return scipy.linalg.solve_triangular(
        L.T, scipy.linalg.solve_triangular(L, b, lower=True, check_finite=False),
        lower=False, check_finite=False)


def nll_chol(hyp, x, y, n, family="product", p=0.5):
    """python/05_tokamak/SympGPR/func.py:143-150: hyp=[lx,ly,sig,sig2n], n=2N."""
    xs, ys = _split(x, n // 2)
    K = build_k_vec(xs, ys, xs, ys, hyp[:3], family, p)
    Ky = K + np.abs(hyp[3]) * np.eye(n)
    L = scipy.linalg.cholesky(Ky, lower=True)
    alpha = solve_cholesky(L, y)
    return 0.5 * y.T.dot(alpha) + np.sum(np.log(L.diagonal()))


def nll_chol_reg(hyp, x, y, n, family="product", p=0.5):
    """python/05_tokamak/SympGPR/func.py:134-141 (plain kernel, n=N)."""
    xs, ys = _split(x, n)
    K = buildkreg_vec(xs, ys, xs, ys, hyp[:3], family, p)
    Ky = K + np.abs(hyp[3]) * np.eye(n)
    L = scipy.linalg.cholesky(Ky, lower=True)
    alpha = solve_cholesky(L, y)
    return 0.5 * y.T.dot(alpha) + np.sum(np.log(L.diagonal()))


def nll_grad_literal(hyp, x, y, n, family="product", p=0.5):
    """python/02_pert_pendulum/func.py:148-162 -- LU inverse, slogdet, dense
    dK and full Kyinv@dK products, exactly as the reference does it."""
    xs, ys = _split(x, n // 2)
    K = build_k_vec(xs, ys, xs, ys, hyp[:3], family, p)
    Ky = K + np.abs(hyp[3]) * np.diag(np.ones(n))
    Kyinv = np.linalg.inv(Ky)
    alpha = Kyinv.dot(y)
    val = 0.5 * y.T.dot(alpha) + 0.5 * np.linalg.slogdet(Ky)[1]
    dK = build_dk(x, x, hyp[:3], family, p)
    grad = np.array([
        -0.5 * alpha.T.dot(dK[0].dot(alpha)) + 0.5 * np.trace(Kyinv.dot(dK[0])),
        -0.5 * alpha.T.dot(dK[1].dot(alpha)) + 0.5 * np.trace(Kyinv.dot(dK[1])),
    ])
    return val, grad


def nll_grad(hyp, x, y, n, family="product", p=0.5, with_sig=False):
    """Same value/gradient as nll_grad_literal through Cholesky + dpotri and
    the elementwise contraction -0.5*sum((aa^T - Ky^-1) o dK) (no n^3 GEMM).
    with_sig adds d/dsig = -0.5*sum((aa^T - Ky^-1) o K/sig) -- the
    mathematically consistent third component (the reference's own third
    component mixes dK[1] and dK[2]; see
    nll_grad3_reference)."""
    xs, ys = _split(x, n // 2)
    K = build_k_vec(xs, ys, xs, ys, hyp[:3], family, p)
    Ky = K + np.abs(hyp[3]) * np.eye(n)
    L = scipy.linalg.cholesky(Ky, lower=True)
    alpha = solve_cholesky(L, y)
    val = 0.5 * y.T.dot(alpha) + np.sum(np.log(L.diagonal()))
    Kinv, info = scipy.linalg.lapack.dpotri(L, lower=1)
    assert info == 0
    Kinv = np.tril(Kinv) + np.tril(Kinv, -1).T
    W = np.outer(alpha, alpha) - Kinv
    dK = build_dk(x, x, hyp[:3], family, p, with_sig=with_sig)
    grad = np.array([-0.5 * np.sum(W * d) for d in dK])
    return val, grad


def nll_grad3_reference(hyp, x, y, n, family="product", p=0.5):
    """python/05_tokamak/SympGPR/func.py:152-168 including its third component
    as written (quadratic term uses dK[1], trace term dK[2])."""
    xs, ys = _split(x, n // 2)
    K = build_k_vec(xs, ys, xs, ys, hyp[:3], family, p)
    Ky = K + np.abs(hyp[3]) * np.diag(np.ones(n))
    Kyinv = np.linalg.inv(Ky)
    L = scipy.linalg.cholesky(Ky, lower=True)
    alpha = solve_cholesky(L, y)
    val = 0.5 * y.T.dot(alpha) + np.sum(np.log(L.diagonal()))
    dK = build_dk(x, x, hyp[:3], family, p, with_sig=True)
    alpha = Kyinv.dot(y)
    grad = np.array([
        -0.5 * alpha.T.dot(dK[0].dot(alpha)) + 0.5 * np.trace(Kyinv.dot(dK[0])),
        -0.5 * alpha.T.dot(dK[1].dot(alpha)) + 0.5 * np.trace(Kyinv.dot(dK[1])),
        -0.5 * alpha.T.dot(dK[1].dot(alpha)) + 0.5 * np.trace(Kyinv.dot(dK[2])),
    ])
    return val, grad


def nll_grad_reg(hyp, x, y, n, family="product", p=0.5):
    """python/02_pert_pendulum/func.py:132-146 via Cholesky/potri (plain kernel)."""
    xs, ys = _split(x, n)
    K = buildkreg_vec(xs, ys, xs, ys, hyp[:3], family, p)
    Ky = K + np.abs(hyp[3]) * np.eye(n)
    L = scipy.linalg.cholesky(Ky, lower=True)
    alpha = solve_cholesky(L, y)
    val = 0.5 * y.T.dot(alpha) + np.sum(np.log(L.diagonal()))
    Kinv, info = scipy.linalg.lapack.dpotri(L, lower=1)
    assert info == 0
    Kinv = np.tril(Kinv) + np.tril(Kinv, -1).T
    W = np.outer(alpha, alpha) - Kinv
    dK = build_dkreg(x, x, hyp[:3], family, p)
    return val, np.array([-0.5 * np.sum(W * d) for d in dK])


def fit_alpha(hyp, x, z, n, reg=False, family="product", p=0.5):
    """alpha = (K + |sig2n| I)^-1 z: what every main.py gets as Kyinv.dot(ztrain)
    (e.g. python/01_pendulum/implicit/main.py:159-165)."""
    if reg:
        xs, ys = _split(x, n)
        K = buildkreg_vec(xs, ys, xs, ys, hyp[:3], family, p)
    else:
        xs, ys = _split(x, n // 2)
        K = build_k_vec(xs, ys, xs, ys, hyp[:3], family, p)
    Ky = K + np.abs(hyp[3]) * np.eye(n)
    L = scipy.linalg.cholesky(Ky, lower=True)
    return solve_cholesky(L, z)


# ----------------------------------------------------------- prediction (M1-M3)
def guessp(x, y, hypp, xtrainp, ytrainp, ztrainp, kyinvp, family="product", p=0.5):
    """sympgpr.f90:62-73."""
    Kstar = buildkreg_vec(np.atleast_1d(x), np.atleast_1d(y), xtrainp, ytrainp, hypp, family, p)
    return float(Kstar[0, :].dot(np.asarray(kyinvp, float).dot(ztrainp)))


def calcq(x, y, xtrain, ytrain, hyp, kyinv, ztrain, family="product", p=0.5):
    """sympgpr.f90:75-86."""
    Kstar = build_k_vec(np.atleast_1d(x), np.atleast_1d(y), xtrain, ytrain, hyp, family, p)
    return float(Kstar[1, :].dot(np.asarray(kyinv, float).dot(ztrain)))


def target(P, x, y, hyp, xtrain, ytrain, ztrain, kyinv, family="product", p=0.5):
    """sympgpr.f90:112-124: f(P) = K*(1,:).(Kyinv ztrain) - p + P."""
    Kstar = build_k_vec(np.atleast_1d(x), np.atleast_1d(P), xtrain, ytrain, hyp, family, p)
    return float(Kstar[0, :].dot(np.asarray(kyinv, float).dot(ztrain))) - float(y) + float(P)


def calcp(x, y, hyp, hypp, xtrainp, ytrainp, ztrainp, kyinvp, xtrain, ytrain, ztrain, kyinv,
          family="product", p=0.5, full_output=False):
    """sympgpr.f90:88-125: hybrd1(target, n=1, x0=guessP, tol=1e-13).

    hybrd1's parameter block (minpack.f90:1570-1577: maxfev=200*(n+1), ml=mu=n-1,
    epsfcn=0, mode=2, diag=1, factor=100) is passed to SciPy's own MINPACK
    hybrd through fsolve.
    """
    pg = guessp(x, y, hypp, xtrainp, ytrainp, ztrainp, kyinvp, family, p)
    alpha = np.asarray(kyinv, float).dot(ztrain)
    nt = len(xtrain)

    def f(P):
        Ks = build_k_vec(np.atleast_1d(x), np.atleast_1d(P[0]), xtrain, ytrain, hyp, family, p)
        return [float(Ks[0, :].dot(alpha)) - float(y) + P[0]]

    sol, info, ier, _ = scipy.optimize.fsolve(
        f, [pg], xtol=1e-13, maxfev=400, diag=[1.0], factor=100, epsfcn=0.0, full_output=True)
    if full_output:
        return float(sol[0]), pg, info["nfev"], ier
    return float(sol[0])


def calcp_expl(x, y, hyp, xtrain, ytrain, ztrain, kyinv, family="sum", p=0.5):
    """Explicit map, momentum step: P = p - F_q(q, p) with the generating function taken at the OLD momentum.
    calcP_expl python/04_standard_map/func.py:174-179 (returns -pGP[0] + y); calcP
    python/01_pendulum/explicit/func_expl.py:107-112 (returns -pGP[0], added to p by its caller :121)."""
    Kstar = build_k_vec(np.atleast_1d(x), np.atleast_1d(y), xtrain, ytrain, hyp, family, p)
    return float(y) - float(Kstar[0, :].dot(np.asarray(kyinv, float).dot(ztrain)))


def applymap_expl(variant, nm, q0, p0, hyp, xtrain, ytrain, ztrain, kyinv, family="sum", p=0.5):
    """Explicit ensemble loops.  variant "standard": applymap_expl python/04_standard_map/func.py:256-285
    (pdiff, p = mod(P, 2pi), q = q + dq NOT wrapped) -> (qmap, pmap, pdiff); variant "pendulum": applymap
    python/01_pendulum/explicit/func_expl.py:114-128 (q = mod(q + dq, 2pi), p not wrapped) -> (qmap, pmap)."""
    q0 = np.asarray(q0, float); p0 = np.asarray(p0, float)
    E = len(q0)
    pmap = np.zeros((nm, E)); qmap = np.zeros((nm, E)); pdiff = np.zeros((nm, E))
    pmap[0] = p0; qmap[0] = q0; pdiff[0] = p0
    for i in range(nm - 1):
        for k in range(E):
            pmap[i + 1, k] = calcp_expl(qmap[i, k], pmap[i, k], hyp, xtrain, ytrain, ztrain, kyinv, family, p)
            if variant == "standard":
                pdiff[i + 1, k] = pdiff[i, k] + (pmap[i + 1, k] - pmap[i, k])
                pmap[i + 1, k] = np.mod(pmap[i + 1, k], TWO_PI)
        for k in range(E):
            if np.isnan(pmap[i + 1, k]):
                qmap[i + 1, k] = np.nan
            else:
                dq = calcq(qmap[i, k], pmap[i + 1, k], xtrain, ytrain, hyp, kyinv, ztrain, family, p)
                qmap[i + 1, k] = dq + qmap[i, k] if variant == "standard" else np.mod(dq + qmap[i, k], TWO_PI)
    if variant == "standard":
        return qmap, pmap, pdiff
    return qmap, pmap


def applymap_tok_split(nphmap, nm, q0, p0, xtrainp, ztrainp, kyinvp, hypp, xtrain, ztrain, kyinv, hyp, family="product", p=0.5):
    """Split map, literal: applymap_tok python/05_tokamak/Split_SympGPR/func.py:184-219 (array layout of
    python/05_tokamak/Split_SympGPR/main.py:91-112).  Whole turns only (`while i < nm - nphmap`); loss test at the
    new angle with phi = 2 pi/nphmap mod(i+1, nphmap) (compute_r ignores it, fieldlines.f90:34-47)."""
    q0 = np.asarray(q0, float); p0 = np.asarray(p0, float)
    E = len(q0)
    pmap = np.zeros((nm, E)); qmap = np.zeros((nm, E))
    pmap[0] = p0; qmap[0] = q0
    Np, Nt = xtrainp.shape[0] // 2, xtrain.shape[0] // 2
    i = 0
    while i < nm - nphmap:
        for m in range(nphmap):
            for k in range(E):
                if np.isnan(pmap[i, k]):
                    pmap[i + 1, k] = np.nan
                else:
                    pmap[i + 1, k] = calcp(qmap[i, k], pmap[i, k], hyp[m], hypp[m], xtrainp[:Np, m], xtrainp[Np:, m],
                                           ztrainp[:, m], kyinvp[m], xtrain[:Nt, m], xtrain[Nt:, m], ztrain[:, m], kyinv[m],
                                           family, p)
            for k in range(E):
                if np.isnan(pmap[i + 1, k]):
                    qmap[i + 1, k] = np.nan
                else:
                    dq = calcq(qmap[i, k], pmap[i + 1, k], xtrain[:Nt, m], xtrain[Nt:, m], hyp[m], kyinv[m], ztrain[:, m],
                               family, p)
                    qmap[i + 1, k] = np.mod(dq + qmap[i, k], TWO_PI)
                    ph = TWO_PI / nphmap * np.mod(i + 1, nphmap)
                    zk = np.array([pmap[i + 1, k] * 1e-2, qmap[i + 1, k], ph])
                    if compute_r(zk, 0.3) > 0.5 or pmap[i + 1, k] < 0.0:
                        pmap[i + 1, k] = np.nan
                        qmap[i + 1, k] = np.nan
            i = i + 1
    return qmap, pmap


def nll_expl(hyp, x, y, n, ind, p=0.5):
    """nll_expl python/04_standard_map/func.py:126-141: the sum kernel's matrix is block diagonal, so lx is
    fitted on the (q,q) block with the first half of the observations (ind = 0) and ly on the (P,P) block
    with the second half (ind = 1).  hyp = [l, sig, sig2n]; x = [q; P] (n values), y = the matching n/2
    observations.  (The reference fills the whole n x n matrix with the other length-scale set to 0 and then
    slices; only the block it keeps is finite, and only that block is formed here.)"""
    N = n // 2
    xq, xP = np.asarray(x[:N], float), np.asarray(x[N:2 * N], float)
    l, sig, sig2n = float(hyp[0]), float(hyp[1]), float(hyp[2])
    f = _fam("sum")
    fn = f.d2kdxdx0_num if ind == 0 else f.d2kdydy0_num
    # rows: x point (b), cols: x0 point (a) -- build_K_expl python/04_standard_map/func.py:104-124
    K = sig * (fn(xq[None, :], xP[None, :], xq[:, None], xP[:, None], l, l) + np.zeros((N, N)))
    Ky = K + abs(sig2n) * np.eye(N)
    L = scipy.linalg.cholesky(Ky, lower=True)
    alpha = solve_cholesky(L, np.asarray(y, float))
    return float(0.5 * np.asarray(y, float).dot(alpha) + np.sum(np.log(np.diag(L))))


# ------------------------------------------------- X1: 2-DOF 4x4-block kernel (NOT in the reference; parity unpinned)
def build_k4(x, x0, hyp):
    """Direct generalisation of build_K (sympgpr.f90:12-38) to F(q1, q2, P1, P2) with the SE kernel of
    python/03_henon_heiles/init_func.py:24-28 in all four variables, l = (lq, lq, lP, lP):
    K[aN+i, bN0+j] = sig (delta_ab/l_a^2 - D_a D_b/(l_a^2 l_b^2)) exp(-sum_c D_c^2/(2 l_c^2)), D = u(x_i) - u(x0_j).
    x = [q1; q2; P1; P2].  Validated by finite differences of the kernel and by its 2x2 sub-blocks
    reproducing the SE x SE matrix of the reference (tests/test_oracle.py)."""
    x = np.asarray(x, float); x0 = np.asarray(x0, float)
    N, N0 = len(x) // 4, len(x0) // 4
    l = np.array([hyp[0], hyp[0], hyp[1], hyp[1]], float)
    U = x.reshape(4, N); U0 = x0.reshape(4, N0)
    D = U[:, :, None] - U0[:, None, :]                          # (4, N, N0)
    E = np.exp(-0.5 * np.sum(D**2 / l[:, None, None]**2, axis=0))
    K = np.empty((4 * N, 4 * N0), order="F")
    for a in range(4):
        for b in range(4):
            blk = -D[a] * D[b] / (l[a]**2 * l[b]**2)
            if a == b:
                blk = blk + 1.0 / l[a]**2
            K[a * N:(a + 1) * N, b * N0:(b + 1) * N0] = hyp[2] * blk * E
    return K


def nll_grad4(hyp, x, y, n, with_sig=False):
    """NLL and its gradient w.r.t. (lq, lP[, sig]) for the 2-DOF kernel, the way nll_grad does it for the 2x2 case
    (python/02_pert_pendulum/func.py:148-162: -0.5 a' dK a + 0.5 tr(Kyinv dK)); dK by central differences of
    build_k4 in the length scales is NOT used -- the derivative blocks are written out analytically here and
    checked against differences in the tests."""
    x = np.asarray(x, float); y = np.asarray(y, float)
    N = n // 4
    lq, lP, sig, noise = float(hyp[0]), float(hyp[1]), float(hyp[2]), abs(float(hyp[3]))
    K = build_k4(x, x, [lq, lP, sig])
    Ky = K + noise * np.eye(n)
    L = scipy.linalg.cholesky(Ky, lower=True)
    alpha = solve_cholesky(L, y)
    val = float(0.5 * y.dot(alpha) + np.sum(np.log(np.diag(L))))
    Kinv = scipy.linalg.cho_solve((L, True), np.eye(n))
    l = np.array([lq, lq, lP, lP])
    U = x.reshape(4, N)
    D = U[:, :, None] - U[:, None, :]
    E = np.exp(-0.5 * np.sum(D**2 / l[:, None, None]**2, axis=0))
    grads = []
    for grp, lg in (((0, 1), lq), ((2, 3), lP)):
        S = sum(D[c]**2 for c in grp) / lg**3
        dK = np.empty((n, n))
        for a in range(4):
            for b in range(4):
                ga, gb = 1.0 / l[a]**2, 1.0 / l[b]**2
                ww = D[a] * D[b] * ga * gb
                base = (ga if a == b else 0.0) - ww
                d = base * S + 2.0 * ((a in grp) + (b in grp)) * ww / lg
                if a == b and a in grp:
                    d = d - 2.0 * ga / lg
                dK[a * N:(a + 1) * N, b * N:(b + 1) * N] = sig * d * E
        grads.append(-0.5 * alpha.dot(dK.dot(alpha)) + 0.5 * np.sum(Kinv * dK))
    if with_sig:
        grads.append(-0.5 * alpha.dot(K.dot(alpha)) / sig + 0.5 * np.sum(Kinv * K) / sig)
    return val, np.array(grads)


def grad_f4(xs, xtrain, alpha, hyp):
    """grad F at the points xs (4, E) of the 2-DOF generating function: rows of build_k4(xs, xtrain) times alpha."""
    xs = np.asarray(xs, float)
    E = xs.shape[1]
    K = build_k4(xs.reshape(-1), xtrain, hyp)
    return (K @ np.asarray(alpha, float)).reshape(4, E)


def applymap4(nm, q0, p0, hyp, xtrain, alpha):
    """2-DOF map loop, the twin of csrc/map.cu map4_kernel (no reference code: SURVEY 8a row X1).  One step solves
    P = p - grad_q F(q, P) by Newton started at P = p with the Jacobian I + d grad_q F / dP taken by central
    differences of grad_f4 here (the kernel uses the closed form), then Q = q + grad_P F(q, P).  Stops as the kernel
    does: step <= 1e-9 max(|P|, 1).  Returns qmap, pmap (nm, 2, E)."""
    q0 = np.asarray(q0, float); p0 = np.asarray(p0, float)
    E = q0.shape[1]
    qmap = np.zeros((nm, 2, E)); pmap = np.zeros((nm, 2, E))
    qmap[0], pmap[0] = q0, p0
    h = 1e-6
    for i in range(nm - 1):
        q, p = qmap[i], pmap[i]
        P = p.copy()
        done = np.zeros(E, bool)
        for it in range(40):
            F = grad_f4(np.vstack((q, P)), xtrain, alpha, hyp)
            J = np.zeros((2, 2, E))
            for c in range(2):
                dP = np.zeros((2, E)); dP[c] = h
                Fp = grad_f4(np.vstack((q, P + dP)), xtrain, alpha, hyp)
                Fm = grad_f4(np.vstack((q, P - dP)), xtrain, alpha, hyp)
                J[:, c] = (Fp[:2] - Fm[:2]) / (2 * h)
            r = P - p + F[:2]
            j00, j01, j10, j11 = 1 + J[0, 0], J[0, 1], J[1, 0], 1 + J[1, 1]
            det = j00 * j11 - j01 * j10
            d0 = (j11 * r[0] - j01 * r[1]) / det
            d1 = (j00 * r[1] - j10 * r[0]) / det
            upd = ~done
            P[0, upd] -= d0[upd]; P[1, upd] -= d1[upd]
            ad = np.maximum(np.abs(d0), np.abs(d1))
            done |= ad <= 1e-9 * np.maximum(np.abs(P).max(axis=0), 1.0)
            if done.all():
                break
        F = grad_f4(np.vstack((q, P)), xtrain, alpha, hyp)
        qmap[i + 1] = q + F[2:]
        pmap[i + 1] = P
    return qmap, pmap


def henon_like_training(N, seed=3):
    """Synthetic 2-DOF symplectic map for the X1 tests: one kick-drift step of a Henon-Heiles-like potential,
    P = p - dt dV/dq(q), Q = q + dt P, V = (q1^2 + q2^2)/2 + q1^2 q2 - q2^3/3; Halton points in [-0.4, 0.4]^4."""
    dt = 0.3
    q1 = -0.4 + 0.8 * halton(N, 2, start=seed); q2 = -0.4 + 0.8 * halton(N, 3, start=seed)
    p1 = -0.4 + 0.8 * halton(N, 5, start=seed); p2 = -0.4 + 0.8 * halton(N, 7, start=seed)
    P1 = p1 - dt * (q1 + 2 * q1 * q2); P2 = p2 - dt * (q2 + q1**2 - q2**2)
    Q1 = q1 + dt * P1; Q2 = q2 + dt * P2
    x = np.concatenate((q1, q2, P1, P2))
    z = np.concatenate((p1 - P1, p2 - P2, Q1 - q1, Q2 - q2))
    return x, z


# ------------------------------------------------------------------- M5 tokamak
def compute_r(z, rstart=0.3):
    """fieldlines.f90:94-107 with f_r :82-91, Ath :34-39, dAthdr :42-47
    (B0 = R0 = 1): exactly 20 Newton iterations."""
    r = rstart
    for _ in range(20):
        yv = z[0] - 1.0 * (r**2 / 2.0 - r**3 / (3.0 * 1.0) * np.cos(z[1]))
        dy = -(1.0 * (r - r**2 / 1.0 * np.cos(z[1])))
        r = r - yv / dy
    return r


# --------------------------------------------------------------- map loops (M4)
MAP_PENDULUM = 0     # q = mod(q + dq, 2pi)                    functions/func.py:216-237
MAP_HENON = 1        # q = q + dq                              functions/func.py:239-260
MAP_STANDARD = 2     # q = mod(q+dq,2pi); p = mod(P,2pi)       04_standard_map/func.py:218-254
MAP_TOKAMAK = 3      # pendulum wrap + loss test               05_tokamak/SympGPR/func.py:182-211


def applymap(kind, nm, q0, p0, hyp, hypp, xtrainp, ytrainp, ztrainp, kyinvp,
             xtrain, ytrain, ztrain, kyinv, family="product", p=0.5):
    """Literal ensemble loop; returns (qmap, pmap[, pdiff]) of shape (nm, Ntest)."""
    q0 = np.asarray(q0, float); p0 = np.asarray(p0, float)
    E = len(q0)
    pmap = np.zeros((nm, E)); qmap = np.zeros((nm, E)); pdiff = np.zeros((nm, E))
    pmap[0] = p0; qmap[0] = q0; pdiff[0] = p0
    for i in range(nm - 1):
        for k in range(E):
            if kind == MAP_TOKAMAK and np.isnan(pmap[i, k]):
                pmap[i + 1, k] = np.nan
                continue
            pmap[i + 1, k] = calcp(qmap[i, k], pmap[i, k], hyp, hypp, xtrainp, ytrainp, ztrainp,
                                   kyinvp, xtrain, ytrain, ztrain, kyinv, family, p)
            if kind == MAP_STANDARD:
                pdiff[i + 1, k] = pdiff[i, k] + (pmap[i + 1, k] - pmap[i, k])
                pmap[i + 1, k] = np.mod(pmap[i + 1, k], TWO_PI)
            if kind == MAP_TOKAMAK:
                zk = np.array([pmap[i + 1, k] * 1e-2, qmap[i, k], 0.0])
                if compute_r(zk, 0.3) > 0.5 or pmap[i + 1, k] < 0.0:
                    pmap[i + 1, k] = np.nan
        for k in range(E):
            if np.isnan(pmap[i + 1, k]):
                qmap[i + 1, k] = np.nan
            else:
                dq = calcq(qmap[i, k], pmap[i + 1, k], xtrain, ytrain, hyp, kyinv, ztrain, family, p)
                if kind == MAP_HENON:
                    qmap[i + 1, k] = dq + qmap[i, k]
                else:
                    qmap[i + 1, k] = np.mod(dq + qmap[i, k], TWO_PI)
    if kind == MAP_STANDARD:
        return qmap, pmap, pdiff
    return qmap, pmap


# ------------------------------------------------------ quality metrics
def energy_pendulum(q, p, U0):
    """energy(x, U0) python/01_pendulum/implicit/func.py:116-117."""
    return np.asarray(p)**2 / 2 + U0 * (1 - np.cos(np.asarray(q) + np.pi))


def aph(r, th, ph, eps, m, n, phase):
    """fieldlines.f90:58-64 (B0 = iota0 = 1, a = 0.5)."""
    return -(r**2 / 2.0 - r**4 / (4.0 * 0.5**2)) * (1.0 + eps * np.cos(m * th + n * ph + phase))


def energy_tok(qmap, pmap, eps, m, phase):
    """energy(qmap, pmap, nm) python/05_tokamak/Split_SympGPR/func.py:234-246 (ph = 0): H (N, nm)."""
    nm, N = qmap.shape
    H = np.zeros((N, nm))
    for i in range(N):
        for k in range(nm):
            r = compute_r(np.array([pmap[k, i] * 1e-2, qmap[k, i], 0.0]), 0.3)
            H[i, k] = -aph(r, qmap[k, i], 0.0, eps, m, 0, phase)
    return H


def quality_eosc(H):
    """Energy oscillation of `quality` python/functions/func.py:268-271: H (nm, Ntest) -> std/mean per orbit."""
    H = np.asarray(H)
    return np.array([np.std(H[:, k]) / np.mean(H[:, k]) for k in range(H.shape[1])])


def standard_map_iterate(k, nm, N, X0):
    """StandardMap / StandardMapIterate python/04_standard_map/main.py:27-39."""
    f = np.zeros((2, N, nm))
    f[:, :, 0] = X0
    for i in range(N):
        for l in range(nm - 1):
            J = f[1, i, l] + k * np.sin(f[0, i, l])
            f[:, i, l + 1] = [f[0, i, l] + J, J]
    return f


# -------------------------------------------------- synthetic workload
def halton(n, base, start=1):
    """Unscrambled van der Corput sequence, indices start..start+n-1."""
    out = np.zeros(n)
    for k in range(n):
        i, f, r = start + k, 1.0, 0.0
        while i > 0:
            f /= base
            r += f * (i % base)
            i //= base
        out[k] = r
    return out


def standard_map_training(N, kchaos=0.9):
    """python/04_standard_map/main.py:27-30,43-59,89-92 with k -> kchaos and
    Halton(2,3) training points on [0,2pi)^2."""
    q = halton(N, 2) * TWO_PI
    pp = halton(N, 3) * TWO_PI
    P = pp + kchaos * np.sin(q)
    Q = q + P
    xtrain = np.hstack((q, P))
    ztrain = np.concatenate((pp - P, Q - q))
    xtrainp = np.hstack((q, pp))
    ztrainp = P - pp
    sig = 2 * np.amax(np.abs(ztrain))**2
    sigp = 2 * np.amax(np.abs(ztrainp))**2
    return dict(q=q, p=pp, Q=Q, P=P, xtrain=xtrain, ztrain=ztrain, xtrainp=xtrainp,
                ztrainp=ztrainp, sig=sig, sigp=sigp)


def timing_hyp(N, sig, sig2n=1e-8):
    """lx = ly = 0.5*2pi/sqrt(N): keeps cond(Ky) ~ 1e4."""
    l = 0.5 * TWO_PI / np.sqrt(N)
    return np.array([l, l, sig, sig2n])
