"""Golden NLL + gradient at BASELINE's FULL size (N = 16384 training pairs, n = 32768) from the CPU oracle.

The oracle's own nll_grad (oracle/oracle.py:195-215) would hold five dense n x n matrices; this script evaluates the
same formulas with the same oracle functions block by block (fill by column blocks, dK by row blocks) so that it
fits in ~20 GB:  K = build_k_vec(...) + |sig2n| I;  L = chol(K) (blocked, LAPACK on blocks);  alpha = L^-T L^-1 z;
NLL = z.alpha/2 + sum log diag L;  W = L^-T L^-1 (triangular solves by column blocks);  grad_t = -1/2 sum_ij (alpha_i alpha_j - W_ij) dK_t[i,j]
(python/02_pert_pendulum/func.py:148-162, python/05_tokamak/SympGPR/func.py:143-150).
Checked against oracle.nll_grad itself at N = 1024 before the full-size run.

    python tests/golden/make_golden_fullsize.py [N]      -> tests/golden/fullsize_nll_N<N>.json   (~10 min, 8 cores)
"""
import json, os, sys, time
import numpy as np
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O


def blocked_cholesky(K, nb):
    """Lower Cholesky factor in place, right-looking by block columns with LAPACK/BLAS calls on blocks only
    (OpenBLAS's own dpotrf segfaults at n = 32768 in this image).  The strict upper triangle is left untouched."""
    n = K.shape[0]
    for k in range(0, n, nb):
        e = min(n, k + nb)
        K[k:e, k:e] = scipy.linalg.cholesky(K[k:e, k:e], lower=True, check_finite=False)
        if e < n:
            Lkk = K[k:e, k:e]
            K[e:, k:e] = scipy.linalg.solve_triangular(Lkk, K[e:, k:e].T, lower=True, check_finite=False).T
            P = K[e:, k:e]
            for j in range(e, n, nb):                       # lower block columns of the trailing matrix
                je = min(n, j + nb)
                K[j:, j:je] -= P[j - e:, :] @ P[j - e:je - e, :].T
    return K


def nll_grad_blocked(hyp, x, z, n, blk=1024, nb=4096):
    N = n // 2
    xs, ys = x[:N], x[N:]
    K = np.empty((n, n), order="F")
    for b0 in range(0, N, blk):
        b = slice(b0, min(N, b0 + blk))
        nb = b.stop - b.start
        Kb = O.build_k_vec(xs, ys, xs[b], ys[b], hyp[:3])
        K[:, b0:b0 + nb] = Kb[:, :nb]
        K[:, N + b0:N + b0 + nb] = Kb[:, nb:]
    K[np.diag_indices(n)] += abs(hyp[3])
    L = blocked_cholesky(K, nb)
    alpha = scipy.linalg.solve_triangular(L, scipy.linalg.solve_triangular(L, z, lower=True, check_finite=False),
                                          lower=True, trans="T", check_finite=False)
    val = 0.5 * z.dot(alpha) + np.sum(np.log(L.diagonal()))
    # inverse by column blocks: W[:, b] = L^-T L^-1 I[:, b]
    W = np.empty((n, n), order="F")
    for b0 in range(0, n, nb):
        be = min(n, b0 + nb)
        rhs = np.zeros((n, be - b0), order="F")
        rhs[np.arange(b0, be), np.arange(be - b0)] = 1.0
        y = scipy.linalg.solve_triangular(L, rhs, lower=True, check_finite=False, overwrite_b=True)
        W[:, b0:be] = scipy.linalg.solve_triangular(L, y, lower=True, trans="T", check_finite=False, overwrite_b=True)
    del L
    grad = np.zeros(2)
    for b0 in range(0, N, blk):
        b = np.arange(b0, min(N, b0 + blk))
        idx = np.concatenate((b, N + b))
        dK = O.build_dk(x, np.concatenate((xs[b], ys[b])), hyp[:3])        # rows: block points, cols: all points
        Wr = W[idx, :]                                                         # rows of the (full, symmetric) inverse
        M = np.outer(alpha[idx], alpha) - Wr
        for t in range(2):
            grad[t] += -0.5 * np.sum(M * dK[t])
    return val, grad


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    # self-check of the blocked evaluation against the oracle proper
    Nc = 1024
    d = O.standard_map_training(Nc)
    hyp = O.timing_hyp(Nc, d["sig"], 1e-8)
    v0, g0 = O.nll_grad(hyp, d["xtrain"], d["ztrain"], 2 * Nc)
    v1, g1 = nll_grad_blocked(hyp, d["xtrain"], d["ztrain"], 2 * Nc, blk=300, nb=384)
    assert abs(v1 - v0) <= 1e-12 * abs(v0) and np.allclose(g1, g0, rtol=1e-11), (v0, v1, g0, g1)
    d = O.standard_map_training(N)
    hyp = O.timing_hyp(N, d["sig"], 1e-8)
    t0 = time.time()
    v, g = nll_grad_blocked(hyp, d["xtrain"], d["ztrain"], 2 * N)
    out = {"N": N, "n": 2 * N, "hyp": list(map(float, hyp)), "nll": float(v), "grad": [float(g[0]), float(g[1])],
           "workload": "oracle.standard_map_training(N), oracle.timing_hyp(N, sig, 1e-8)", "seconds": time.time() - t0,
           "selfcheck_N1024": {"nll_rel": abs(v1 - v0) / abs(v0), "grad_rel": float(np.abs(g1 - g0).max() / np.abs(g0).max())}}
    path = os.path.join(ROOT, "tests", "golden", f"fullsize_nll_N{N}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
