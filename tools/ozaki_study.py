"""Round-2 groundwork (CPU only): how many INT8 slices would an Ozaki-split GEMM need for the trailing updates of the
Cholesky factorisation to keep the NLL within BASELINE's 1e-9?  NumPy emulation: operands are scaled per row (A) /
per column (B) to [-1, 1), cut into `s` slices of `bits` bits, slice products are exact integer GEMMs (here: int64), and
C = sum_{p+q <= s+1} A_p B_q 2^(-bits (p+q)) with the products of equal p+q added exactly before conversion -- the
scheme an INT8 tcgen05 path would run.  Only the trailing SYRK/GEMM updates of a blocked
right-looking Cholesky are emulated (they are all of the n^3 work); panels and the diagonal blocks stay in FP64.

    python tools/ozaki_study.py [N]      -> table on stdout
"""
import sys
import numpy as np
import scipy.linalg

sys.path.insert(0, ".")
from oracle import oracle as O


def split(M, s, bits, axis):
    """M scaled along `axis` (rows of A: axis=1 -> one power-of-two scale per row) into s integer slices of `bits` bits."""
    mx = np.max(np.abs(M), axis=axis, keepdims=True)
    e = np.ceil(np.log2(np.where(mx > 0, mx, 1.0))) + 1
    X = M / 2.0**e                                   # |X| < 0.5
    slices = []
    for _ in range(s):
        X = X * 2.0**bits
        q = np.round(X)                              # |q| <= 2^(bits-1): fits a signed (bits+1)-bit integer
        slices.append(q.astype(np.int64))
        X = X - q
    return slices, e


def split256(M, s):
    """The scheme csrc/ozaki.cu runs since round-2 state "j": s balanced radix-256 digits of the row-scaled mantissa
    (v = rint(x 2^(8 s - 1 - e)); signed bytes from the bottom with a carry); returned most significant first."""
    mx = np.max(np.abs(M), axis=1, keepdims=True)
    f, e = np.frexp(np.where(mx > 0, mx, 1.0))
    e = e + (f >= 0.995)
    v = np.rint(np.ldexp(M, (8 * s - 1 - e).astype(np.int64) if s <= 6 else (8 * s - 1 - e).astype(np.int64)))
    v = v.astype(object) if s > 7 else v.astype(np.int64)      # 63 bits: beyond float64-exact int64 conversion only in the last bits
    v = np.array(v, dtype=object)
    digits = []
    for _ in range(s):
        dgt = np.vectorize(lambda t: ((int(t) + 128) % 256) - 128, otypes=[object])(v)
        digits.append(np.array(dgt, dtype=np.int64))
        v = (v - dgt) // 256
    return digits[::-1], e


def ozaki256_gemm_nt(A, B, s):
    As, ea = split256(A, s)
    Bs, eb = split256(B, s)
    C = np.zeros((A.shape[0], B.shape[0]))
    for dcls in range(s - 1, -1, -1):                # smallest contributions first
        acc = np.zeros((A.shape[0], B.shape[0]), dtype=np.int64)
        for p in range(dcls + 1):
            acc += As[p] @ Bs[dcls - p].T
        C += acc.astype(np.float64) * 2.0**(-8 * dcls)
    return C * 2.0**(ea - 7) * 2.0**(eb - 7).T


def ozaki_gemm_nt(A, B, s, bits):
    """A (m,k) @ B(n,k)^T with both operands cut into s slices."""
    As, ea = split(A, s, bits, 1)
    Bs, eb = split(B, s, bits, 1)
    C = np.zeros((A.shape[0], B.shape[0]))
    for g in range(2, s + 2):                        # p + q = g, 1-based slice indices
        acc = np.zeros((A.shape[0], B.shape[0]), dtype=np.int64)
        for p in range(1, g):
            q = g - p
            if p <= s and q <= s:
                acc += As[p - 1] @ Bs[q - 1].T       # exact: |entries| <= k 2^(2 bits - 2)
        C += acc.astype(np.float64) * 2.0**(-bits * g)
    return C * 2.0**ea * 2.0**eb.T


def blocked_cholesky(K, nb, gemm):
    n = K.shape[0]
    K = K.copy()
    for k in range(0, n, nb):
        e = min(n, k + nb)
        K[k:e, k:e] = scipy.linalg.cholesky(K[k:e, k:e], lower=True)
        if e < n:
            K[e:, k:e] = scipy.linalg.solve_triangular(K[k:e, k:e], K[e:, k:e].T, lower=True).T
            P = K[e:, k:e]
            K[e:, e:] -= gemm(P, P)
    return np.tril(K)


def nll_of(L, z):
    w = scipy.linalg.solve_triangular(L, z, lower=True)
    return 0.5 * w.dot(w) + np.sum(np.log(np.diag(L)))


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    d = O.standard_map_training(N)
    print(f"N = {N} (n = {2 * N}), block 128; relative NLL error of the Cholesky with Ozaki-split trailing updates vs FP64")
    print("model                      cond(Ky)   " + "  ".join(f"s={s},b={b}" for s, b in ((6, 6), (7, 6), (8, 6), (7, 7), (8, 7), (9, 7))))
    for name, fac in (("timing hyp (0.5 2pi/sqrtN)", 1.0), ("map model (1.0 2pi/sqrtN)", 2.0)):
        hyp = O.timing_hyp(N, d["sig"], 1e-8)
        hyp[:2] *= fac
        K = O.build_k_vec(d["xtrain"][:N], d["xtrain"][N:], d["xtrain"][:N], d["xtrain"][N:], hyp[:3]) + hyp[3] * np.eye(2 * N)
        z = d["ztrain"]
        ref = nll_of(blocked_cholesky(K, 128, lambda A, B: A @ B.T), z)
        cond = np.linalg.cond(K)
        errs = []
        for s, b in ((6, 6), (7, 6), (8, 6), (7, 7), (8, 7), (9, 7)):
            try:
                v = nll_of(blocked_cholesky(K, 128, lambda A, B: ozaki_gemm_nt(A, B, s, b)), z)
                errs.append(abs(v - ref) / abs(ref))
            except np.linalg.LinAlgError:
                errs.append(float("nan"))
        print(f"{name:26s} {cond:9.2e}   " + "  ".join(f"{e:8.1e}" for e in errs))
        errs = []
        for sdig in (5, 6, 7, 8):
            try:
                v = nll_of(blocked_cholesky(K, 128, lambda A, B: ozaki256_gemm_nt(A, B, sdig)), z)
                errs.append(abs(v - ref) / abs(ref))
            except np.linalg.LinAlgError:
                errs.append(float("nan"))
        print(f"{'  radix-256 digits 5,6,7,8':26s} {'':9s}   " + "  ".join(f"{e:8.1e}" for e in errs))
