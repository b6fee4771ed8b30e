// Cholesky factor and its inverse with the level-3 work on the INT8 tensor pipe (opt-in, sgp_set_ozaki_ex): the route past
// the DMMA ceiling for the two stages of an NLL+gradient evaluation that ozaki_lauum() does not cover.
//
// Replaces, together with ozaki_lauum, the same LAPACK calls as chol.cu (scipy.linalg.cholesky
// python/05_tokamak/SympGPR/func.py:147, np.linalg.inv python/02_pert_pendulum/func.py:152).
//
// The evaluation needs X = L^-1 (alpha = X^T X z, K^-1 = X^T X, log det = -sum log X(i,i)) but never L itself, so factor and
// inverse are ONE recursion over the lower triangle, in place:
//     node(A):  A = [A11 .; A21 A22]
//       1. node(A11)                      -> X11 = L11^-1 in place of A11
//       2. A21 <- A21 X11^T               =  L21                       (B = X11, lower:  k <  64 (tn + 1))
//       3. A22 <- A22 - L21 L21^T                                      (lower tiles only)
//       4. A21 <- L21 X11                                              (B = X11^T:       k >= 64 tn)
//       5. node(A22)                      -> X22 in place of A22
//       6. A21 <- -X22 (L21 X11)          =  X21                       (A = X22, lower:  k < 128 (tm + 1))
// Blocks of <= leaf_n rows are factored and inverted by the DMMA kernels (potrf_ll, grouped trtri: the chain of diagonal tiles
// is latency-bound, not GEMM-bound); every product above them is one ozaki_gemm_sliced() -- 4 h^3 flop per node with halves of
// h rows, 98 % of the 2 n^3 / 3 in nodes with h >= 4096 at n = 32768.  A sliced product reads only the INT8 slices, so its
// FP64 output may overwrite the operand it was sliced from: no scratch matrix, and L21's slices serve steps 3 and 4.
#include "chol.cuh"
#include "ozaki.cuh"

namespace sgp {

namespace {

#define AT(A, lda, rt, ct) ((A) + (long)(rt) * TILE + (long)(ct) * TILE * (lda))

struct FactInv {
    Ctx& c;
    int ns, leaf_t;
    double* A; long lda;
    double *Dinv, *logparts, *T;
    int* info;
    void *bufA, *bufB;
    double *z, *w, *part;      // factor-only variant: right-hand side (updated in place), solution of L w = z, gemv scratch
};

// inverse = true: the block is replaced by X = L^-1.  inverse = false (factor-only variant, for evaluations that need the value
// alone): X11 is still formed (step 2 needs it) but A21 keeps L21, steps 4 and 6 are skipped and the (2,2) block recurses in
// the same form -- n^3/3 + n^3/24 + ... = 0.38 n^3 flop instead of 0.67 n^3; the forward substitution L w = z rides along:
// w1 = X11 z1, z2 -= L21 w1 before the (2,2) block, and a leaf solves its own part inside potrf_ll.
int node(FactInv& f, int j0, int mt, bool inverse)
{
    double* A11 = AT(f.A, f.lda, j0, j0);
    if (mt <= f.leaf_t) {
        const long nb = (long)mt * TILE;
        double* Dinv = f.Dinv + (long)j0 * TILE * TILE;
        if (!inverse) return potrf(f.c, A11, nb, f.lda, Dinv, f.logparts + j0, f.info, f.z + (long)j0 * TILE, f.w + (long)j0 * TILE);
        SGP_TRY(potrf(f.c, A11, nb, f.lda, Dinv, f.logparts + j0, f.info));
        return trtri(f.c, A11, nb, f.lda, Dinv, f.T);
    }
    const int m1 = mt / 2, m2 = mt - m1;
    const long h1 = (long)m1 * TILE, h2 = (long)m2 * TILE;
    double* A21 = AT(f.A, f.lda, j0 + m1, j0);
    double* A22 = AT(f.A, f.lda, j0 + m1, j0 + m1);
    Ctx& c = f.c;
    const int ns = f.ns;
    SGP_TRY(node(f, j0, m1, true));
    // 2. L21 = A21 X11^T
    OzSliced SA = ozaki_carve(f.bufA, h2, h1, ns), SB = ozaki_carve(f.bufB, h1, h1, ns);
    SGP_TRY(ozaki_slice(c, ns, A21, f.lda, h2, h1, OZ_MN, 0, SA));
    SGP_TRY(ozaki_slice(c, ns, A11, f.lda, h1, h1, OZ_MN, 1, SB));
    SGP_TRY(ozaki_gemm_sliced(c, ns, SA, SB, h2, h1, 1.0, 0.0, A21, f.lda, OZ_KHI_TN, 0));
    // 3. A22 -= L21 L21^T
    SGP_TRY(ozaki_slice(c, ns, A21, f.lda, h2, h1, OZ_MN, 0, SA));
    SGP_TRY(ozaki_gemm_sliced(c, ns, SA, SA, h2, h2, -1.0, 1.0, A22, f.lda, 0, 1));
    if (!inverse) {
        double* z1 = f.z + (long)j0 * TILE;
        double* w1 = f.w + (long)j0 * TILE;
        SGP_TRY(gemv_blocked(c, A11, f.lda, h1, h1, 1, z1, w1, 0, f.part));                      // w1 = X11 z1
        SGP_TRY(gemv_blocked(c, A21, f.lda, h2, h1, 0, w1, z1 + h1, 1, f.part));                 // z2 -= L21 w1
        return node(f, j0 + m1, m2, false);
    }
    // 4. A21 = L21 X11
    SGP_TRY(ozaki_slice(c, ns, A11, f.lda, h1, h1, OZ_K, 2, SB));
    SGP_TRY(ozaki_gemm_sliced(c, ns, SA, SB, h2, h1, 1.0, 0.0, A21, f.lda, OZ_KLO_TN, 0));
    // 5.
    SGP_TRY(node(f, j0 + m1, m2, true));
    // 6. X21 = -X22 (L21 X11)
    SA = ozaki_carve(f.bufA, h2, h2, ns);
    SB = ozaki_carve(f.bufB, h1, h2, ns);
    SGP_TRY(ozaki_slice(c, ns, A22, f.lda, h2, h2, OZ_MN, 1, SA));
    SGP_TRY(ozaki_slice(c, ns, A21, f.lda, h1, h2, OZ_K, 0, SB));
    return ozaki_gemm_sliced(c, ns, SA, SB, h2, h1, -1.0, 0.0, A21, f.lda, OZ_KHI_TM, 0);
}

}  // namespace

size_t ozaki_factinv_workspace_bytes(long n_pad, int ns)
{
    const long nt = n_pad / TILE, hmax = (nt - nt / 2) * TILE;
    return 2 * (ozaki_sliced_bytes(hmax, hmax, ns) + 256);
}

static int oz_chol_check(const char* who, int ns, long leaf_n, long n_pad, long lda, size_t work_bytes)
{
    if (n_pad <= 0 || n_pad % TILE || (lda & 1) || leaf_n < TILE) { set_error("%s: bad arguments", who); return ST_BADARG; }
    if (ns < 4 || ns > 8) { set_error("%s: 4..8 slices, got %d", who, ns); return ST_BADARG; }
    if (work_bytes < ozaki_factinv_workspace_bytes(n_pad, ns)) { set_error("%s: workspace too small", who); return ST_BADARG; }
    return ST_OK;
}

int ozaki_factinv(Ctx& c, int ns, long leaf_n, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, double* T,
                  void* work, size_t work_bytes)
{
    SGP_TRY(oz_chol_check("ozaki_factinv", ns, leaf_n, n_pad, lda, work_bytes));
    const size_t half = ozaki_factinv_workspace_bytes(n_pad, ns) / 2;
    FactInv f{c, ns, (int)(leaf_n / TILE), A, lda, Dinv, logparts, T, info, work, (char*)work + half, nullptr, nullptr, nullptr};
    return node(f, 0, (int)(n_pad / TILE), true);
}

size_t ozaki_factor_solve_scratch_doubles(long n_pad)
{
    const long nt = n_pad / TILE, h = (nt - nt / 2) * TILE;
    return gemv_scratch_doubles(h, h) + 16;
}

int ozaki_factor_solve(Ctx& c, int ns, long leaf_n, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, double* T,
                       double* z, double* w, double* part, void* work, size_t work_bytes)
{
    SGP_TRY(oz_chol_check("ozaki_factor_solve", ns, leaf_n, n_pad, lda, work_bytes));
    if (!z || !w || !part) { set_error("ozaki_factor_solve: bad arguments"); return ST_BADARG; }
    const size_t half = ozaki_factinv_workspace_bytes(n_pad, ns) / 2;
    FactInv f{c, ns, (int)(leaf_n / TILE), A, lda, Dinv, logparts, T, info, work, (char*)work + half, z, w, part};
    return node(f, 0, (int)(n_pad / TILE), false);
}

}  // namespace sgp
