// Opt-in FP64 GEMM from the INT8 tensor pipe (tcgen05.mma kind::i8, Ozaki splitting); see ozaki.cu, ozaki_chol.cu.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace sgp {

// self test of the tcgen05 plumbing: one 128 x 64 x K INT8 product on the tensor pipe against a plain integer kernel;
// *mismatches = number of differing INT32 outputs (0 = pass); the two probes are element (5,3) of the reference / the result
int i8mma_selftest(Ctx& c, int K, int* mismatches, int* probe_ref, int* probe_got);

// ---- building blocks: an operand split once, used by any number of products ------------------------------------------
// One operand (R rows x K) as ns signed 7-bit slices of its row-scaled mantissas: sl[s][r][k] (k contiguous; Rp, Kp = R, K
// rounded up to 128), ex[r] = row exponents, rowmax = scratch of the split.  ozaki_carve() lays one out in caller memory of
// ozaki_sliced_bytes(R, K, ns) bytes.
struct OzSliced {
    int8_t* sl = nullptr;
    int* ex = nullptr;
    unsigned long long* rowmax = nullptr;
    long Rp = 0, Kp = 0;
};
size_t ozaki_sliced_bytes(long R, long K, int ns);
OzSliced ozaki_carve(void* buf, long R, long K, int ns);
// storage order of the FP64 source: OZ_MN element (r, k) at X[r + k ld]; OZ_K element (r, k) at X[k + r ld]
constexpr int OZ_MN = 0, OZ_K = 1;
// tri: 0 = every entry valid; 1 = only k <= r (OZ_MN view of a lower-triangular matrix); 2 = only k >= r (OZ_K view of one)
int ozaki_slice(Ctx& c, int ns, const double* X, long ld, long R, long K, int layout, int tri, const OzSliced& out);
// k-range of the output tile (tm: 128 rows of A, tn: 64 rows of B): lower / upper limits that triangular operands allow
constexpr int OZ_KLO_TM = 1;      // k >= 128 tm        (A(m, k) = 0 for k < m: A is the OZ_K view of a lower-triangular matrix)
constexpr int OZ_KLO_TN = 2;      // k >= 64 tn         (B likewise)
constexpr int OZ_KHI_TM = 4;      // k < 128 (tm + 1)   (A(m, k) = 0 for k > m: A lower triangular)
constexpr int OZ_KHI_TN = 8;      // k < 64 (tn + 1)    (B lower triangular)
// C (M x N, column-major, ldc) = alpha A B^T + beta C; lower != 0: only tiles that touch the lower triangle (64 tn <= 128 tm + 127)
int ozaki_gemm_sliced(Ctx& c, int ns, const OzSliced& A, const OzSliced& B, long M, long N, double alpha, double beta, double* C, long ldc,
                      int kmode, int lower);

// ---- whole products -----------------------------------------------------------------------------------------------------
// C (M x N, column-major) = alpha A B^T + beta C from INT8 slice products; A (M x K), B (N x K): element (r, k) at ptr[r + k ld];
// ns = 4..8 slices per operand; work: ozaki_workspace_bytes(M, N, K, ns) bytes of device scratch.  Enqueues on c.stream.
size_t ozaki_workspace_bytes(long M, long N, long K, int ns);
int ozaki_slice_operands(Ctx& c, int ns, long M, long N, long K, const double* A, long lda, const double* B, long ldb, void* work, size_t work_bytes);
int ozaki_gemm_presliced(Ctx& c, int ns, long M, long N, long K, double alpha, double beta, double* C, long ldc, void* work, size_t work_bytes);
int ozaki_gemm(Ctx& c, int ns, long M, long N, long K, double alpha, const double* A, long lda, const double* B, long ldb, double beta,
               double* C, long ldc, void* work, size_t work_bytes);

// lauum on the INT8 pipe: W = X^T X (lower tiles) for the lower-triangular inverse factor X; work: ozaki_lauum_workspace_bytes
size_t ozaki_lauum_workspace_bytes(long n_pad, int ns);
int ozaki_lauum(Ctx& c, int ns, const double* X, long n_pad, long ldx, double* W, long ldw, void* work, size_t work_bytes);

// Cholesky factor AND its inverse on the INT8 pipe (ozaki_chol.cu): the lower triangle of A (n_pad x n_pad, column-major, lda;
// n_pad a multiple of 128) is replaced by X = L^-1 with A = L L^T; logparts[t] = sum of log L(i,i) over tile t; *info as potrf.
// Blocks of at most leaf_n rows are factored and inverted by the DMMA kernels (potrf_ll, trtri), everything above them is
// four sliced products per node.  work: ozaki_factinv_workspace_bytes(n_pad, ns); Dinv, T: as for potrf / trtri of a leaf.
size_t ozaki_factinv_workspace_bytes(long n_pad, int ns);
int ozaki_factinv(Ctx& c, int ns, long leaf_n, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, double* T,
                  void* work, size_t work_bytes);

// Factor-only variant for evaluations that need the value alone (0.38 n^3 flop): the diagonal blocks of the recursion hold
// inverse factors, the off-diagonal blocks L itself, and the forward substitution L w = z rides along (z is updated in place;
// leaves solve their part inside potrf_ll).  part: ozaki_factor_solve_scratch_doubles(n_pad) doubles; work as for ozaki_factinv.
size_t ozaki_factor_solve_scratch_doubles(long n_pad);
int ozaki_factor_solve(Ctx& c, int ns, long leaf_n, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, double* T,
                       double* z, double* w, double* part, void* work, size_t work_bytes);

}  // namespace sgp
