// Dense FP64 building blocks (see chol.cu).
#pragma once
#include "common.cuh"
#include "dmma_gemm.cuh"
#include "dmma_gemm_ws.cuh"

namespace sgp {

// Cholesky factor in place; with y and w given also the forward substitution L w = y (y is left alone)
int potrf(Ctx& c, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, double* y = nullptr, double* w = nullptr);
// one persistent cooperative kernel (potrf_ll.cu); flags: potrf_ll_flag_bytes(n_pad) bytes of device scratch
int potrf_ll(Ctx& c, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, int* flags, const double* y,
             double* w);
size_t potrf_ll_flag_bytes(long n_pad);
// blocked backward substitution with the tile inverses: L^T alpha = w (w destroyed); the forward half is fused into potrf_ll
int trsv_bwd(Ctx& c, const double* L, long n_pad, long lda, const double* Dinv, double* w, double* alpha);
// alpha = X^T w for the explicit lower-triangular inverse factor X (entries above the diagonal are not read)
int gemv_t_lower(Ctx& c, const double* X, long n_pad, long ldx, const double* w, double* alpha);
// w = X z for the explicit lower-triangular X; part: trmv_lower_scratch_doubles(n_pad) doubles of device scratch
int trmv_lower(Ctx& c, const double* X, long n_pad, long ldx, const double* z, double* w, double* part);
size_t trmv_lower_scratch_doubles(long n_pad);
// y = A x (mode 0) / y -= A x (mode 1), A rows x cols column-major (tri: square lower triangular); part: gemv_scratch_doubles
int gemv_blocked(Ctx& c, const double* A, long ld, long rows, long cols, int tri, const double* x, double* y, int mode, double* part);
size_t gemv_scratch_doubles(long rows, long cols);
// alpha = W y for the symmetric W given by its lower triangle (tiles on and below the diagonal complete)
int symv_lower(Ctx& c, const double* W, long n_pad, long ldw, const double* y, double* alpha);
int trtri(Ctx& c, double* A, long n_pad, long lda, const double* Dinv, double* T);
size_t trtri_workspace_doubles(long n_pad);
int lauum(Ctx& c, const double* X, long n_pad, long lda, double* W, long ldw);

int ref_gemm(Ctx& c, int al, int bl, const GemmArgs& g, double* out);
int dmma_gemm(Ctx& c, int al, int bl, const GemmArgs& g);

}  // namespace sgp
