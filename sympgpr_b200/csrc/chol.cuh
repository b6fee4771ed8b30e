// Dense FP64 building blocks (see chol.cu).
#pragma once
#include "common.cuh"
#include "dmma_gemm.cuh"
#include "dmma_gemm_ws.cuh"

namespace sgp {

int potrf(Ctx& c, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info);
// one persistent cooperative kernel (potrf_ll.cu); flags: potrf_ll_flag_bytes(n_pad) bytes of device scratch
int potrf_ll(Ctx& c, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, int* flags);
size_t potrf_ll_flag_bytes(long n_pad);
int potrs(Ctx& c, const double* L, long n_pad, long lda, const double* Dinv, double* y, double* w, double* alpha);
int trtri(Ctx& c, double* A, long n_pad, long lda, const double* Dinv, double* T);
size_t trtri_workspace_doubles(long n_pad);
int lauum(Ctx& c, const double* X, long n_pad, long lda, double* W, long ldw);

int ref_gemm(Ctx& c, int al, int bl, const GemmArgs& g, double* out);
int dmma_gemm(Ctx& c, int al, int bl, const GemmArgs& g);

}  // namespace sgp
