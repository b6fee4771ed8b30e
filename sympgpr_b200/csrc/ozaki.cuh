// Opt-in FP64 GEMM from the INT8 tensor pipe (tcgen05.mma kind::i8, Ozaki splitting); see ozaki.cu.
#pragma once
#include "common.cuh"

namespace sgp {

// self test of the tcgen05 plumbing: one 128 x 64 x K INT8 product on the tensor pipe against a plain integer kernel;
// *mismatches = number of differing INT32 outputs (0 = pass); the two probes are element (5,3) of the reference / the result
int i8mma_selftest(Ctx& c, int K, int* mismatches, int* probe_ref, int* probe_got);

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

}  // namespace sgp
