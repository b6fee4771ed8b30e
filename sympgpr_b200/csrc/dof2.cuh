// 2-DOF (4 x 4 block) Hessian kernel: fill and gradient contraction (see dof2.cu).
#pragma once
#include "common.cuh"

namespace sgp {

// x / x0: [q1; q2; P1; P2] (4 arrays of N / N0 values); K: (4N x 4N0), column-major
int fill4(Ctx& c, const double* x, long N, const double* x0, long N0, double lq, double lP, double sig, double* K, long ld);
int fill4_sym(Ctx& c, const double* x, long N, double lq, double lP, double sig, double noise, double* K, long ld, long n_pad);
long grad4_num_partials(long N);
int grad4_contract(Ctx& c, const double* x, long N, double lq, double lP, double sig, const double* Kinv, long ld,
                   const double* alpha, double* partial);

}  // namespace sgp
