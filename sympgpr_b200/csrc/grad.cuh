// Gradient contraction (see grad.cu).
#pragma once
#include "common.cuh"

namespace sgp {

long grad_num_partials(long N);
int grad_contract(Ctx& c, int fam, int reg, const Pt* pts, long N, const HypC& h, const double* Kinv, long ld,
                  const double* alpha, double* partial);

}  // namespace sgp
