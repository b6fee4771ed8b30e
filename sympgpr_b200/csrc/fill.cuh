// Kernel-matrix fills (see fill.cu).
#pragma once
#include "common.cuh"

namespace sgp {

int make_points(Ctx& c, int fam, double p, const double* x, const double* y, long n, Pt* out);
int fill_hess(Ctx& c, int fam, const Pt* pb, long N, const Pt* pa, long N0, const HypC& h, double* K, long ld);
int fill_hess_sym(Ctx& c, int fam, const Pt* pts, long N, const HypC& h, double noise, double* K, long ld, long n_pad);
int fill_reg(Ctx& c, int fam, const Pt* pb, long N, const Pt* pa, long N0, const HypC& h, double* K, long ld);
// which: 0 kernel value; 1 / 2 the (q,q) / (P,P) Hessian block alone
int fill_reg_sym(Ctx& c, int fam, const Pt* pts, long N, const HypC& h, double noise, double* K, long ld, long n_pad, int which = 0);

}  // namespace sgp
