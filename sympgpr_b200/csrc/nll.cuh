// NLL / gradient / fit pipeline on device buffers (see nll.cu).
#pragma once
#include "common.cuh"

namespace sgp {

struct NllJob {
    int fam;            // Family
    double per;         // period parameter p (0.5 for the committed product kernel)
    int reg;            // 0: derivative (Hessian-block) kernel, order n = 2N; 1: plain kernel, n = N;
                        // 2 / 3: the (q,q) / (P,P) Hessian block alone, n = N (nll_expl, 04_standard_map/func.py:126-141)
                        // 4: 2-DOF 4x4-block derivative kernel, n = 4N, d_x = [q1; q2; P1; P2] (dof2.cu, not in the reference)
    double hyp[4];      // lx, ly, sig, sig2n
    long n;             // matrix order
    const double* d_x;  // device, 2N: [x(0:N); y(0:N)]   (the reference's xin layout)
    const double* d_z;  // device, n: observations
    int ngrad;          // 0: value only; 2 or 3: also gradient (d/dlx, d/dly[, d/dsig])
    double* d_res;      // device, RES_DOUBLES
    double* d_alpha;    // device n or nullptr
    double* d_kinv;     // device n*n (ld n, full symmetric) or nullptr
    double* d_L;        // device n*n (ld n, lower, zeros above) or nullptr
};

// Enqueues the whole evaluation on c.stream; no host synchronisation.
int nll_enqueue(Ctx& c, const NllJob& job);

}  // namespace sgp
