// Batched symplectic map application (see map.cu).
#pragma once
#include "common.cuh"

namespace sgp {

enum MapKind : int {
    MAP_PENDULUM = 0,   // q <- mod(q + dq, 2pi)                         python/functions/func.py:216-237
    MAP_HENON = 1,      // q <- q + dq                                   python/functions/func.py:239-260
    MAP_STANDARD = 2,   // q <- mod(q + dq, 2pi), p <- mod(P, 2pi), pdiff python/04_standard_map/func.py:218-254
    MAP_TOKAMAK = 3,    // pendulum wrap + loss test (compute_r)         python/05_tokamak/SympGPR/func.py:182-211
};

struct MapArgs {
    // ordinary GP (guess): features + alphap, padded to map_pad(np)
    const double *gu, *gv, *gy, *ga;
    long np_pad;
    // symplectic GP: features + alpha (q part, P part), padded to map_pad(nt)
    const double *tu, *tv, *ty, *taq, *taP;
    long nt_pad;
    HypC h, hp;
    int kind;
    long E, nsteps;
    const double *q0, *p0;
    // history: row r (= step / out_every) of orbit k at [r*step_stride + k*orbit_stride]; out_every = 0: none
    double *qout, *pout, *pdiff;
    long step_stride, orbit_stride, out_every;
    double *qfinal, *pfinal;
    unsigned long long* stats;   // [0] residual evaluations, [1] solver exits without convergence
};

long map_pad(long n);
int map_prepare(Ctx& c, int fam, double per, const double* x, const double* y, long n, double* u, double* v, double* yo);
int map_pad_copy(Ctx& c, const double* src, long n, double* dst);
int map_launch(Ctx& c, int fam, int solver, const MapArgs& a);

}  // namespace sgp
