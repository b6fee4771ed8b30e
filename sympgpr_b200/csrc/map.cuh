// Batched symplectic map application (see map.cu).
#pragma once
#include "common.cuh"

namespace sgp {

enum MapKind : int {
    MAP_PENDULUM = 0,   // q <- mod(q + dq, 2pi)                         python/functions/func.py:216-237
    MAP_HENON = 1,      // q <- q + dq                                   python/functions/func.py:239-260
    MAP_STANDARD = 2,   // q <- mod(q + dq, 2pi), p <- mod(P, 2pi), pdiff python/04_standard_map/func.py:218-254
    MAP_TOKAMAK = 3,    // pendulum wrap + loss test (compute_r)         python/05_tokamak/SympGPR/func.py:182-211
    MAP_STANDARD_EXPL = 4,  // p <- mod(P, 2pi), pdiff, q <- q + dq (no wrap): applymap_expl python/04_standard_map/func.py:256-285
    MAP_TOKAMAK_SPLIT = 5,  // split map: model (step-1) mod nmodels, loss test at the NEW angle, lost orbits NaN in q and p:
                            // applymap_tok python/05_tokamak/Split_SympGPR/func.py:184-219
};

constexpr int MAP_CHUNK = 64;        // training points per chunk (one bulk copy)
constexpr int MAP_GF = 4;            // fields of a guess-GP chunk:      u, v, y, alpha
constexpr int MAP_TF = 5;            // fields of a symplectic-GP chunk: u, v, y, alpha_q, alpha_P

// One learned map on the device.  Training sets in chunked structure-of-arrays form: chunk c =
// [field0(64) field1(64) ...], padded with neutral points (alpha = 0) to a whole number of chunks, at least two.
struct MapModelDev {
    const double* gch; int nchg;     // ordinary GP (guess)
    const double* tch; int ncht;     // symplectic GP
    HypC h, hp;
};

struct MapArgs {
    const MapModelDev* models;       // device table; step s uses models[(s - 1) % nmodels] (split maps cycle, others have one)
    int nmodels;
    int kind;
    long E, nsteps;
    const double *q0, *p0;
    // history: row r (= step / out_every) of orbit k at [r*step_stride + k*orbit_stride]; out_every = 0: none
    double *qout, *pout, *pdiff;
    long step_stride, orbit_stride, out_every;
    double *qfinal, *pfinal;         // last state; also carries the state from one work item to the next
    double* pdstate;                 // running pdiff between work items (only when pdiff != nullptr)
    unsigned long long* stats;       // [0] residual evaluations, [1] solver exits without convergence
    // work distribution: a work item is slice_steps steps of one batch of 32 orbits; runnable batches wait
    // in a FIFO (map.cu).  All of it is zeroed before launch.
    long slice_steps;
    unsigned long long* ticket;      // [0] pop counter, [1] scheduler error word, [2] push counter
    unsigned long long* slots;       // nbatches queue slots: (position + 1) << 32 | batch, or ... | 0xffffffff once consumed
    int* progress;                   // nbatches: slices already done
};

long map_chunks(long n);                                   // chunks a set of n points occupies
size_t map_model_doubles(long np, long nt);
size_t map_sched_bytes(long E);                            // ticket + slice_done scratch
// chunked layouts from plain arrays (device pointers)
int map_prepare_guess(Ctx& c, int fam, double per, const double* x, const double* y, const double* alpha, long n, double* gch);
int map_prepare_sympl(Ctx& c, int fam, double per, const double* x, const double* y, const double* alpha, long n, double* tch);
// a.ticket / a.slots / a.progress / a.slice_steps are filled in by map_launch from `sched` (map_sched_bytes(E) bytes)
int map_launch(Ctx& c, int fam, int solver, MapArgs a, void* sched);

}  // namespace sgp
