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

// energy function of the fused quality metric (python/functions/func.py:262-272 `quality`: Eosc = std(H)/mean(H) per orbit)
enum EnergyKind : int {
    ENERGY_NONE = 0,
    ENERGY_PENDULUM = 1,   // H = p^2/2 + U0 (1 - cos(q + pi)), epar[0] = U0: energy() python/01_pendulum/implicit/func.py:116-117
    ENERGY_TOKAMAK = 2,    // H = -Aph(r, q, 0), r = compute_r([p 1e-2, q, 0], 0.3), epar = {eps, m, phase}:
                           // energy() python/05_tokamak/Split_SympGPR/func.py:234-246, Aph fieldlines.f90:58-64
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
    int coop_max;                    // up to this many unconverged lanes of a step are served by cooperative passes (map.cu)
    int newton_max_nb, newton_max_b; // Newton give-up limits (evaluations) without / with a bracket
    int start_delta;                 // 1: the guess GP predicts P - p (SGP_SOLVER_NEWTON_DELTA): the solver starts at p + guess
    const double *q0, *p0;
    // history: row r (= step / out_every) of orbit k at [r*step_stride + k*orbit_stride]; out_every = 0: none
    double *qout, *pout, *pdiff;
    long step_stride, orbit_stride, out_every;
    double *qfinal, *pfinal;         // last state; also carries the state from one work item to the next
    double* pdstate;                 // running pdiff between work items (only when pdiff != nullptr)
    unsigned long long* stats;       // [0] residual evaluations, [1] solver exits without convergence
    // fused quality metrics (ekind != ENERGY_NONE): running mean / sum of squares of H - H(0) per orbit over the rows
    // 0, e_every, 2 e_every, ... (Welford), so that no history has to be written; first mapped state for `gd`
    int ekind;
    long e_every;
    double epar[4];
    double *ek, *emean, *em2;        // E each: H(0), mean of H - H(0), sum of squared deviations
    double *q1, *p1;                 // state after e_every steps (row 1 of the sampled history) or nullptr
    double *eosc, *ehmean;           // E each, written with the last step: std(H)/mean(H) and mean(H)
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
// 2-DOF map (map.cu, "2-DOF map"): chunks the set needs, and the launch on device arrays (x, alpha: (4, n); q0, p0, finals: (2, E);
// histories (rows, 2, E) or nullptr; d_set: map4_chunks(n) * 256 doubles of scratch)
long map4_chunks(long n);
int map4_run(Ctx& c, const double* d_x, const double* d_alpha, long n, double lq, double lP, double sig, double* d_set, long E,
             long nsteps, const double* d_q0, const double* d_p0, double* d_qhist, double* d_phist, long out_every, double* d_qfinal,
             double* d_pfinal, unsigned long long* d_stats);
// StandardMapIterate (python/04_standard_map/main.py:32-39) on device arrays: X0 (2, N), f (2, N, nm)
int standard_map_iterate(Ctx& c, double kk, long nm, long N, const double* d_X0, double* d_f);

}  // namespace sgp
