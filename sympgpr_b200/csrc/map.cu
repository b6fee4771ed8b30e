// Batched application of the learned symplectic map: one thread per orbit, many map steps per
// launch, training set streamed through shared memory.
//
// Replaces the Python/Fortran ensemble loops
//   applymap        python/functions/func.py:216-237          (q wrapped mod 2pi)
//   applymap_henon  python/functions/func.py:239-260          (no wrap)
//   applymap        python/04_standard_map/func.py:218-254    (+ p wrapped, running pdiff)
//   applymap_tok    python/05_tokamak/SympGPR/func.py:182-211 (+ loss test via compute_r)
//   sympgpr.f90::applymap_tok :128-177, ::calcP :88-125, ::calcq :75-86, ::guessP :62-73
// with alpha = Kyinv*ztrain hoisted out of every evaluation (the reference recomputes the
// O(Nt^2) matvec per call, sympgpr.f90:72,85,121).
//
// One map step of one orbit (q,p):
//   1. P0 = sum_j sigp k(a_j;(q,p)) alphap_j                        (ordinary GP, guess)
//   2. solve  f(P) = sum_j sig [kxx_j aq_j + kxy_j aP_j](q,P) - p + P = 0   (Hybrd1 or Newton1)
//   3. dq   = sum_j sig [kxy_j aq_j + kyy_j aP_j](q,P)
//   4. post-step variant (wraps / loss test), see MapKind.
// Every sum is a sweep of the whole block over the training set in 512-point chunks staged in
// shared memory; threads read the same training point at the same time (broadcast), so the
// kernel is bound by the FP64 pipe (about 45 DFMA-equivalents per pair incl. one exp), not by
// shared memory or HBM.
#include "map.cuh"

#include "hybrd.cuh"

namespace sgp {

constexpr int MAP_THREADS = 128;
constexpr int MAP_CH = 512;           // training points per shared-memory chunk
constexpr int MAP_MIN_BLOCKS = 5;     // 640 threads / SM: registers capped at 96
constexpr double TWO_PI = 6.283185307179586;

__device__ __forceinline__ double np_mod(double a, double b)
{   // numpy.mod for b > 0
    double r = fmod(a, b);
    if (r != 0.0 && r < 0.0) r += b;
    return r;
}

// fieldlines.f90:94-107 with f_r :82-91, Ath :34-39, dAthdr :42-47 (B0 = R0 = 1): 20 Newton steps
__host__ __device__ inline double compute_r_dev(double pth, double th, double rstart)
{
    double r = rstart;
    const double ct = cos(th);
    for (int k = 0; k < 20; k++) {
        const double yv = pth - (r * r / 2.0 - r * r * r / 3.0 * ct);
        const double dy = -(r - r * r * ct);
        r = r - yv / dy;
    }
    return r;
}

enum SweepMode : int { SW_GUESS = 0, SW_F_DF = 1, SW_F = 2, SW_DQ = 3 };

// One block-wide sweep.  fields: guess GP 4 (u, v, y, alpha), symplectic GP 5 (u, v, y, aq, aP).
template <int FAM, int MODE>
__device__ __forceinline__ void sweep(const double* const* __restrict__ fld, long n_pad, double* sm, const Pt& b,
                                      const HypC& h, bool active, double& o0, double& o1)
{
    constexpr int NF = (MODE == SW_GUESS) ? 4 : 5;
    double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
    for (long c0 = 0; c0 < n_pad; c0 += MAP_CH) {
        __syncthreads();
#pragma unroll
        for (int f = 0; f < NF; f++) {
            const double* g = fld[f] + c0;
            for (int idx = threadIdx.x; idx < MAP_CH; idx += MAP_THREADS) sm[f * MAP_CH + idx] = g[idx];
        }
        __syncthreads();
        if (active) {
#pragma unroll 2
            for (int j = 0; j < MAP_CH; j += 2) {
                const double2 u2 = *reinterpret_cast<const double2*>(sm + 0 * MAP_CH + j);
                const double2 v2 = *reinterpret_cast<const double2*>(sm + 1 * MAP_CH + j);
                const double2 y2 = *reinterpret_cast<const double2*>(sm + 2 * MAP_CH + j);
                const double2 a2 = *reinterpret_cast<const double2*>(sm + 3 * MAP_CH + j);
                Pt p0, p1;
                p0.u = u2.x; p0.v = v2.x; p0.y = y2.x;
                p1.u = u2.y; p1.v = v2.y; p1.y = y2.y;
                const Pair<FAM> q0(p0, b, h), q1(p1, b, h);
                if (MODE == SW_GUESS) {
                    s0 += q0.k() * a2.x;
                    s1 += q1.k() * a2.y;
                } else {
                    const double2 c2 = *reinterpret_cast<const double2*>(sm + 4 * MAP_CH + j);
                    if (MODE == SW_F_DF || MODE == SW_F) {
                        s0 += q0.kxx(h) * a2.x + q0.kxy(h) * c2.x;
                        s1 += q1.kxx(h) * a2.y + q1.kxy(h) * c2.y;
                        if (MODE == SW_F_DF) {
                            t0 += q0.kxx_yb(h) * a2.x + q0.kxy_yb(h) * c2.x;
                            t1 += q1.kxx_yb(h) * a2.y + q1.kxy_yb(h) * c2.y;
                        }
                    } else {
                        s0 += q0.kxy(h) * a2.x + q0.kyy(h) * c2.x;
                        s1 += q1.kxy(h) * a2.y + q1.kyy(h) * c2.y;
                    }
                }
            }
        }
    }
    o0 = h.sig * (s0 + s1);
    o1 = h.sig * (t0 + t1);
}

template <int FAM, int SOLVER>
__global__ void __launch_bounds__(MAP_THREADS, MAP_MIN_BLOCKS)
map_kernel(MapArgs a)
{
    __shared__ __align__(16) double sm[5 * MAP_CH];
    __shared__ unsigned long long s_evals;
    __shared__ unsigned int s_fail;
    if (threadIdx.x == 0) { s_evals = 0ull; s_fail = 0u; }
    __syncthreads();

    const long k = (long)blockIdx.x * MAP_THREADS + threadIdx.x;
    const bool mine = k < a.E;
    double q = mine ? a.q0[k] : 0.0;
    double p = mine ? a.p0[k] : 0.0;
    double pd = p;                       // running pdiff (standard map)
    unsigned long long evals = 0ull;
    unsigned int fails = 0u;

    const double* gf[4] = {a.gu, a.gv, a.gy, a.ga};
    const double* tf[5] = {a.tu, a.tv, a.ty, a.taq, a.taP};

    if (mine && a.out_every > 0) {
        a.qout[k * a.orbit_stride] = q;
        a.pout[k * a.orbit_stride] = p;
        if (a.pdiff) a.pdiff[k * a.orbit_stride] = pd;
    }

    for (long step = 1; step <= a.nsteps; step++) {
        // an orbit that is already NaN stays NaN (tokamak: explicit test, func.py:192-193; the
        // other variants propagate it through the arithmetic)
        bool alive = mine && (q == q) && (p == p);
        Pt b;
        if (alive) { b = make_pt<FAM>(q, p, a.h.p); } else { b.u = 0; b.v = 1; b.y = 0; }

        double pg, dummy;
        sweep<FAM, SW_GUESS>(gf, a.np_pad, sm, b, a.hp, alive, pg, dummy);
        if (alive && !(fabs(pg) <= DBL_MAX)) alive = false;

        double P;
        if (SOLVER == 0) {
            Hybrd1 sv;
            sv.start(alive ? pg : 0.0);
            if (!alive) sv.phase = 3;
            for (;;) {
                const bool run = !sv.done();
                b.y = sv.query();
                double F, dF;
                sweep<FAM, SW_F>(tf, a.nt_pad, sm, b, a.h, run, F, dF);
                if (run) { sv.feed(F - p + b.y); evals++; }
                if (__syncthreads_and(sv.done())) break;
            }
            P = sv.root();
            if (alive && sv.info != 1) fails++;
        } else {
            Newton1 sv;
            sv.start(alive ? pg : 0.0);
            if (!alive) sv.phase = 3;
            for (;;) {
                const bool run = !sv.done();
                b.y = sv.query();
                double F, dF;
                sweep<FAM, SW_F_DF>(tf, a.nt_pad, sm, b, a.h, run, F, dF);
                if (run) { sv.feed(F - p + b.y, 1.0 + dF); evals++; }
                if (__syncthreads_and(sv.done())) break;
            }
            P = sv.root();
            if (alive && sv.info != 1 && sv.info != 3) fails++;
        }

        double Pst = P;
        if (alive) {
            if (a.kind == MAP_STANDARD) {
                pd = pd + (P - p);
                Pst = np_mod(P, TWO_PI);
            } else if (a.kind == MAP_TOKAMAK) {
                const double r = compute_r_dev(P * 1e-2, q, 0.3);
                if (r > 0.5 || P < 0.0) Pst = nan("");
            }
        } else {
            Pst = nan("");
        }
        const bool qalive = alive && (Pst == Pst);
        b.y = Pst;
        double dq;
        sweep<FAM, SW_DQ>(tf, a.nt_pad, sm, b, a.h, qalive, dq, dummy);
        if (qalive) evals++;
        double qn;
        if (!qalive) qn = nan("");
        else if (a.kind == MAP_HENON) qn = dq + q;
        else qn = np_mod(dq + q, TWO_PI);
        if (mine && !alive) pd = nan("");
        q = qn;
        p = Pst;

        if (mine && a.out_every > 0 && (step % a.out_every) == 0) {
            const long row = step / a.out_every;
            a.qout[row * a.step_stride + k * a.orbit_stride] = q;
            a.pout[row * a.step_stride + k * a.orbit_stride] = p;
            if (a.pdiff) a.pdiff[row * a.step_stride + k * a.orbit_stride] = pd;
        }
    }
    if (mine) {
        a.qfinal[k] = q;
        a.pfinal[k] = p;
    }
    atomicAdd(&s_evals, evals);
    atomicAdd(&s_fail, fails);
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(a.stats + 0, s_evals);
        atomicAdd(a.stats + 1, (unsigned long long)s_fail);
    }
}

// training features in structure-of-arrays form, padded with neutral points (alpha = 0)
template <int FAM>
__global__ void map_prep_kernel(const double* __restrict__ x, const double* __restrict__ y, long n, long n_pad, double p,
                                double* __restrict__ u, double* __restrict__ v, double* __restrict__ yo)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    if (i < n) {
        const Pt t = make_pt<FAM>(x[i], y[i], p);
        u[i] = t.u; v[i] = t.v; yo[i] = t.y;
    } else {
        u[i] = 0.0; v[i] = 1.0; yo[i] = 0.0;
    }
}

__global__ void pad_copy_kernel(const double* __restrict__ src, long n, long n_pad, double* __restrict__ dst)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) dst[i] = (i < n) ? src[i] : 0.0;
}

long map_pad(long n) { return round_up(n > 0 ? n : 1, MAP_CH); }

int map_prepare(Ctx& c, int fam, double per, const double* x, const double* y, long n, double* u, double* v, double* yo)
{
    const long n_pad = map_pad(n);
    const unsigned g = (unsigned)((n_pad + 255) / 256);
    if (fam == FAM_SQ) map_prep_kernel<FAM_SQ><<<g, 256, 0, c.stream>>>(x, y, n, n_pad, per, u, v, yo);
    else map_prep_kernel<FAM_PRODUCT><<<g, 256, 0, c.stream>>>(x, y, n, n_pad, per, u, v, yo);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

int map_pad_copy(Ctx& c, const double* src, long n, double* dst)
{
    const long n_pad = map_pad(n);
    pad_copy_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, c.stream>>>(src, n, n_pad, dst);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

int map_launch(Ctx& c, int fam, int solver, const MapArgs& a)
{
    if (a.E <= 0) return ST_OK;
    const unsigned g = (unsigned)((a.E + MAP_THREADS - 1) / MAP_THREADS);
#define ML(F, S) map_kernel<F, S><<<g, MAP_THREADS, 0, c.stream>>>(a)
    if (solver == 0) {
        switch (fam) {
        case FAM_PRODUCT: ML(FAM_PRODUCT, 0); break;
        case FAM_SQ: ML(FAM_SQ, 0); break;
        case FAM_SUM: ML(FAM_SUM, 0); break;
        default: set_error("unknown kernel family %d", fam); return ST_BADARG;
        }
    } else {
        switch (fam) {
        case FAM_PRODUCT: ML(FAM_PRODUCT, 1); break;
        case FAM_SQ: ML(FAM_SQ, 1); break;
        case FAM_SUM: ML(FAM_SUM, 1); break;
        default: set_error("unknown kernel family %d", fam); return ST_BADARG;
        }
    }
#undef ML
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

}  // namespace sgp
