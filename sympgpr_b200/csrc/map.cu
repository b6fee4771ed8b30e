// Batched application of the learned symplectic map: one thread per orbit, one autonomous warp per
// batch of 32 orbits, many map steps per launch, training set streamed through shared memory.
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
// Every sum is a sweep over the training set.
//
// Execution model (B200): the grid is persistent and every WARP is its own worker -- there is no
// block-level synchronisation anywhere.  A warp pulls work items (32 orbits x slice_steps steps) from
// an atomic ticket, keeps the orbit and solver state of its 32 orbits in registers, and streams the
// training set through its private double buffer in shared memory: lane 0 issues one cp.async.bulk
// (TMA engine, SASS UBLKCP) per 64-point chunk, completion is counted on a per-buffer mbarrier, and
// the copy of chunk c+2 is in flight while the warp evaluates chunk c with broadcast LDS.128 reads
// (two training points per load).  The 32 lanes advance their solvers in lock step only within the
// warp, so a slow root solve delays 31 neighbours, not a thread block -- and not for long: once at most
// MAP_COOP_MAX lanes are still iterating, each of them is served by a COOPERATIVE pass in which the 32 lanes
// share the training set for that one orbit (1/32 of the arithmetic of a full pass).  Slicing the step
// loop into work items keeps all SMs busy to the end whatever the ensemble size.  The kernel is bound
// by the FP64 pipe (about 50 DP instructions per orbit-point pair incl. one exp), not by shared
// memory, L2 or HBM: a chunk of 2.5 KB feeds 32 x 64 pair evaluations.
//
// Summation order is fixed (full pass: chunk by chunk, two interleaved partial sums; cooperative pass:
// per-lane partial sums, then a fixed xor-shuffle tree), so results do not depend on scheduling; which of
// the two a given evaluation uses depends on how many lanes of the batch were still iterating.
#include "map.cuh"

#include <cstdlib>

#include "hybrd.cuh"
#include "mbar.cuh"

namespace sgp {

constexpr int MAP_THREADS = 128;
constexpr int MAP_WARPS = MAP_THREADS / 32;
constexpr int MAP_COOP_MAX = 16;                   // up to this many unconverged lanes are served by cooperative passes
constexpr int MAP_BLOCKS_PER_SM = 4;               // 16 warps / SM, <= 128 registers per thread
constexpr int MAP_BUF_DOUBLES = MAP_TF * MAP_CHUNK; // one buffer holds a chunk of either set
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

// fieldlines.f90:58-64 at ph = 0 (B0 = iota0 = 1, a = 0.5): Aph = -(r^2/2 - r^4/(4 a^2)) (1 + eps cos(m th + phase))
__host__ __device__ inline double aph_dev(double r, double th, double eps, double m, double phase)
{
    const double r2 = r * r;
    return -(r2 / 2.0 - r2 * r2 / (4.0 * 0.25)) * (1.0 + eps * cos(m * th + phase));
}

__device__ __forceinline__ double orbit_energy(int ekind, double e0, double e1, double e2, double q, double p)
{
    if (ekind == ENERGY_PENDULUM) return p * p / 2.0 + e0 * (1.0 - cos(q + 3.141592653589793));
    const double r = compute_r_dev(p * 1e-2, q, 0.3);
    return -aph_dev(r, q, e0, e1, e2);
}

enum SweepMode : int { SW_GUESS = 0, SW_F_DF = 1, SW_F = 2, SW_DQ = 3 };

// 2^(j/64), j = 0..63, correctly rounded (filled by the host on first launch; copied to shared memory per block)
__constant__ double c_exp2_tab[64];

// exp(x) for x <= 0 in the sweeps: N = round(64 x / ln2) = 64 k + j, r = x - N ln2/64 (two-term
// Cody-Waite, |r| <= ln2/128), exp(x) = 2^k * T[j] * (1 + q), q = e^r - 1 by a degree-5 polynomial;
// T from a 64-entry table in shared memory.  About 12 DP instructions and one LDS, against 21 for the
// table-free exp_neg of forms.cuh; relative error < 1.2 ulp.  Arguments below -700 are clamped (the
// result, < 1e-304, stands for a true value that is smaller still); the sweeps never pass NaN (dead
// orbits are masked out before).
__device__ __forceinline__ double exp_neg_tab(double x, const double* __restrict__ tab)
{
    const double C = 92.33248261689366;             // 64 / ln 2
    const double SHIFT = 6755399441055744.0;        // 1.5 * 2^52
    const double L_HI = 0.010830424695086549;       // ln2/64, low 20 mantissa bits zero
    const double L_LO = 1.162596423439437e-12;
    // clamp to >= -700 (x <= 0: the order of magnitudes is the unsigned order of the high words)
    const unsigned hic = min((unsigned)__double2hiint(x), 0xC085E000u);
    const double xc = __hiloint2double((int)hic, __double2loint(x));
    const double t = fma(xc, C, SHIFT);
    const int N = __double2loint(t);
    const double kd = t - SHIFT;
    double r = fma(kd, -L_HI, xc);
    r = fma(kd, -L_LO, r);
    double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = q * r;
    const double T = tab[(N & 63) << 5];               // the table is replicated per lane (bank-conflict free), tab = base + lane
    double res = fma(T, q, T);
    return __hiloint2double(__double2hiint(res) + ((N >> 6) << 20), __double2loint(res));
}

// per-warp chunk stream: two buffers, one mbarrier each; `seq` counts the chunks consumed so far
struct Stream {
    double* buf;                 // 2 * MAP_BUF_DOUBLES
    unsigned long long* bar;     // 2
    uint32_t seq;
    unsigned int passes;         // sweeps this warp has run (utilisation statistics)
    const double* primed;        // the set whose chunks 0 and 1 are in flight between sweeps
};

__device__ __forceinline__ void stream_issue(const Stream& s, uint32_t slot, const double* src, uint32_t bytes)
{
    mbar_arrive_expect_tx(s.bar + slot, bytes);
    bulk_g2s(s.buf + slot * MAP_BUF_DOUBLES, src, bytes, s.bar + slot);
}

// Accumulators of one sweep: two interleaved partial sums (even / odd training points) of up to four
// weighted sums; what they mean depends on the mode (see the end of sweep()).
struct Acc {
    double a0, a1, b0, b1, c0, c1, d0, d1;
};

// One training point against the lane's query point b, product (periodic x SE) and SE x SE families,
// with every hyper-parameter factor pulled out of the sum (they are applied once at the end of sweep()):
//   E = exp(-dy^2 hy - s^2 hx),  bxx = lx^2 cos(2 p dx) - s^2 c^2 | lx^2 - dx^2,  odd = s c | dx
//   F    = sig        (cxx sum E bxx aq         - cxy sum E dy odd aP)            kernels.f90:58-94
//   dF   = sig / ly^2 (cxx sum E dy bxx aq      + cxy sum E (ly^2 - dy^2) odd aP) kernels.f90:95-132
//   dq   = sig        (cyy sum E (ly^2-dy^2) aP - cxy sum E dy odd aq)
//   pg   = sigp sum E alpha                                                        kernels.f90:1-11
template <int FAM, int MODE>
__device__ __forceinline__ void point_eval(double au, double av, double ay, double w0, double w1, const Pt& b, double nhx,
                                           double nhy, double lx2, double ly2, const double* __restrict__ tab, double& A,
                                           double& B, double& C, double& D)
{
    const double dy = ay - b.y;
    double s, odd, bxx;
    if (FAM == FAM_SQ) {
        s = au - b.u;
        odd = s;
        bxx = fma(-s, s, lx2);
    } else {
        s = fma(au, b.v, -(av * b.u));
        const double c = fma(av, b.v, au * b.u);
        odd = s * c;
        const double s2 = s * s;
        bxx = fma(lx2, fma(-2.0, s2, 1.0), -(odd * odd));
    }
    const double E = exp_neg_tab(fma(s * s, nhx, (dy * dy) * nhy), tab);
    if (MODE == SW_GUESS) {
        A = fma(E, w0, A);
    } else if (MODE == SW_F_DF || MODE == SW_F) {
        const double ea = E * (bxx * w0);        // E bxx aq
        const double eb = E * (odd * w1);        // E odd aP
        A += ea;
        B = fma(dy, eb, B);
        if (MODE == SW_F_DF) {
            C = fma(dy, ea, C);
            D = fma(fma(-dy, dy, ly2), eb, D);
        }
    } else {
        const double eb = E * (odd * w0);        // E odd aq
        A = fma(fma(-dy, dy, ly2) * E, w1, A);   // E (ly^2 - dy^2) aP
        B = fma(dy, eb, B);
    }
}

// One sweep of the warp over a whole training set (nch chunks of NF fields); the first two chunks are
// already in flight.  While chunk c is evaluated, chunk c+2 is fetched -- for the last two chunks that
// is chunk 0/1 of the set the NEXT sweep will read (nxt, nxf fields).
// COOP = true: ONE query point (the caller broadcasts a straggler's b to all lanes) and the 32 lanes share the
// training set -- lane l evaluates points 2l, 2l+1 of every 64-point chunk -- then the partial sums are added across
// the warp in a fixed order.  1/32 of the arithmetic of a full pass: this is how the last few unconverged lanes of a
// step are served, instead of a full pass in which 31 lanes idle (see the solver loops in map_kernel).
template <int FAM, int MODE, bool COOP = false>
__device__ __forceinline__ void sweep(Stream& st, const double* __restrict__ set, int nch, const double* __restrict__ nxt, int nxf,
                                      int lane, const Pt& b, const HypC& h, const double* __restrict__ tab, bool active,
                                      double& o0, double& o1)
{
    constexpr int NF = (MODE == SW_GUESS) ? MAP_GF : MAP_TF;
    Acc z;
    z.a0 = z.a1 = z.b0 = z.b1 = z.c0 = z.c1 = z.d0 = z.d1 = 0.0;
    const double nhx = -h.hx, nhy = -h.hy, lx2 = h.lx2, ly2 = h.ly2;
    const bool any = COOP ? true : __any_sync(0xffffffffu, active);
    if (!COOP) st.passes++;
    for (int c = 0; c < nch; c++) {
        const uint32_t slot = st.seq & 1u, par = (st.seq >> 1) & 1u;
        mbar_wait(st.bar + slot, par);
        const double* sm = st.buf + slot * MAP_BUF_DOUBLES;
        if (any) {
#pragma unroll 2
            for (int j = COOP ? 2 * lane : 0; j < MAP_CHUNK; j += COOP ? MAP_CHUNK : 2) {
                const double2 u2 = *reinterpret_cast<const double2*>(sm + 0 * MAP_CHUNK + j);
                const double2 v2 = *reinterpret_cast<const double2*>(sm + 1 * MAP_CHUNK + j);
                const double2 y2 = *reinterpret_cast<const double2*>(sm + 2 * MAP_CHUNK + j);
                const double2 a2 = *reinterpret_cast<const double2*>(sm + 3 * MAP_CHUNK + j);
                double2 c2 = make_double2(0.0, 0.0);
                if (MODE != SW_GUESS) c2 = *reinterpret_cast<const double2*>(sm + 4 * MAP_CHUNK + j);
                if (FAM != FAM_SUM) {
                    point_eval<FAM, MODE>(u2.x, v2.x, y2.x, a2.x, c2.x, b, nhx, nhy, lx2, ly2, tab, z.a0, z.b0, z.c0, z.d0);
                    point_eval<FAM, MODE>(u2.y, v2.y, y2.y, a2.y, c2.y, b, nhx, nhy, lx2, ly2, tab, z.a1, z.b1, z.c1, z.d1);
                } else {
                    // sum family (kernels_expl_per_q_sq_p.f90): generic closed forms; a0/a1 = value, c0/c1 = derivative
                    Pt p0, p1;
                    p0.u = u2.x; p0.v = v2.x; p0.y = y2.x;
                    p1.u = u2.y; p1.v = v2.y; p1.y = y2.y;
                    const Pair<FAM> q0(p0, b, h), q1(p1, b, h);
                    if (MODE == SW_GUESS) {
                        z.a0 += q0.k() * a2.x;
                        z.a1 += q1.k() * a2.y;
                    } else if (MODE == SW_F_DF || MODE == SW_F) {
                        z.a0 += q0.kxx(h) * a2.x + q0.kxy(h) * c2.x;
                        z.a1 += q1.kxx(h) * a2.y + q1.kxy(h) * c2.y;
                        if (MODE == SW_F_DF) {
                            z.c0 += q0.kxx_yb(h) * a2.x + q0.kxy_yb(h) * c2.x;
                            z.c1 += q1.kxx_yb(h) * a2.y + q1.kxy_yb(h) * c2.y;
                        }
                    } else {
                        z.a0 += q0.kxy(h) * a2.x + q0.kyy(h) * c2.x;
                        z.a1 += q1.kxy(h) * a2.y + q1.kyy(h) * c2.y;
                    }
                }
            }
        }
        __syncwarp();                                        // every lane is done with this buffer
        if (lane == 0) {
            if (c + 2 < nch) stream_issue(st, slot, set + (size_t)(c + 2) * NF * MAP_CHUNK, NF * MAP_CHUNK * sizeof(double));
            else stream_issue(st, slot, nxt + (size_t)(c + 2 - nch) * nxf * MAP_CHUNK, nxf * MAP_CHUNK * sizeof(double));
        }
        st.seq++;
    }
    st.primed = nxt;
    double A = z.a0 + z.a1, B = z.b0 + z.b1, C = z.c0 + z.c1, D = z.d0 + z.d1;
    if (COOP) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            A += __shfl_xor_sync(0xffffffffu, A, o);
            if (MODE != SW_GUESS) B += __shfl_xor_sync(0xffffffffu, B, o);
            if (MODE == SW_F_DF) {
                C += __shfl_xor_sync(0xffffffffu, C, o);
                D += __shfl_xor_sync(0xffffffffu, D, o);
            }
        }
    }
    if (FAM == FAM_SUM) {
        o0 = h.sig * A;
        o1 = h.sig * C;
    } else if (MODE == SW_GUESS) {
        o0 = h.sig * A;
        o1 = 0.0;
    } else if (MODE == SW_F_DF || MODE == SW_F) {
        o0 = h.sig * (h.cxx * A - h.cxy * B);
        o1 = h.sig * h.ily2 * (h.cxx * C + h.cxy * D);
    } else {
        o0 = h.sig * (h.cyy * A - h.cxy * B);
        o1 = 0.0;
    }
}

template <int FAM, int SOLVER>
__global__ void __launch_bounds__(MAP_THREADS, MAP_BLOCKS_PER_SM)
map_kernel(MapArgs a)
{
    __shared__ __align__(16) double s_buf[MAP_WARPS][2 * MAP_BUF_DOUBLES];
    __shared__ unsigned long long s_bar[MAP_WARPS][2];
    // exp table, one copy per lane: entry j of lane l at [j*32 + l], so that the 32 random indices of a warp never
    // meet in a bank (the single 512-byte table cost 28 % extra shared-memory wavefronts, profiles/r01_ncu_prof_map_e.txt)
    __shared__ double s_tab_all[64 * 32];
    for (int i = threadIdx.x; i < 64 * 32; i += MAP_THREADS) s_tab_all[i] = c_exp2_tab[i >> 5];
    const double* s_tab = s_tab_all + (threadIdx.x & 31);
    __syncthreads();                                         // the only block-wide barrier: before any work
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Stream st;
    st.buf = s_buf[warp];
    st.bar = s_bar[warp];
    st.seq = 0;
    st.passes = 0u;
    if (lane == 0) {
        mbar_init(st.bar + 0, 1);
        mbar_init(st.bar + 1, 1);
        mbar_fence_init();
    }
    __syncwarp();
    // prime the stream with the first two chunks of the set a step starts with: the guess set of its model,
    // or the symplectic set for the explicit map, which has no guess sweep
    constexpr int first_nf = (SOLVER == 2) ? MAP_TF : MAP_GF;
    {
        const double* fs = (SOLVER == 2) ? a.models[0].tch : a.models[0].gch;
        if (lane == 0) {
            stream_issue(st, 0, fs, first_nf * MAP_CHUNK * sizeof(double));
            stream_issue(st, 1, fs + first_nf * MAP_CHUNK, first_nf * MAP_CHUNK * sizeof(double));
        }
        st.primed = fs;
    }

    const long nbatches = (a.E + 31) / 32;
    const long nslices = a.nsteps > 0 ? (a.nsteps + a.slice_steps - 1) / a.slice_steps : 1;
    const unsigned long long nitems = (unsigned long long)nbatches * (unsigned long long)nslices;
    unsigned long long evals = 0ull;
    unsigned int fails = 0u;
    volatile unsigned long long* verr = a.ticket + 1;      // scheduler error word (a queue wait timed out)

    for (;;) {
        // ---- pop the next work item from the FIFO of runnable batches --------------------------------
        // Position t of the queue is claimed with one atomic; positions 0..nbatches-1 are the batches in
        // their initial state, later positions are filled (in completion order) by warps that finished a
        // slice of a batch with steps left.  A claimed position that is not filled yet means there is no
        // runnable work at the moment (tail of the launch): wait for the push.
        unsigned long long tk = 0ull;
        if (lane == 0) tk = atomicAdd(a.ticket, 1ull);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= nitems) break;
        long batch = (long)tk, slice = 0;
        int ok = 1;
        if (lane == 0) {
            unsigned long long* slot = a.slots + (tk % (unsigned long long)nbatches);
            if (tk >= (unsigned long long)nbatches) {
                const unsigned long long t0 = globaltimer();
                unsigned n = 0;
                unsigned long long v;
                while (((v = ld_acquire_u64(slot)) >> 32) != ((tk + 1ull) & 0xffffffffull) || (unsigned)v == 0xffffffffu) {
                    __nanosleep(200);
                    if ((++n & 255u) == 0u && (*verr != 0ull || globaltimer() - t0 > 20000000000ull)) { ok = 0; break; }
                }
                batch = (long)(unsigned)v;
                if (ok) slice = a.progress[batch];
            }
            if (ok) st_release_u64(slot, ((tk + 1ull) << 32) | 0xffffffffull);      // consumed: the slot may be refilled
            else atomicExch(a.ticket + 1, 1ull);
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) break;
        batch = __shfl_sync(0xffffffffu, batch, 0);
        slice = __shfl_sync(0xffffffffu, slice, 0);
        const long k = batch * 32 + lane;
        const bool mine = k < a.E;
        double q = 0.0, p = 0.0, pd = 0.0;
        if (mine) {
            if (slice == 0) { q = a.q0[k]; p = a.p0[k]; pd = p; }
            else { q = a.qfinal[k]; p = a.pfinal[k]; pd = a.pdstate ? a.pdstate[k] : 0.0; }
        }
        if (slice == 0 && mine && a.ekind != ENERGY_NONE) {
            a.ek[k] = orbit_energy(a.ekind, a.epar[0], a.epar[1], a.epar[2], q, p);
            a.emean[k] = 0.0;
            a.em2[k] = 0.0;
            if (a.nsteps < a.e_every) { a.eosc[k] = 0.0 / a.ek[k]; a.ehmean[k] = a.ek[k]; }     // a single sample
        }
        if (slice == 0 && mine && a.out_every > 0) {
            a.qout[k * a.orbit_stride] = q;
            a.pout[k * a.orbit_stride] = p;
            if (a.pdiff) a.pdiff[k * a.orbit_stride] = pd;
        }
        const long step_begin = slice * a.slice_steps + 1;
        long step_end = step_begin + a.slice_steps - 1;
        if (step_end > a.nsteps) step_end = a.nsteps;

        for (long step = step_begin; step <= step_end; step++) {
            const MapModelDev& M = a.models[(step - 1) % a.nmodels];
            const MapModelDev& Mn = a.models[step % a.nmodels];                   // model of the next step
            const double* first_set = (SOLVER == 2) ? M.tch : M.gch;
            const double* next_first = (SOLVER == 2) ? Mn.tch : Mn.gch;
            if (st.primed != first_set) {
                // the chunks in flight belong to another model's set (a work item that starts in the middle of a
                // split-map turn): discard them and prime again
                mbar_wait(st.bar + (st.seq & 1u), (st.seq >> 1) & 1u);
                mbar_wait(st.bar + ((st.seq + 1u) & 1u), ((st.seq + 1u) >> 1) & 1u);
                __syncwarp();
                st.seq += 2;
                if (lane == 0) {
                    stream_issue(st, st.seq & 1u, first_set, first_nf * MAP_CHUNK * sizeof(double));
                    stream_issue(st, (st.seq + 1u) & 1u, first_set + first_nf * MAP_CHUNK, first_nf * MAP_CHUNK * sizeof(double));
                }
                st.primed = first_set;
            }
            // an orbit that is already NaN stays NaN (tokamak: explicit test, func.py:192-193; the
            // other variants propagate it through the arithmetic)
            bool alive = mine && (q == q) && (p == p);
            Pt b;
            if (alive) { b = make_pt<FAM>(q, p, M.h.p); } else { b.u = 0; b.v = 1; b.y = 0; }

            double pg = 0.0, dummy;
            if (SOLVER != 2) {
                sweep<FAM, SW_GUESS>(st, M.gch, M.nchg, M.tch, MAP_TF, lane, b, M.hp, s_tab, alive, pg, dummy);
                if (alive && !(fabs(pg) <= DBL_MAX)) alive = false;
            }

            double P;
            if (SOLVER == 2) {
                // explicit map (no root solve): P = p - F_q(q, p), the generating function taken at the OLD
                // momentum -- calcP_expl python/04_standard_map/func.py:174-179, calcP
                // python/01_pendulum/explicit/func_expl.py:107-112.  No guess GP.
                b.y = p;
                double F, dF;
                sweep<FAM, SW_F>(st, M.tch, M.ncht, M.tch, MAP_TF, lane, b, M.h, s_tab, alive, F, dF);
                if (alive) evals++;
                P = p - F;
            } else if (SOLVER == 0) {
                Hybrd1 sv;
                sv.start(alive ? pg : 0.0);
                if (!alive) sv.phase = 3;
                for (;;) {
                    const unsigned pend = __ballot_sync(0xffffffffu, !sv.done());
                    if (pend == 0u) break;
                    const double yq = sv.query();
                    double F, dF;
                    if (__popc(pend) <= a.coop_max) {
                        // the last few lanes: one cooperative pass each (1/32 of the arithmetic of a full pass)
                        for (unsigned m = pend; m != 0u; m &= m - 1u) {
                            const int src = __ffs(m) - 1;
                            Pt bs;
                            bs.u = __shfl_sync(0xffffffffu, b.u, src);
                            bs.v = __shfl_sync(0xffffffffu, b.v, src);
                            bs.y = __shfl_sync(0xffffffffu, yq, src);
                            sweep<FAM, SW_F, true>(st, M.tch, M.ncht, M.tch, MAP_TF, lane, bs, M.h, s_tab, true, F, dF);
                            if (lane == src) { sv.feed(F - p + yq); evals++; }
                        }
                        st.passes++;                         // (counted as one pass in the utilisation statistics)
                        continue;
                    }
                    const bool run = !sv.done();
                    b.y = yq;
                    sweep<FAM, SW_F>(st, M.tch, M.ncht, M.tch, MAP_TF, lane, b, M.h, s_tab, run, F, dF);
                    if (run) { sv.feed(F - p + b.y); evals++; }
                }
                P = sv.root();
                if (alive && sv.info != 1) fails++;
            } else {
                Newton1 sv;
                // the reference starts at the guess GP's prediction as it is (sympgpr.f90:104-107) even where that GP was
                // trained on P - p (scripts 03/04/05); start_delta adds p, which is the consistent start for such a model
                sv.start(alive ? (a.start_delta ? p + pg : pg) : 0.0, a.newton_max_nb, a.newton_max_b);
                if (!alive) sv.phase = 3;
                for (;;) {
                    const unsigned pend = __ballot_sync(0xffffffffu, !sv.done());
                    if (pend == 0u) break;
                    const double yq = sv.query();
                    double F, dF;
                    if (__popc(pend) <= a.coop_max) {
                        // the last few lanes: one cooperative pass each (1/32 of the arithmetic of a full pass)
                        for (unsigned m = pend; m != 0u; m &= m - 1u) {
                            const int src = __ffs(m) - 1;
                            Pt bs;
                            bs.u = __shfl_sync(0xffffffffu, b.u, src);
                            bs.v = __shfl_sync(0xffffffffu, b.v, src);
                            bs.y = __shfl_sync(0xffffffffu, yq, src);
                            sweep<FAM, SW_F_DF, true>(st, M.tch, M.ncht, M.tch, MAP_TF, lane, bs, M.h, s_tab, true, F, dF);
                            if (lane == src) { sv.feed(F - p + yq, 1.0 + dF); evals++; }
                        }
                        st.passes++;                         // (counted as one pass in the utilisation statistics)
                        continue;
                    }
                    const bool run = !sv.done();
                    b.y = yq;
                    sweep<FAM, SW_F_DF>(st, M.tch, M.ncht, M.tch, MAP_TF, lane, b, M.h, s_tab, run, F, dF);
                    if (run) { sv.feed(F - p + b.y, 1.0 + dF); evals++; }
                }
                P = sv.root();
                if (alive && sv.info != 1 && sv.info != 3) fails++;
            }

            double Pst = P;
            if (alive) {
                if (a.kind == MAP_STANDARD || a.kind == MAP_STANDARD_EXPL) {
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
            sweep<FAM, SW_DQ>(st, M.tch, M.ncht, next_first, first_nf, lane, b, M.h, s_tab, qalive, dq, dummy);
            if (qalive) evals++;
            double qn;
            if (!qalive) qn = nan("");
            else if (a.kind == MAP_HENON || a.kind == MAP_STANDARD_EXPL) qn = dq + q;
            else qn = np_mod(dq + q, TWO_PI);
            if (a.kind == MAP_TOKAMAK_SPLIT && qalive) {
                // loss test at the NEW angle; a lost orbit is NaN in both coordinates (Split_SympGPR/func.py:212-217;
                // compute_r does not depend on the toroidal angle, fieldlines.f90:34-47,94-107)
                const double r = compute_r_dev(Pst * 1e-2, qn, 0.3);
                if (r > 0.5 || Pst < 0.0) { Pst = nan(""); qn = nan(""); }
            }
            if (mine && !alive) pd = nan("");
            q = qn;
            p = Pst;

            if (mine && a.ekind != ENERGY_NONE && (step % a.e_every) == 0) {
                // Welford update with the sample H - H(0); the accumulators live in global memory (L2) between steps:
                // 40 bytes per orbit-step against ~1e5 pair evaluations, and no registers held across the sweeps
                const double x = orbit_energy(a.ekind, a.epar[0], a.epar[1], a.epar[2], q, p) - a.ek[k];
                const double n = (double)(step / a.e_every + 1);
                double mean = a.emean[k];
                const double dl = x - mean;
                mean += dl / n;
                const double m2 = a.em2[k] + dl * (x - mean);
                a.emean[k] = mean;
                a.em2[k] = m2;
                if (step == a.e_every && a.q1) { a.q1[k] = q; a.p1[k] = p; }
                if (step + a.e_every > a.nsteps) {           // last sampled row
                    const double hm = a.ek[k] + mean;
                    a.ehmean[k] = hm;
                    a.eosc[k] = sqrt(m2 / n) / hm;
                }
            }
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
            if (a.pdstate) a.pdstate[k] = pd;
        }
        if (slice + 1 < nslices) {
            // ---- push the batch back: its state is in qfinal/pfinal, progress = slices done ------------
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                a.progress[batch] = (int)slice + 1;
                const unsigned long long pt = (unsigned long long)nbatches + atomicAdd(a.ticket + 2, 1ull);
                unsigned long long* slot = a.slots + (pt % (unsigned long long)nbatches);
                const unsigned long long want = ((pt - (unsigned long long)nbatches + 1ull) << 32) | 0xffffffffull;
                const unsigned long long t0 = globaltimer();
                unsigned n = 0;
                while (ld_acquire_u64(slot) != want) {       // previous occupant of the slot not consumed yet (rare)
                    __nanosleep(100);
                    if ((++n & 255u) == 0u && (*verr != 0ull || globaltimer() - t0 > 20000000000ull)) { atomicExch(a.ticket + 1, 1ull); break; }
                }
                st_release_u64(slot, ((pt + 1ull) << 32) | (unsigned long long)(unsigned)batch);
            }
        }
    }

    // drain the two chunk copies that are still in flight before the shared memory goes away
    mbar_wait(st.bar + (st.seq & 1u), (st.seq >> 1) & 1u);
    mbar_wait(st.bar + ((st.seq + 1u) & 1u), ((st.seq + 1u) >> 1) & 1u);

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        evals += __shfl_xor_sync(0xffffffffu, evals, o);
        fails += __shfl_xor_sync(0xffffffffu, fails, o);
    }
    if (lane == 0) {
        if (evals) atomicAdd(a.stats + 0, evals);
        if (fails) atomicAdd(a.stats + 1, (unsigned long long)fails);
        if (st.passes) atomicAdd(a.ticket + 3, (unsigned long long)st.passes);  // passes over a training set, per warp (utilisation statistics)
    }
}

// ------------------------------------------------------------------------------------------
// model preparation: chunked structure-of-arrays layout, neutral padding (u,v,y) = (0,1,0), alpha = 0
// ------------------------------------------------------------------------------------------
long map_chunks(long n)
{
    long c = (n + MAP_CHUNK - 1) / MAP_CHUNK;
    return c < 2 ? 2 : c;
}

size_t map_model_doubles(long np, long nt)
{
    return (size_t)map_chunks(np) * MAP_GF * MAP_CHUNK + (size_t)map_chunks(nt) * MAP_TF * MAP_CHUNK;
}

size_t map_sched_bytes(long E)
{
    const size_t nb = (size_t)((E + 31) / 32);
    return 4 * sizeof(unsigned long long) + nb * sizeof(unsigned long long) + (nb + 4) * sizeof(int);
}

template <int FAM, int NF>
__global__ void map_prep_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ alpha,
                                long n, long npad, double p, double* __restrict__ out)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    double* ch = out + (i / MAP_CHUNK) * (long)(NF * MAP_CHUNK) + (i % MAP_CHUNK);
    if (i < n) {
        const Pt t = make_pt<FAM>(x[i], y[i], p);
        ch[0] = t.u; ch[MAP_CHUNK] = t.v; ch[2 * MAP_CHUNK] = t.y;
        ch[3 * MAP_CHUNK] = alpha[i];
        if (NF == MAP_TF) ch[4 * MAP_CHUNK] = alpha[n + i];
    } else {
        ch[0] = 0.0; ch[MAP_CHUNK] = 1.0; ch[2 * MAP_CHUNK] = 0.0; ch[3 * MAP_CHUNK] = 0.0;
        if (NF == MAP_TF) ch[4 * MAP_CHUNK] = 0.0;
    }
}

int map_prepare_guess(Ctx& c, int fam, double per, const double* x, const double* y, const double* alpha, long n, double* gch)
{
    const long npad = map_chunks(n) * MAP_CHUNK;
    const unsigned g = (unsigned)((npad + 255) / 256);
    if (fam == FAM_SQ) map_prep_kernel<FAM_SQ, MAP_GF><<<g, 256, 0, c.stream>>>(x, y, alpha, n, npad, per, gch);
    else map_prep_kernel<FAM_PRODUCT, MAP_GF><<<g, 256, 0, c.stream>>>(x, y, alpha, n, npad, per, gch);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

int map_prepare_sympl(Ctx& c, int fam, double per, const double* x, const double* y, const double* alpha, long n, double* tch)
{
    const long npad = map_chunks(n) * MAP_CHUNK;
    const unsigned g = (unsigned)((npad + 255) / 256);
    if (fam == FAM_SQ) map_prep_kernel<FAM_SQ, MAP_TF><<<g, 256, 0, c.stream>>>(x, y, alpha, n, npad, per, tch);
    else map_prep_kernel<FAM_PRODUCT, MAP_TF><<<g, 256, 0, c.stream>>>(x, y, alpha, n, npad, per, tch);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

static int init_exp_table();

// ---------------------------------------------------------------------------------------------------------
// 2-DOF map (BASELINE config 3): prediction side of the 4 x 4-block derivative kernel of dof2.cu.
// NOT IN THE REFERENCE (it reduces Henon-Heiles to a section map with 2 x 2 blocks); parity unpinned, the oracle
// twin is oracle.applymap4.  Generating function F(q1, q2, P1, P2), observations [p - P; Q - q] = grad F:
//     grad_a F(x*) = sig sum_j E_j (at_ja - w_a s_j),   D = x* - x_j,  w_a = D_a / l_a^2,  at = alpha / l^2,
//     s_j = sum_b D_b at_jb,  E_j = exp(-sum_b D_b w_b / 2)                       (row a of build_k4 times alpha)
// One map step: solve  r(P) = P - p + grad_q F(q, P) = 0  (2 equations) by Newton with the analytic 2 x 2 Jacobian
//     d grad_a F / dP_c = -sig sum_j E_j (w_c (at_ja - w_a s_j) + w_a at_jc),   a in q, c in P,
// started at P = p, then  Q = q + grad_P F(q, P).  Warp = 32 orbits in lock step; the training set (8 doubles per
// point: x and at) streams through the warp's double buffer in chunks of 32 points, one bulk copy each.
// ---------------------------------------------------------------------------------------------------------
constexpr int M4_CHUNK = 32;
constexpr int M4_F = 8;

__global__ void map4_prep_kernel(const double* __restrict__ x, const double* __restrict__ alpha, long n, long npad, double gq,
                                 double gP, double* __restrict__ out)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    double* ch = out + (i / M4_CHUNK) * (long)(M4_F * M4_CHUNK) + (i % M4_CHUNK);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        ch[c * M4_CHUNK] = (i < n) ? x[c * n + i] : 0.0;
        ch[(4 + c) * M4_CHUNK] = (i < n) ? alpha[c * n + i] * (c < 2 ? gq : gP) : 0.0;     // padding: alpha = 0
    }
}

template <bool NEWTON>
__device__ __forceinline__ void sweep4(Stream& st, const double* __restrict__ set, int nch, int lane, const double (&xs)[4],
                                       const double (&g)[4], const double* __restrict__ tab, double (&F)[2], double (&J)[4])
{
    constexpr uint32_t bytes = M4_F * M4_CHUNK * sizeof(double);
    F[0] = F[1] = 0.0;
    J[0] = J[1] = J[2] = J[3] = 0.0;
    if (lane == 0) {
        stream_issue(st, st.seq & 1u, set, bytes);
        stream_issue(st, (st.seq + 1u) & 1u, set + M4_F * M4_CHUNK, bytes);
    }
    for (int c = 0; c < nch; c++) {
        const uint32_t slot = st.seq & 1u, par = (st.seq >> 1) & 1u;
        mbar_wait(st.bar + slot, par);
        const double* sm = st.buf + slot * MAP_BUF_DOUBLES;
#pragma unroll 2
        for (int j = 0; j < M4_CHUNK; j++) {
            double D[4], at[4], w[4];
#pragma unroll
            for (int b = 0; b < 4; b++) { D[b] = xs[b] - sm[b * M4_CHUNK + j]; at[b] = sm[(4 + b) * M4_CHUNK + j]; w[b] = D[b] * g[b]; }
            const double e = fma(D[0], w[0], fma(D[1], w[1], fma(D[2], w[2], D[3] * w[3])));
            const double E = exp_neg_tab(-0.5 * e, tab);
            const double sj = fma(D[0], at[0], fma(D[1], at[1], fma(D[2], at[2], D[3] * at[3])));
            if (NEWTON) {
                const double t0 = fma(-w[0], sj, at[0]), t1 = fma(-w[1], sj, at[1]);
                F[0] = fma(E, t0, F[0]);
                F[1] = fma(E, t1, F[1]);
                J[0] = fma(E, fma(w[2], t0, w[0] * at[2]), J[0]);       // d F_0 / d P_1
                J[1] = fma(E, fma(w[3], t0, w[0] * at[3]), J[1]);       // d F_0 / d P_2
                J[2] = fma(E, fma(w[2], t1, w[1] * at[2]), J[2]);       // d F_1 / d P_1
                J[3] = fma(E, fma(w[3], t1, w[1] * at[3]), J[3]);       // d F_1 / d P_2
            } else {
                F[0] = fma(E, fma(-w[2], sj, at[2]), F[0]);
                F[1] = fma(E, fma(-w[3], sj, at[3]), F[1]);
            }
        }
        __syncwarp();
        if (lane == 0 && c + 2 < nch) stream_issue(st, slot, set + (size_t)(c + 2) * M4_F * M4_CHUNK, bytes);
        st.seq++;
    }
    st.passes++;
}

struct Map4Args {
    const double* set; int nch;          // chunked training set
    double g[4], sig;
    long E, nsteps;
    const double *q0, *p0;               // (2, E) each
    double *qout, *pout;                 // history (rows, 2, E) or nullptr
    long out_every;
    double *qfinal, *pfinal;             // (2, E)
    unsigned long long* stats;           // [0] Newton evaluations, [1] exits without convergence
};

__global__ void __launch_bounds__(MAP_THREADS, 3) map4_kernel(Map4Args a)
{
    __shared__ __align__(16) double s_buf[MAP_WARPS][2 * MAP_BUF_DOUBLES];
    __shared__ unsigned long long s_bar[MAP_WARPS][2];
    __shared__ double s_tab_all[64 * 32];
    for (int i = threadIdx.x; i < 64 * 32; i += MAP_THREADS) s_tab_all[i] = c_exp2_tab[i >> 5];
    const double* s_tab = s_tab_all + (threadIdx.x & 31);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Stream st;
    st.buf = s_buf[warp]; st.bar = s_bar[warp]; st.seq = 0; st.passes = 0u; st.primed = nullptr;
    if (lane == 0) { mbar_init(st.bar + 0, 1); mbar_init(st.bar + 1, 1); mbar_fence_init(); }
    __syncwarp();
    const double g[4] = {a.g[0], a.g[1], a.g[2], a.g[3]};
    unsigned long long evals = 0ull;
    unsigned int fails = 0u;
    const long nbatches = (a.E + 31) / 32;
    const long nwarps = (long)gridDim.x * MAP_WARPS;
    for (long batch = (long)blockIdx.x * MAP_WARPS + warp; batch < nbatches; batch += nwarps) {
        const long k = batch * 32 + lane;
        const bool mine = k < a.E;
        double q[2] = {0.0, 0.0}, p[2] = {0.0, 0.0};
        if (mine) { q[0] = a.q0[k]; q[1] = a.q0[a.E + k]; p[0] = a.p0[k]; p[1] = a.p0[a.E + k]; }
        if (mine && a.out_every > 0) {
            a.qout[k] = q[0]; a.qout[a.E + k] = q[1]; a.pout[k] = p[0]; a.pout[a.E + k] = p[1];
        }
        for (long step = 1; step <= a.nsteps; step++) {
            bool alive = mine && q[0] == q[0] && q[1] == q[1] && p[0] == p[0] && p[1] == p[1];
            double P[2] = {p[0], p[1]};
            bool done = !alive;
            int it = 0;
            while (__any_sync(0xffffffffu, !done)) {
                const double xs[4] = {alive ? q[0] : 0.0, alive ? q[1] : 0.0, done ? 0.0 : P[0], done ? 0.0 : P[1]};
                double F[2], J[4];
                sweep4<true>(st, a.set, a.nch, lane, xs, g, s_tab, F, J);
                if (!done) {
                    evals++;
                    it++;
                    const double r0 = P[0] - p[0] + a.sig * F[0], r1 = P[1] - p[1] + a.sig * F[1];
                    const double j00 = 1.0 - a.sig * J[0], j01 = -a.sig * J[1], j10 = -a.sig * J[2], j11 = 1.0 - a.sig * J[3];
                    const double det = j00 * j11 - j01 * j10;
                    const double d0 = (j11 * r0 - j01 * r1) / det, d1 = (j00 * r1 - j10 * r0) / det;
                    if (!(fabs(d0) <= DBL_MAX) || !(fabs(d1) <= DBL_MAX)) {     // singular Jacobian / non-finite residual
                        P[0] = P[1] = nan(""); done = true; fails++;
                    } else {
                        P[0] -= d0; P[1] -= d1;
                        const double ad = fmax(fabs(d0), fabs(d1));
                        if (ad <= 1e-9 * fmax(fmax(fabs(P[0]), fabs(P[1])), 1.0)) done = true;      // next error O(ad^2)
                        else if (it >= 40) { done = true; fails++; }
                    }
                }
            }
            alive = alive && P[0] == P[0] && P[1] == P[1];
            {
                const double xs[4] = {alive ? q[0] : 0.0, alive ? q[1] : 0.0, alive ? P[0] : 0.0, alive ? P[1] : 0.0};
                double F[2], J[4];
                sweep4<false>(st, a.set, a.nch, lane, xs, g, s_tab, F, J);
                if (alive) {
                    evals++;
                    q[0] += a.sig * F[0]; q[1] += a.sig * F[1];
                    p[0] = P[0]; p[1] = P[1];
                } else if (mine) {
                    q[0] = q[1] = p[0] = p[1] = nan("");
                }
            }
            if (mine && a.out_every > 0 && (step % a.out_every) == 0) {
                const long row = step / a.out_every;
                double* qo = a.qout + row * 2 * a.E;
                double* po = a.pout + row * 2 * a.E;
                qo[k] = q[0]; qo[a.E + k] = q[1]; po[k] = p[0]; po[a.E + k] = p[1];
            }
        }
        if (mine) { a.qfinal[k] = q[0]; a.qfinal[a.E + k] = q[1]; a.pfinal[k] = p[0]; a.pfinal[a.E + k] = p[1]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        evals += __shfl_xor_sync(0xffffffffu, evals, o);
        fails += __shfl_xor_sync(0xffffffffu, fails, o);
    }
    if (lane == 0) {
        if (evals) atomicAdd(a.stats + 0, evals);
        if (fails) atomicAdd(a.stats + 1, (unsigned long long)fails);
    }
}

long map4_chunks(long n)
{
    long c = (n + M4_CHUNK - 1) / M4_CHUNK;
    return c < 2 ? 2 : c;
}

// x (4, n), alpha (4, n) device arrays -> chunked set (map4_chunks(n) * 8 * 32 doubles) -> nm - 1 steps of E orbits
int map4_run(Ctx& c, const double* d_x, const double* d_alpha, long n, double lq, double lP, double sig, double* d_set, long E,
             long nsteps, const double* d_q0, const double* d_p0, double* d_qhist, double* d_phist, long out_every, double* d_qfinal,
             double* d_pfinal, unsigned long long* d_stats)
{
    SGP_TRY(init_exp_table());
    const long nch = map4_chunks(n), npad = nch * M4_CHUNK;
    const double gq = 1.0 / (lq * lq), gP = 1.0 / (lP * lP);
    map4_prep_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c.stream>>>(d_x, d_alpha, n, npad, gq, gP, d_set);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    if (E <= 0) return ST_OK;
    Map4Args a;
    a.set = d_set; a.nch = (int)nch;
    a.g[0] = a.g[1] = gq; a.g[2] = a.g[3] = gP; a.sig = sig;
    a.E = E; a.nsteps = nsteps; a.q0 = d_q0; a.p0 = d_p0;
    a.qout = d_qhist; a.pout = d_phist; a.out_every = (d_qhist && d_phist) ? out_every : 0;
    a.qfinal = d_qfinal; a.pfinal = d_pfinal; a.stats = d_stats;
    const long nbatches = (E + 31) / 32;
    long blocks = (nbatches + MAP_WARPS - 1) / MAP_WARPS;
    const long cap = (long)(c.sm_count > 0 ? c.sm_count : 148) * 3;
    if (blocks > cap) blocks = cap;
    map4_kernel<<<(unsigned)blocks, MAP_THREADS, 0, c.stream>>>(a);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

// StandardMapIterate python/04_standard_map/main.py:27-39 (training / reference orbits of the standard map, no wrap):
// f (2, N, nm) C order, f[:, i, 0] = X0[:, i], J' = J + k sin(th), th' = th + J'.  One thread per orbit.
__global__ void standard_map_iterate_kernel(double kk, long nm, long N, const double* __restrict__ X0, double* __restrict__ f)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double th = X0[i], J = X0[N + i];
    double* fth = f + i * nm;
    double* fJ = f + (N + i) * nm;
    fth[0] = th; fJ[0] = J;
    for (long l = 1; l < nm; l++) {
        J = J + kk * sin(th);
        th = th + J;
        fth[l] = th; fJ[l] = J;
    }
}

int standard_map_iterate(Ctx& c, double kk, long nm, long N, const double* d_X0, double* d_f)
{
    if (N <= 0 || nm <= 0) return ST_OK;
    standard_map_iterate_kernel<<<(unsigned)((N + 127) / 128), 128, 0, c.stream>>>(kk, nm, N, d_X0, d_f);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

static int init_exp_table()
{
    static bool done[64] = {};
    int dev = 0;
    SGP_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || done[dev]) return ST_OK;
    double tab[64];
    for (int j = 0; j < 64; j++) tab[j] = (double)exp2l((long double)j / 64.0L);
    SGP_CUDA(cudaMemcpyToSymbol(c_exp2_tab, tab, sizeof(tab)));
    done[dev] = true;
    return ST_OK;
}

int map_launch(Ctx& c, int fam, int solver, MapArgs a, void* sched)
{
    if (a.E <= 0) return ST_OK;
    SGP_TRY(init_exp_table());
    const long nbatches = (a.E + 31) / 32;
    SGP_CUDA(cudaMemsetAsync(sched, 0, map_sched_bytes(a.E), c.stream));
    a.ticket = (unsigned long long*)sched;
    a.slots = a.ticket + 4;
    a.progress = (int*)(a.slots + nbatches);
    // work items of about 1/16 of the step loop, 1..16 steps each: enough items to balance the tail
    long ss = a.nsteps / 16;
    if (ss < 1) ss = 1;
    if (ss > 16) ss = 16;
    if (a.nmodels > 1) ss = (ss + a.nmodels - 1) / a.nmodels * a.nmodels;      // work items start at a turn boundary of a split map
    a.slice_steps = ss;
    a.start_delta = (solver == 3) ? 1 : 0;
    a.newton_max_nb = 30; a.newton_max_b = 60;
    a.coop_max = MAP_COOP_MAX;
    {   // measurement overrides
        static const char* e_nb = getenv("SGP_NEWTON_MAX_NB");
        static const char* e_b = getenv("SGP_NEWTON_MAX_B");
        static const char* e_cm = getenv("SGP_MAP_COOP_MAX");
        if (e_cm) a.coop_max = atoi(e_cm);
        if (e_nb && atoi(e_nb) > 0) a.newton_max_nb = atoi(e_nb);
        if (e_b && atoi(e_b) > 0) a.newton_max_b = atoi(e_b);
    }
    const long warps_needed = nbatches;
    long blocks = (warps_needed + MAP_WARPS - 1) / MAP_WARPS;
    const long cap = (long)(c.sm_count > 0 ? c.sm_count : 148) * MAP_BLOCKS_PER_SM;
    if (blocks > cap) blocks = cap;
    const unsigned g = (unsigned)blocks;
#define ML(F, S) map_kernel<F, S><<<g, MAP_THREADS, 0, c.stream>>>(a)
    if (solver == 0) {
        switch (fam) {
        case FAM_PRODUCT: ML(FAM_PRODUCT, 0); break;
        case FAM_SQ: ML(FAM_SQ, 0); break;
        case FAM_SUM: ML(FAM_SUM, 0); break;
        default: set_error("unknown kernel family %d", fam); return ST_BADARG;
        }
    } else if (solver == 2) {
        switch (fam) {
        case FAM_PRODUCT: ML(FAM_PRODUCT, 2); break;
        case FAM_SQ: ML(FAM_SQ, 2); break;
        case FAM_SUM: ML(FAM_SUM, 2); break;
        default: set_error("unknown kernel family %d", fam); return ST_BADARG;
        }
    } else if (solver == 1 || solver == 3) {
        switch (fam) {
        case FAM_PRODUCT: ML(FAM_PRODUCT, 1); break;
        case FAM_SQ: ML(FAM_SQ, 1); break;
        case FAM_SUM: ML(FAM_SUM, 1); break;
        default: set_error("unknown kernel family %d", fam); return ST_BADARG;
        }
    } else {
        set_error("unknown solver %d", solver); return ST_BADARG;
    }
#undef ML
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

}  // namespace sgp
