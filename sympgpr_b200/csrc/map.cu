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
// warp (__any_sync), so a slow root solve delays 31 neighbours, not a thread block; slicing the step
// loop into work items keeps all SMs busy to the end whatever the ensemble size.  The kernel is bound
// by the FP64 pipe (about 50 DP instructions per orbit-point pair incl. one exp), not by shared
// memory, L2 or HBM: a chunk of 2.5 KB feeds 32 x 64 pair evaluations.
//
// Summation order is fixed (chunk by chunk, two interleaved partial sums), so results do not depend
// on scheduling.
#include "map.cuh"

#include "hybrd.cuh"
#include "mbar.cuh"

namespace sgp {

constexpr int MAP_THREADS = 128;
constexpr int MAP_WARPS = MAP_THREADS / 32;
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

enum SweepMode : int { SW_GUESS = 0, SW_F_DF = 1, SW_F = 2, SW_DQ = 3 };

// per-warp chunk stream: two buffers, one mbarrier each; `seq` counts the chunks consumed so far
struct Stream {
    double* buf;                 // 2 * MAP_BUF_DOUBLES
    unsigned long long* bar;     // 2
    uint32_t seq;
};

__device__ __forceinline__ void stream_issue(const Stream& s, uint32_t slot, const double* src, uint32_t bytes)
{
    mbar_arrive_expect_tx(s.bar + slot, bytes);
    bulk_g2s(s.buf + slot * MAP_BUF_DOUBLES, src, bytes, s.bar + slot);
}

// One sweep of the warp over a whole training set (nch chunks of NF fields); the first two chunks are
// already in flight.  While chunk c is evaluated, chunk c+2 is fetched -- for the last two chunks that
// is chunk 0/1 of the set the NEXT sweep will read (nxt, NXF fields).
template <int FAM, int MODE>
__device__ __forceinline__ void sweep(Stream& st, const double* __restrict__ set, int nch, const double* __restrict__ nxt, int nxf,
                                      int lane, const Pt& b, const HypC& h, bool active, double& o0, double& o1)
{
    constexpr int NF = (MODE == SW_GUESS) ? MAP_GF : MAP_TF;
    double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
    const bool any = __any_sync(0xffffffffu, active);
    for (int c = 0; c < nch; c++) {
        const uint32_t slot = st.seq & 1u, par = (st.seq >> 1) & 1u;
        mbar_wait(st.bar + slot, par);
        const double* sm = st.buf + slot * MAP_BUF_DOUBLES;
        if (any) {
#pragma unroll 2
            for (int j = 0; j < MAP_CHUNK; j += 2) {
                const double2 u2 = *reinterpret_cast<const double2*>(sm + 0 * MAP_CHUNK + j);
                const double2 v2 = *reinterpret_cast<const double2*>(sm + 1 * MAP_CHUNK + j);
                const double2 y2 = *reinterpret_cast<const double2*>(sm + 2 * MAP_CHUNK + j);
                const double2 a2 = *reinterpret_cast<const double2*>(sm + 3 * MAP_CHUNK + j);
                Pt p0, p1;
                p0.u = u2.x; p0.v = v2.x; p0.y = y2.x;
                p1.u = u2.y; p1.v = v2.y; p1.y = y2.y;
                const Pair<FAM> q0(p0, b, h), q1(p1, b, h);
                if (MODE == SW_GUESS) {
                    s0 += q0.k() * a2.x;
                    s1 += q1.k() * a2.y;
                } else {
                    const double2 c2 = *reinterpret_cast<const double2*>(sm + 4 * MAP_CHUNK + j);
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
        __syncwarp();                                        // every lane is done with this buffer
        if (lane == 0) {
            if (c + 2 < nch) stream_issue(st, slot, set + (size_t)(c + 2) * NF * MAP_CHUNK, NF * MAP_CHUNK * sizeof(double));
            else stream_issue(st, slot, nxt + (size_t)(c + 2 - nch) * nxf * MAP_CHUNK, nxf * MAP_CHUNK * sizeof(double));
        }
        st.seq++;
    }
    o0 = h.sig * (s0 + s1);
    o1 = h.sig * (t0 + t1);
}

template <int FAM, int SOLVER>
__global__ void __launch_bounds__(MAP_THREADS, MAP_BLOCKS_PER_SM)
map_kernel(MapArgs a)
{
    __shared__ __align__(16) double s_buf[MAP_WARPS][2 * MAP_BUF_DOUBLES];
    __shared__ unsigned long long s_bar[MAP_WARPS][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Stream st;
    st.buf = s_buf[warp];
    st.bar = s_bar[warp];
    st.seq = 0;
    if (lane == 0) {
        mbar_init(st.bar + 0, 1);
        mbar_init(st.bar + 1, 1);
        mbar_fence_init();
    }
    __syncwarp();
    // prime the stream with the first two chunks of the guess set (every work item starts with a guess sweep)
    if (lane == 0) {
        stream_issue(st, 0, a.gch, MAP_GF * MAP_CHUNK * sizeof(double));
        stream_issue(st, 1, a.gch + MAP_GF * MAP_CHUNK, MAP_GF * MAP_CHUNK * sizeof(double));
    }

    const long nbatches = (a.E + 31) / 32;
    const long nslices = a.nsteps > 0 ? (a.nsteps + a.slice_steps - 1) / a.slice_steps : 1;
    const unsigned long long nitems = (unsigned long long)nbatches * (unsigned long long)nslices;
    unsigned long long evals = 0ull;
    unsigned int fails = 0u;
    volatile unsigned long long* verr = a.ticket + 1;      // scheduler error word (dependency wait timed out)

    for (;;) {
        unsigned long long tk = 0ull;
        if (lane == 0) tk = atomicAdd(a.ticket, 1ull);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= nitems) break;
        const long batch = (long)(tk % (unsigned long long)nbatches);
        const long slice = (long)(tk / (unsigned long long)nbatches);
        // the previous slice of these orbits must be finished (its state is in qfinal/pfinal)
        if (slice > 0) {
            int ok = 1;
            if (lane == 0) {
                const unsigned long long t0 = globaltimer();
                unsigned n = 0;
                while (ld_acquire(a.slice_done + batch) < (int)slice) {
                    __nanosleep(200);
                    if ((++n & 255u) == 0u && (*verr != 0ull || globaltimer() - t0 > 20000000000ull)) { ok = 0; break; }
                }
                if (!ok) atomicExch(a.ticket + 1, 1ull);
            }
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (!ok) break;
        }
        const long k = batch * 32 + lane;
        const bool mine = k < a.E;
        double q = 0.0, p = 0.0, pd = 0.0;
        if (mine) {
            if (slice == 0) { q = a.q0[k]; p = a.p0[k]; pd = p; }
            else { q = a.qfinal[k]; p = a.pfinal[k]; pd = a.pdstate ? a.pdstate[k] : 0.0; }
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
            // an orbit that is already NaN stays NaN (tokamak: explicit test, func.py:192-193; the
            // other variants propagate it through the arithmetic)
            bool alive = mine && (q == q) && (p == p);
            Pt b;
            if (alive) { b = make_pt<FAM>(q, p, a.h.p); } else { b.u = 0; b.v = 1; b.y = 0; }

            double pg, dummy;
            sweep<FAM, SW_GUESS>(st, a.gch, a.nchg, a.tch, MAP_TF, lane, b, a.hp, alive, pg, dummy);
            if (alive && !(fabs(pg) <= DBL_MAX)) alive = false;

            double P;
            if (SOLVER == 0) {
                Hybrd1 sv;
                sv.start(alive ? pg : 0.0);
                if (!alive) sv.phase = 3;
                while (__any_sync(0xffffffffu, !sv.done())) {
                    const bool run = !sv.done();
                    b.y = sv.query();
                    double F, dF;
                    sweep<FAM, SW_F>(st, a.tch, a.ncht, a.tch, MAP_TF, lane, b, a.h, run, F, dF);
                    if (run) { sv.feed(F - p + b.y); evals++; }
                }
                P = sv.root();
                if (alive && sv.info != 1) fails++;
            } else {
                Newton1 sv;
                sv.start(alive ? pg : 0.0);
                if (!alive) sv.phase = 3;
                while (__any_sync(0xffffffffu, !sv.done())) {
                    const bool run = !sv.done();
                    b.y = sv.query();
                    double F, dF;
                    sweep<FAM, SW_F_DF>(st, a.tch, a.ncht, a.tch, MAP_TF, lane, b, a.h, run, F, dF);
                    if (run) { sv.feed(F - p + b.y, 1.0 + dF); evals++; }
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
            sweep<FAM, SW_DQ>(st, a.tch, a.ncht, a.gch, MAP_GF, lane, b, a.h, qalive, dq, dummy);
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
            if (a.pdstate) a.pdstate[k] = pd;
        }
        if (slice + 1 < nslices) {
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release(a.slice_done + batch, (int)slice + 1);
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
    return 16 + (nb + 4) * sizeof(int);
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

int map_launch(Ctx& c, int fam, int solver, MapArgs a, void* sched)
{
    if (a.E <= 0) return ST_OK;
    const long nbatches = (a.E + 31) / 32;
    SGP_CUDA(cudaMemsetAsync(sched, 0, map_sched_bytes(a.E), c.stream));
    a.ticket = (unsigned long long*)sched;
    a.slice_done = (int*)((char*)sched + 16);
    // work items of about 1/16 of the step loop, 1..16 steps each: enough items to balance the tail
    long ss = a.nsteps / 16;
    if (ss < 1) ss = 1;
    if (ss > 16) ss = 16;
    a.slice_steps = ss;
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
