// Left-looking tile Cholesky as ONE persistent cooperative kernel (the whole potrf is one launch).
//
// Replaces scipy.linalg.cholesky (LAPACK dpotrf) of python/05_tokamak/SympGPR/func.py:147 and
// python/02_pert_pendulum/func.py:136,195 on the 2N x 2N Hessian-block matrix.
//
// The matrix is cut into 128 x 128 tiles.  Tile (i,j), i >= j, is one task:
//     C'   = A(i,j) - sum_{k<j} L(i,k) L(j,k)^T      one register-accumulated DMMA GEMM of depth 128 j
//     i==j : L(j,j) = chol(C'), Dinv_j = L(j,j)^-1    in shared memory by the 8 consumer warps
//     i> j : L(i,j) = C' Dinv_j^T                     a second, depth-128 pass through the same ring
// so every flop of the update runs in a long-k main loop (no read-modify-write of the trailing matrix,
// no short-k GEMM launches, no wave quantisation) and the only global traffic is the operand stream.
// CTAs (one per SM, warp-specialised exactly like dmma_gemm_ws.cuh: producer warp + full/empty
// mbarrier ring + 8 DMMA consumer warps) take tasks round-robin from a fixed order
//     DIAG(0) | (1,0) DIAG(1) (2,0) ... (nt-1,0) | (2,1) DIAG(2) (3,1) ... | ...
// i.e. column by column, with the next column's diagonal tile pulled forward to second place: its
// factorisation (the latency-bound step) then overlaps the rest of the current column.  A task only
// depends on tasks earlier in that order, all CTAs are co-resident (cooperative launch), so the CTA
// holding the oldest unfinished task can always proceed.  Dependencies are tracked with one ready
// flag per tile in global memory (release store after the tile is written, acquire load by the
// producer lane before it streams the tile).
//
// A failed dependency wait (10 s) sets an abort word; every wait in the kernel returns early once it
// is set and the launcher reports an error -- the kernel cannot hang the device.
#include "chol.cuh"

#include <stdlib.h>
#include <vector>

namespace sgp {

namespace {

constexpr int LL_LD = TILE + 1;                 // row stride of the diagonal tile in shared memory
constexpr int LL_CONSUMERS = WS_CONSUMER_WARPS * 32;

struct LLArgs {
    double* A; long lda; int nt;
    double* Dinv;          // nt tiles, ld 128
    double* logparts;      // nt
    int* info;
    int* ready;            // nt*nt, ready[i + j*nt] = 1 once L(i,j) (and Dinv_j for i == j) is final
    int* slab;             // nt*8: slab[j*8 + s] counts the warps (2) that have stored columns [16 s, 16 s + 16) of the
                           // SUB-DIAGONAL tile L(j+1, j): the next diagonal task streams that tile slab by slab
    int* abort;
    // optional fused forward substitution  L w = y  (potrs, first half): the diagonal task of column j also
    // accumulates v_j = sum_{k<j} L(j,k) w_k from the slabs it streams anyway, then w_j = Dinv_j (y_j - v_j)
    const double* y;       // n_pad (zero in the padding) or nullptr
    double* w;             // n_pad
    unsigned long long* dbg;   // optional (SGP_LL_TRACE=1): 8 globaltimer stamps per task [task*8 + k], else nullptr
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(LL_CONSUMERS) : "memory"); }

// one lane spins until the tile's ready flag is set (or abort)
__device__ __forceinline__ void wait_ready(const int* flag, int* abort, int* info)
{
    if (ld_acquire(flag) != 0) return;
    const unsigned long long t0 = globaltimer();
    unsigned n = 0;
    while (ld_acquire(flag) == 0) {
        __nanosleep(100);
        if ((++n & 255u) == 0u) {
            if (*(volatile int*)abort) return;
            if (globaltimer() - t0 > 10000000000ull) { atomicExch(abort, 1); atomicExch(info, -1); return; }
        }
    }
}

// one lane spins until a counter has reached `need` (or abort)
__device__ __forceinline__ void wait_count(const int* cnt, int need, int* abort, int* info)
{
    if (ld_acquire(cnt) >= need) return;
    const unsigned long long t0 = globaltimer();
    unsigned n = 0;
    while (ld_acquire(cnt) < need) {
        __nanosleep(40);
        if ((++n & 255u) == 0u) {
            if (*(volatile int*)abort) return;
            if (globaltimer() - t0 > 10000000000ull) { atomicExch(abort, 1); atomicExch(info, -1); return; }
        }
    }
}

// task order (see the header comment): ticket -> tile (i,j)
__device__ __forceinline__ void ll_ticket(long t, int nt, int& i, int& j)
{
    if (t == 0) { i = 0; j = 0; return; }
    const long tp = t - 1;
    const double b = 2.0 * nt + 1.0;
    long c = (long)((b - sqrt(b * b - 8.0 * (double)tp)) * 0.5);
    if (c < 0) c = 0;
    if (c > nt - 2) c = nt - 2;
    auto off = [nt](long q) { return q * nt - q * (q - 1) / 2; };
    while (c < nt - 2 && off(c + 1) <= tp) c++;
    while (c > 0 && off(c) > tp) c--;
    const long r = tp - off(c);
    if (r == 0) { i = (int)c + 1; j = (int)c; }
    else if (r == 1) { i = (int)c + 1; j = (int)c + 1; }
    else { i = (int)(c + r); j = (int)c; }
}

// producer: nk slabs of (A rows at pa, B rows at pb), both LAYOUT_MN, k advancing by GK columns
__device__ __forceinline__ void ll_produce(const double* pa, long lda, const double* pb, long ldb, int nk, uint32_t& it, double* smem,
                                           unsigned long long* full, unsigned long long* empty, int lane, const volatile int* abort)
{
    for (int kb = 0; kb < nk; kb++, it++) {
        const int s = (int)(it % WS_STAGES);
        const uint32_t ph = (it / WS_STAGES) & 1u;
        mbar_wait_ab(empty + s, ph ^ 1u, abort);
        if (lane == 0) mbar_arrive_expect_tx(full + s, 2u * WS_SLAB_BYTES);
        __syncwarp();
        double* sa = smem + (size_t)s * 2 * STAGE_DOUBLES;
        const long kk = (long)kb * GK;
        produce_slab<LAYOUT_MN>(sa, pa, lda, 0, kk, lane, full + s);
        produce_slab<LAYOUT_MN>(sa + STAGE_DOUBLES, pb, ldb, 0, kk, lane, full + s);
    }
}

// consumers: acc += sum over nk slabs
// MATVEC: additionally v += A_slab * wv[k] for the row (tid & 127) of the A slab, k-half (tid >> 7)
template <bool MATVEC>
__device__ __forceinline__ void ll_consume(double (&acc)[8][4][2], int nk, uint32_t& it, const double* smem, unsigned long long* full,
                                           unsigned long long* empty, int wm, int wn, int g, int t, int lane, const volatile int* abort,
                                           const double* wv = nullptr, int tid = 0, double* vout = nullptr)
{
    double v = 0.0;
    for (int kb = 0; kb < nk; kb++, it++) {
        const int s = (int)(it % WS_STAGES);
        const uint32_t ph = (it / WS_STAGES) & 1u;
        mbar_wait_ab(full + s, ph, abort);
        const double* sa = smem + (size_t)s * 2 * STAGE_DOUBLES;
        const double* sb = sa + STAGE_DOUBLES;
        if (MATVEC) {
            const int m = tid & (TILE - 1), kh = (tid >> 7) * (GK / 2);
            const double* wk = wv + (long)kb * GK + kh;
#pragma unroll
            for (int kk = 0; kk < GK / 2; kk++) v = fma(sa[(kh + kk) * LDMN + m], __ldcg(wk + kk), v);
        }
        double af[2][8], bf[2][4];
        load_frags_a<LAYOUT_MN>(sa, wm, g, t, 0, af[0]);
        load_frags_b<LAYOUT_MN>(sb, wn, g, t, 0, bf[0]);
#pragma unroll
        for (int k4 = 0; k4 < GK / 4; k4++) {
            const int cur = k4 & 1, nxt = cur ^ 1;
            if (k4 + 1 < GK / 4) {
                load_frags_a<LAYOUT_MN>(sa, wm, g, t, (k4 + 1) * 4, af[nxt]);
                load_frags_b<LAYOUT_MN>(sb, wn, g, t, (k4 + 1) * 4, bf[nxt]);
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    if (MATVEC) *vout = v;
}

// ---------------------------------------------------------------------------------------------------------
// Diagonal tile: Cholesky factor L and inverse X = L^-1 of a 128 x 128 block in shared memory, blocked 4 x 4
// with 32 x 32 blocks so that block-wide barriers are per PHASE (about 30), not per column (the unblocked
// version took 123 us + 138 us and paced the whole factorisation for n <= 16k; tools/ll_trace.py).
//   layout: S[r*LL_LD + c]; lower triangle (r >= c): the matrix, then L.  The inverse is kept TRANSPOSED in the
//   strict upper triangle: X(r,c) = S[c*LL_LD + r] for r > c, diagonal X(c,c) = xd[c].
//   per block column kb:  warp 0 factors the 32 x 32 diagonal block and inverts it (warp-synchronous);
//                         all: panel L(i, kb) = A(i, kb) X_kk^T; trailing A(i,c) -= L(i,kb) L(c,kb)^T
//   then per block row bi = 1..3:  T = sum_k L(bi,k) X(k, 0..bi-1),  X(bi, 0..bi-1) = -X_bibi T.
// ---------------------------------------------------------------------------------------------------------
constexpr int DB = 32;

// unroll factor of the 16-step column loops of the two warp-level 32 x 32 routines below (measured: see DESIGN.md 4)
#ifndef LL_UNR
#define LL_UNR 2
#endif
#define LL_STR2(x) #x
#define LL_STR(x) LL_STR2(x)
#define LL_UNROLL _Pragma(LL_STR(unroll LL_UNR))


// warp 0 only.  Factor the diagonal block at k0 in place (lower), write its inverse transposed into the block's strict
// upper triangle and 1/l_cc into xd[].  Register-resident: lane = row, the row lives in 32 registers, columns
// travel by shuffles; no shared-memory round trip on the dependent chain.  Compact code (two rolled loops of 32
// steps; a fully unrolled triangular version is 25 000 instructions and runs at instruction-fetch speed) and
// branch-free (selects, so the shuffles need no re-convergence).  __noinline__: own register allocation.
//   factor : the register array is shifted by one per column, so the current column is always row[0]:
//            d = row_j[0]; l = row[0] / sqrt(d); row[c-1] = row[c] - l * L(j+c, j)
//   inverse: lane r accumulates row r of X = L^-1 in x[0..31] (x = e_r at the start);
//            step k: lane k scales its finished row by 1/l_kk; lanes r > k: x[c] -= L(r,k) X(k,c)
__device__ __noinline__ void warp_factor32(double* S, int k0, double* xd, int* info, int col0, double* lbuf)
{
    // Column j of the block travels through shared memory instead of 31 double shuffles (62 SHFL per column bound the old
    // version at ~480 clocks per column): every lane stores its L(r, j) into a 32-double line (two lines, alternating), one
    // __syncwarp, then the line is read back as broadcast 16-byte loads.  Fully unrolled (about 1500 instructions), the row
    // stays in registers with fixed indices, same operations and order as before: bit-identical results.
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    double* Sr = S + (k0 + lane) * LL_LD + k0;
    double row[DB];
#pragma unroll
    for (int c = 0; c < DB; c++) row[c] = (c <= lane) ? Sr[c] : 0.0;
    double myinv = 1.0;
    int bad = 0;
#pragma unroll
    for (int j = 0; j < DB; j++) {
        double d = __shfl_sync(FULL, row[j], j);
        const bool neg = !(d > 0.0);                     // also NaN
        bad = (neg && bad == 0) ? j + 1 : bad;
        d = neg ? 1.0 : d;
        const double inv = rsqrt(d);
        const double lj = (lane == j) ? d * inv : row[j] * inv;
        myinv = (lane == j) ? inv : myinv;
        Sr[j] = (lane >= j) ? lj : Sr[j];                // (plain select + store: no divergent branch)
        if (j < DB - 1) {
            double* line = lbuf + (j & 1) * DB;
            line[lane] = lj;
            __syncwarp();
            int c = j + 1;
            if (c & 1) {
                row[c] = fma(-lj, line[c], row[c]);
                c++;
            }
#pragma unroll
            for (; c < DB; c += 2) {
                const double2 l2 = *reinterpret_cast<const double2*>(line + c);         // L(c, j), L(c + 1, j)
                row[c] = fma(-lj, l2.x, row[c]);
                row[c + 1] = fma(-lj, l2.y, row[c + 1]);
            }
        }
    }
    if (bad != 0 && lane == 0) atomicCAS(info, 0, col0 + k0 + bad);
    xd[k0 + lane] = myinv;
    __syncwarp();
}

// One warp, after warp_factor32: inverse of the factored diagonal block (reads L and xd from shared memory, writes X^T into the
// block's strict upper triangle).  Runs while another warp factors the next block / the other warps form the off-diagonal part
// of the tile inverse.  Lane c solves L x = e_c for column c of X by forward substitution, on its own: no shuffles, no
// synchronisation -- every L(r, k) is a broadcast load (all lanes read the same address), the column lives in 32 registers,
// rows above the diagonal come out as exact zeros by themselves.  496 FMAs per lane on four partial sums per row, fully
// unrolled (the shuffle version that accumulated ROWS of X took 7.8 us per block and set the pace of the factorisation phases).
__device__ __noinline__ void warp_invert32(double* S, int k0, const double* xd)
{
    const int lane = threadIdx.x & 31;
    const double* L0 = S + k0 * LL_LD + k0;                  // L(r, k) = L0[r * LL_LD + k]
    double x[DB];
#pragma unroll
    for (int r = 0; r < DB; r++) {
        const double* Lr = L0 + r * LL_LD;
        double s[4] = {(r == lane) ? 1.0 : 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < r; k++) {
            const int slot = (k == r - 1) ? 0 : 1 + (k % 3);     // the newest x joins the chain that ends the row
            s[slot] = fma(-Lr[k], x[k], s[slot]);
        }
        x[r] = ((s[1] + s[2]) + s[3] + s[0]) * xd[k0 + r];
    }
    // X(r, c), r > c, transposed into the strict upper triangle of the block: lane c owns row k0 + c of S
    double* Sc = S + (k0 + lane) * LL_LD + k0;
#pragma unroll
    for (int r = 1; r < DB; r++)
        if (r > lane) Sc[r] = x[r];
    __syncwarp();
}

// named barrier over the first NWARPS... a group of GROUP threads (all of them must call it): id 1 = the 8 consumer warps
template <int BARID, int GROUP>
__device__ __forceinline__ void group_bar() { asm volatile("bar.sync %0, %1;\n" ::"n"(BARID), "n"(GROUP) : "memory"); }

// Off-diagonal blocks of the tile inverse by DMMA, block rows bi0 .. bi1, by a group of NWARPS warps (wid = index of the
// calling warp inside the group, BARID = the group's named barrier; 8 x 8 output tiles handed out round-robin).  For block
// row bi (rows r0 = 32 bi ...), with X known for the rows above and for the diagonal block bi:
//     T(rr, c)      =  sum_{k = c .. r0-1} L(r0+rr, k) X(k, c)           (32 x r0, into scratch[c*32 + rr])
//     X(r0+rr, c)   = -sum_{q <= rr} X_bibi(rr, q) T(q, c)               (stored transposed: S[c*LL_LD + r0+rr])
// Fragments come straight from the tile's storage (X transposed in the strict upper triangle, diagonal in xd), so the
// triangular structure is a select per operand.  The scalar version of this phase took 31 us per tile and sat on the
// critical path of the whole factorisation (every tile (i,j) waits for Dinv_j); as 8 x 8 x 4 DMMAs it is a few us.
// phase 1 of block row bi: T = L(bi, 0..bi-1) X(0..bi-1, 0..bi-1) into scratch (needs X of the rows above only)
template <int NWARPS>
__device__ __forceinline__ void tile_inverse_T(const double* S, const double* xd, double* scratch, int wid, int lane, int bi)
{
    const int g = lane >> 2, t = lane & 3;
    const int r0 = bi * DB;
    const int ntile = (DB / 8) * (r0 / 8);                       // 4 x (4 bi) output tiles of 8 x 8
    for (int tl = wid; tl < ntile; tl += NWARPS) {
        const int mi = tl & 3, ni = tl >> 2;
        const int c = ni * 8 + g;                                // this lane's B column
        const double* La = S + (r0 + mi * 8 + g) * LL_LD;        // this lane's A row: L(r0 + rr, .)
        const double* Xc = S + c * LL_LD;                        // Xc[k] = X(k, c) for k > c
        const double xcc = xd[c];
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;           // two chains: the k-range is a multiple of 8
        for (int kk = ni * 8; kk < r0; kk += 8) {
            const int k = kk + t, k2 = k + 4;
            double b = (k > c) ? Xc[k] : 0.0;
            b = (k == c) ? xcc : b;
            double b2 = (k2 > c) ? Xc[k2] : 0.0;
            b2 = (k2 == c) ? xcc : b2;
            dmma884(c0, c1, La[k], b);
            dmma884(d0, d1, La[k2], b2);
        }
        scratch[(ni * 8 + 2 * t) * DB + mi * 8 + g] = c0 + d0;
        scratch[(ni * 8 + 2 * t + 1) * DB + mi * 8 + g] = c1 + d1;
    }
}

// phase 2 of block row bi: X(bi, 0..bi-1) = -X_bibi T (needs the inverse of diagonal block bi)
template <int NWARPS>
__device__ __forceinline__ void tile_inverse_X(double* S, const double* xd, const double* scratch, int wid, int lane, int bi)
{
    const int g = lane >> 2, t = lane & 3;
    const int r0 = bi * DB;
    const int ntile = (DB / 8) * (r0 / 8);
    for (int tl = wid; tl < ntile; tl += NWARPS) {
        const int mi = tl & 3, ni = tl >> 2;
        const int rr = mi * 8 + g;                               // this lane's A row: X_bibi(rr, .)
        const double* Tc = scratch + (ni * 8 + g) * DB;          // this lane's B column: T(., c)
        const double xrr = xd[r0 + rr];
        double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
        for (int kk = 0; kk < mi * 8 + 8; kk += 8) {
            const int q = kk + t, q2 = q + 4;
            double a = (q < rr) ? S[(r0 + q) * LL_LD + r0 + rr] : 0.0;
            a = (q == rr) ? xrr : a;
            double a2 = (q2 < rr) ? S[(r0 + q2) * LL_LD + r0 + rr] : 0.0;
            a2 = (q2 == rr) ? xrr : a2;
            dmma884(c0, c1, a, Tc[q]);
            dmma884(d0, d1, a2, Tc[q2]);
        }
        S[(ni * 8 + 2 * t) * LL_LD + r0 + rr] = -(c0 + d0);
        S[(ni * 8 + 2 * t + 1) * LL_LD + r0 + rr] = -(c1 + d1);
    }
}

// Panel of block column kb for one row of the tile (one thread per row, nrows a multiple of 32, so whole warps take part):
// l_c = (a_c - sum_{k<c} l_k L_kk(c,k)) / L_kk(c,c), c = 0..31.  The row lives in 32 registers and both loops are fully
// unrolled (496 FMAs on four partial sums per column, the term with the newest l_{c-1} added last so that the dependent
// chain per column is one FMA + the combine): the rolled shared-memory version spent ~5 us per call waiting on its own
// stores.  __noinline__: own register allocation.
__device__ __noinline__ void panel_row32(double* S, int k0, int R0, const double* xd, int tid)
{
    double* Ai = S + (R0 + tid) * LL_LD + k0;
    double row[DB];
#pragma unroll
    for (int c = 0; c < DB; c++) row[c] = Ai[c];
#pragma unroll
    for (int c = 0; c < DB; c++) {
        const double* Lc = S + (k0 + c) * LL_LD + k0;
        double s[4] = {row[c], 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < c; k++) {
            const int slot = (k == c - 1) ? 0 : 1 + (k % 3);     // the last term joins the chain that ends the column
            s[slot] = fma(-row[k], Lc[k], s[slot]);
        }
        row[c] = ((s[1] + s[2]) + s[3] + s[0]) * xd[k0 + c];
    }
#pragma unroll
    for (int c = 0; c < DB; c++) Ai[c] = row[c];
}

// all 256 consumer threads; S lower = matrix on entry, L on exit; upper/xd = X^T; scratch: >= 32*96 doubles; logs: 128 doubles
//   per block column kb (two phases, one block-wide barrier each):
//     phase 1   warp 0: Cholesky of the 32 x 32 diagonal block kb (register-resident, shuffles)
//               warp 1: inverse of diagonal block kb-1 -- one step behind: the block inverses are needed only by the tile
//                       inverse at the end, so they never sit on the chain of block factorisations
//     phase 2   all:    panel L(i, kb) = A(i, kb) L_kk^-T by forward substitution (one thread per row, in place), barrier,
//                       trailing update A(i, c) -= L(i, kb) L(c, kb)^T as 8 x 8 x 4 DMMAs on the lower 8 x 8 tiles
//   then the inverse of the tile: warp 0 inverts the last diagonal block while warps 1-7 form block rows 1 and 2 of the
//   off-diagonal part; block row 3 by all eight warps.
__device__ __noinline__ void diag_tile_factor_invert(double* S, double* xd, double* scratch, double* logs, int tid, int* info, int col0,
                                                     double* logout, unsigned long long* stamp)
{
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    for (int kb = 0; kb < TILE / DB; kb++) {
        const int k0 = kb * DB, R0 = k0 + DB, nrows = TILE - R0;
        unsigned long long tw0 = 0ull;
        if (stamp && tid == 0) tw0 = globaltimer();
        __syncwarp();                                        // (trace only diverges lane 0: re-converge, or the shuffles take their slow path)
        if (warp == 0) warp_factor32(S, k0, xd, info, col0, scratch);
        else if (warp == 1 && kb > 0) warp_invert32(S, k0 - DB, xd);
        if (stamp && tid == 0) stamp[3] += globaltimer() - tw0;      // trace: time in the warp-level block factorisations
        consumer_bar();
        if (nrows > 0) {
            if (tid < nrows) panel_row32(S, k0, R0, xd, tid);
            consumer_bar();                                  // the panel is complete
            // trailing update A(i, c) -= sum_k L(i, k0+k) L(c, k0+k), R0 <= c <= i: lower 8 x 8 tiles, 8 DMMAs each
            const int nb = nrows / 8, ntl = nb * (nb + 1) / 2;
            for (int tl = warp; tl < ntl; tl += WS_CONSUMER_WARPS) {
                int mi = (int)((sqrtf(8.0f * (float)tl + 1.0f) - 1.0f) * 0.5f);
                while ((mi + 1) * (mi + 2) / 2 <= tl) mi++;
                while (mi * (mi + 1) / 2 > tl) mi--;
                const int ni = tl - mi * (mi + 1) / 2;
                const double* La = S + (R0 + mi * 8 + g) * LL_LD + k0 + t;     // A(row g, k t)
                const double* Lb = S + (R0 + ni * 8 + g) * LL_LD + k0 + t;     // B(k t, col g) = L(R0 + 8 ni + g, k)
                double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < DB; kk += 8) {
                    dmma884(c0, c1, La[kk], Lb[kk]);
                    dmma884(d0, d1, La[kk + 4], Lb[kk + 4]);
                }
                const int i = R0 + mi * 8 + g, c = R0 + ni * 8 + 2 * t;
                // (entries above the diagonal belong to the X^T storage of the diagonal blocks: only c <= i is written)
                if (c <= i) S[i * LL_LD + c] -= c0 + d0;
                if (c + 1 <= i) S[i * LL_LD + c + 1] -= c1 + d1;
            }
            consumer_bar();
        }
    }
    // sum(log diag L) = -sum(log xd): one log per thread, fixed-order reduction by warp 0
    if (tid < TILE) logs[tid] = -log(xd[tid]);
    consumer_bar();
    if (stamp && tid == 0) *stamp = globaltimer();           // factor done (trace only)
    __syncwarp();
    if (warp == 0) {
        double sl = (logs[lane] + logs[lane + 32]) + (logs[lane + 64] + logs[lane + 96]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sl += __shfl_xor_sync(0xffffffffu, sl, o);
        if (lane == 0) *logout = sl;
        warp_invert32(S, TILE - DB, xd);
    } else {
        // warps 1-7, meanwhile: block rows 1 and 2 of the off-diagonal inverse and the first phase of block row 3
        // (everything that does not need the inverse of the last diagonal block)
        constexpr int NG = WS_CONSUMER_WARPS - 1;
        tile_inverse_T<NG>(S, xd, scratch, warp - 1, lane, 1);
        group_bar<2, NG * 32>();
        tile_inverse_X<NG>(S, xd, scratch, warp - 1, lane, 1);
        group_bar<2, NG * 32>();
        tile_inverse_T<NG>(S, xd, scratch, warp - 1, lane, 2);
        group_bar<2, NG * 32>();
        tile_inverse_X<NG>(S, xd, scratch, warp - 1, lane, 2);
        group_bar<2, NG * 32>();
        tile_inverse_T<NG>(S, xd, scratch, warp - 1, lane, 3);
    }
    consumer_bar();
    tile_inverse_X<WS_CONSUMER_WARPS>(S, xd, scratch, warp, lane, 3);
    consumer_bar();
}

// columns [wn + 16 HALF, wn + 16 HALF + 16) of a warp's 64 x 32 accumulator block to the tile in global memory
template <int HALF>
__device__ __forceinline__ void store_cols16(const double (&acc)[8][4][2], double* tile, long lda, int wm, int wn, int g, int t)
{
#pragma unroll
    for (int ii = 0; ii < 8; ii++) {
#pragma unroll
        for (int jj = 2 * HALF; jj < 2 * HALF + 2; jj++) {
            double* c0 = tile + (wm + ii * 8 + g) + (long)(wn + jj * 8 + 2 * t) * lda;
            c0[0] = acc[ii][jj][0];
            c0[lda] = acc[ii][jj][1];
        }
    }
}

// TRACE: record per-task time stamps (tools/ll_trace.py); a separate instantiation so that the production
// kernel carries no trace state through the register-tight main loop
// (the helpers of the diagonal tile are __noinline__ so that their register allocation stays out of the DMMA main loop:
// 0 bytes spilled)
template <bool TRACE>
__global__ void __launch_bounds__(WS_THREADS, 1) potrf_ll_kernel(LLArgs a)
{
    extern __shared__ __align__(16) double smem[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)WS_STAGES * 2 * STAGE_DOUBLES);
    unsigned long long* empty = full + WS_STAGES;
    unsigned long long* c2p = empty + WS_STAGES;           // consumers -> producer, once per task
    int* s_stop = reinterpret_cast<int*>(c2p + 1);         // two slots (alternating per task)
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int nt = a.nt;
    const long ntasks = (long)nt * (nt + 1) / 2;
    const volatile int* vabort = a.abort;

    if (tid == 0) {
        for (int s = 0; s < WS_STAGES; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, WS_CONSUMER_WARPS);
        }
        mbar_init(c2p, 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == WS_CONSUMER_WARPS) {
        // ------------------------------------------------------------------ producer warp
        uint32_t it = 0, seq = 0;
        for (long task = blockIdx.x; task < ntasks; task += gridDim.x, seq++) {
            if (*vabort) return;
            int i, j;
            ll_ticket(task, nt, i, j);
            const double* rowi = a.A + (long)i * TILE;
            const double* rowj = a.A + (long)j * TILE;
            for (int kb = 0; kb < j; kb++) {
                if (i == j && kb == j - 1) {
                    // the tile L(j, j-1) comes off the critical path of the whole factorisation (its task had to wait for
                    // Dinv_{j-1}): take it slab by slab as that task stores its column slabs, instead of waiting for the tile
                    for (int sl = 0; sl < TILE / GK; sl++) {
                        if (lane == 0) wait_count(a.slab + (long)kb * (TILE / GK) + sl, 2, a.abort, a.info);
                        __syncwarp();
                        fence_proxy_async();
                        const double* src = rowi + ((long)kb * TILE + (long)sl * GK) * a.lda;
                        ll_produce(src, a.lda, src, a.lda, 1, it, smem, full, empty, lane, vabort);
                    }
                    continue;
                }
                if (lane == 0) {
                    wait_ready(a.ready + i + (long)kb * nt, a.abort, a.info);
                    if (i != j) wait_ready(a.ready + j + (long)kb * nt, a.abort, a.info);
                }
                __syncwarp();
                fence_proxy_async();
                ll_produce(rowi + (long)kb * TILE * a.lda, a.lda, rowj + (long)kb * TILE * a.lda, a.lda, TILE / GK, it, smem, full,
                           empty, lane, vabort);
            }
            // consumers: C' stored (i > j) or diagonal tile finished with the ring memory (i == j)
            mbar_wait_ab(c2p, seq & 1u, vabort);
            if (i != j) {
                if (lane == 0) wait_ready(a.ready + j + (long)j * nt, a.abort, a.info);
                __syncwarp();
                fence_proxy_async();
                ll_produce(a.A + (long)i * TILE + (long)j * TILE * a.lda, a.lda, a.Dinv + (long)j * TILE * TILE, TILE, TILE / GK, it,
                           smem, full, empty, lane, vabort);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64;
    const int wn = (warp >> 1) * 32;
    uint32_t it = 0, seq = 0;
    for (long task = blockIdx.x; task < ntasks; task += gridDim.x, seq++) {
        // uniform stop decision for the 8 consumer warps (they share named barriers below)
        if (tid == 0) s_stop[seq & 1u] = *vabort;
        consumer_bar();
        if (s_stop[seq & 1u]) return;

        int i, j;
        ll_ticket(task, nt, i, j);
        if (i == j) {
            // the diagonal tile A(j,j) is read once, after the main loop, on the critical path of the whole factorisation:
            // pull it into L2 now (128 columns x 1 KB = 1024 lines, 4 per thread)
            const char* tp = reinterpret_cast<const char*>(a.A + (long)j * TILE + (long)j * TILE * a.lda);
            for (int ln = tid; ln < TILE * 8; ln += LL_CONSUMERS)
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(tp + (long)(ln >> 3) * a.lda * 8 + (ln & 7) * 128));
        }
        unsigned long long* dbg = (TRACE && a.dbg) ? a.dbg + task * 8 : nullptr;
        if (TRACE && dbg && tid == 0) { dbg[0] = globaltimer(); dbg[7] = ((unsigned long long)i << 32) | (unsigned)j; }
        double acc[8][4][2];
#pragma unroll
        for (int ii = 0; ii < 8; ii++)
#pragma unroll
            for (int jj = 0; jj < 4; jj++) acc[ii][jj][0] = acc[ii][jj][1] = 0.0;
        const bool solve = (i == j) && a.y != nullptr;
        double vpart = 0.0;
        if (solve) ll_consume<true>(acc, j * (TILE / GK), it, smem, full, empty, wm, wn, g, t, lane, vabort, a.w, tid, &vpart);
        else ll_consume<false>(acc, j * (TILE / GK), it, smem, full, empty, wm, wn, g, t, lane, vabort);

        if (TRACE && dbg && tid == 0) dbg[1] = globaltimer();           // main loop done
        double* tile = a.A + (long)i * TILE + (long)j * TILE * a.lda;
        if (i == j) {
            // ---- diagonal tile: S = A(j,j) - acc in shared memory (ring is idle: the producer waits on c2p)
            consumer_bar();                                  // every warp is out of the main loop
            double* S = smem;
            double* tmp = smem + TILE * LL_LD;
#pragma unroll
            for (int ii = 0; ii < 8; ii++) {
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    const int row = wm + ii * 8 + g, col = wn + jj * 8 + 2 * t;
                    if (row >= col) S[row * LL_LD + col] = tile[row + (long)col * a.lda] - acc[ii][jj][0];
                    if (row >= col + 1) S[row * LL_LD + col + 1] = tile[row + (long)(col + 1) * a.lda] - acc[ii][jj][1];
                }
            }
            consumer_bar();
            double* xd = tmp + 3 * TILE;                     // diagonal of the inverse
            double* scratch = tmp + 4 * TILE;                // 32 x 96 doubles
            if (TRACE && dbg && tid == 0) dbg[2] = globaltimer();       // S built
            diag_tile_factor_invert(S, xd, scratch, tmp, tid, a.info, j * TILE, a.logparts + j, (TRACE && dbg) ? dbg + 3 : nullptr);
            if (TRACE && dbg && tid == 0) dbg[4] = globaltimer();       // factored and inverted
            if (solve) {
                // w_j = X (y_j - v_j); the two k-halves of v_j are in vpart
                tmp[tid] = vpart;
                consumer_bar();
                if (tid < TILE) tmp[2 * TILE + tid] = a.y[(long)j * TILE + tid] - (tmp[tid] + tmp[tid + TILE]);
                consumer_bar();
                if (tid < TILE) {
                    double sacc = xd[tid] * tmp[2 * TILE + tid];
                    for (int c = 0; c < tid; c++) sacc = fma(S[c * LL_LD + tid], tmp[2 * TILE + c], sacc);
                    a.w[(long)j * TILE + tid] = sacc;
                }
            }
            // Dinv_j (and w_j) first: they are all that the tiles below and the later diagonal tasks wait for.  L(j,j) itself
            // is read by no task of this kernel (only off-diagonal tiles are operands), so its store comes after the flag.
            double* Dj = a.Dinv + (long)j * TILE * TILE;
            for (int idx = tid; idx < TILE * TILE; idx += LL_CONSUMERS) {
                const int rr = idx & (TILE - 1), c = idx >> 7;
                Dj[rr + c * TILE] = (rr > c) ? S[c * LL_LD + rr] : (rr == c ? xd[c] : 0.0);
            }
            __threadfence();
            consumer_bar();                                  // Dinv_j and w_j stored and fenced
            if (tid == 0) {
                st_release(a.ready + j + (long)j * nt, 1);
                if (TRACE && dbg) dbg[5] = globaltimer();
            }
            for (int idx = tid; idx < TILE * TILE; idx += LL_CONSUMERS) {
                const int rr = idx & (TILE - 1), c = idx >> 7;
                tile[rr + (long)c * a.lda] = (rr >= c) ? S[rr * LL_LD + c] : 0.0;
            }
            consumer_bar();                                  // S no longer used
            if (tid == 0) mbar_arrive(c2p);                  // the producer may refill the ring
        } else {
            // ---- off-diagonal tile: C' = A(i,j) - acc, in place; then L(i,j) = C' Dinv_j^T through the ring
#pragma unroll
            for (int ii = 0; ii < 8; ii++) {
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    double* c0 = tile + (wm + ii * 8 + g) + (long)(wn + jj * 8 + 2 * t) * a.lda;
                    double* c1 = c0 + a.lda;
                    *c0 = *c0 - acc[ii][jj][0];
                    *c1 = *c1 - acc[ii][jj][1];
                }
            }
            __threadfence();
            fence_proxy_async();
            consumer_bar();
            if (tid == 0) mbar_arrive(c2p);                  // producer: C' is in global memory
            if (TRACE && dbg && tid == 0) dbg[2] = globaltimer();
#pragma unroll
            for (int ii = 0; ii < 8; ii++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) acc[ii][jj][0] = acc[ii][jj][1] = 0.0;
            if (i == j + 1) {
                // sub-diagonal tile: Dinv_j is lower triangular, so columns [16 s, 16 s + 16) of L(i,j) are final after k-slab s.
                // The two warps that own them store them at once and bump the slab counter: the diagonal task of the next
                // column streams the tile slab by slab behind this pass instead of starting after it.
                int* cnt = a.slab + (long)j * (TILE / GK);
#pragma unroll
                for (int sp = 0; sp < TILE / GK / 2; sp++) {
                    ll_consume<false>(acc, 1, it, smem, full, empty, wm, wn, g, t, lane, vabort);
                    if (wn == 32 * sp) {
                        store_cols16<0>(acc, tile, a.lda, wm, wn, g, t);
                        __threadfence();
                        __syncwarp();
                        if (lane == 0) atomicAdd(cnt + 2 * sp, 1);
                    }
                    ll_consume<false>(acc, 1, it, smem, full, empty, wm, wn, g, t, lane, vabort);
                    if (wn == 32 * sp) {
                        store_cols16<1>(acc, tile, a.lda, wm, wn, g, t);
                        __threadfence();
                        __syncwarp();
                        if (lane == 0) atomicAdd(cnt + 2 * sp + 1, 1);
                    }
                }
            } else {
                ll_consume<false>(acc, TILE / GK, it, smem, full, empty, wm, wn, g, t, lane, vabort);
#pragma unroll
                for (int ii = 0; ii < 8; ii++) {
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) {
                        double* c0 = tile + (wm + ii * 8 + g) + (long)(wn + jj * 8 + 2 * t) * a.lda;
                        double* c1 = c0 + a.lda;
                        *c0 = acc[ii][jj][0];
                        *c1 = acc[ii][jj][1];
                    }
                }
            }
            __threadfence();
            consumer_bar();
            if (tid == 0) st_release(a.ready + i + (long)j * nt, 1);
            if (TRACE && dbg && tid == 0) dbg[5] = globaltimer();
        }
    }
}

constexpr size_t LL_SMEM = (size_t)WS_STAGES * 2 * STAGE_DOUBLES * sizeof(double) + (2 * WS_STAGES + 1) * sizeof(unsigned long long) + 16;

}  // namespace

size_t potrf_ll_flag_bytes(long n_pad)
{
    const size_t nt = (size_t)(n_pad / TILE);
    return (nt * nt + 4 + nt * (TILE / GK)) * sizeof(int);
}

// flags: potrf_ll_flag_bytes(n_pad) bytes of device scratch
int potrf_ll(Ctx& c, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, int* flags, const double* y,
             double* w)
{
    static bool configured[64] = {};
    if (first_use_on_current_device(configured)) {
        SGP_CUDA(cudaFuncSetAttribute(potrf_ll_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LL_SMEM));
        SGP_CUDA(cudaFuncSetAttribute(potrf_ll_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LL_SMEM));
    }
    static_assert((size_t)TILE * LL_LD * sizeof(double) + (4 * TILE + 32 * 96) * sizeof(double) <= (size_t)WS_STAGES * 2 * STAGE_DOUBLES * sizeof(double),
                  "diagonal tile scratch must fit in the operand ring");
    const int nt = (int)(n_pad / TILE);
    const size_t nflags = (size_t)nt * nt;
    SGP_CUDA(cudaMemsetAsync(flags, 0, potrf_ll_flag_bytes(n_pad), c.stream));
    LLArgs a;
    a.A = A; a.lda = lda; a.nt = nt; a.Dinv = Dinv; a.logparts = logparts; a.info = info;
    a.ready = flags; a.abort = flags + nflags; a.slab = flags + nflags + 4;
    a.y = (y && w) ? y : nullptr; a.w = w;
    a.dbg = nullptr;
    // SGP_LL_TRACE=1: per-task time stamps (tools/ll_trace.py), dumped to $SGP_LL_TRACE_FILE after the launch
    static const int trace = [] { const char* v = getenv("SGP_LL_TRACE"); return (v && v[0] == '1') ? 1 : 0; }();
    static DBuf tracebuf;
    const long ntasks_all = (long)nt * (nt + 1) / 2;
    if (trace) {
        SGP_TRY(tracebuf.reserve((size_t)ntasks_all * 8 * sizeof(unsigned long long)));
        SGP_CUDA(cudaMemsetAsync(tracebuf.p, 0, (size_t)ntasks_all * 8 * sizeof(unsigned long long), c.stream));
        a.dbg = tracebuf.as<unsigned long long>();
    }
    const long ntasks = (long)nt * (nt + 1) / 2;
    const int sms = c.sm_count > 0 ? c.sm_count : 148;
    const unsigned grid = (unsigned)(ntasks < sms ? ntasks : sms);
    void* args[] = {&a};
    const void* kern = trace ? (const void*)potrf_ll_kernel<true> : (const void*)potrf_ll_kernel<false>;
    SGP_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(WS_THREADS), args, LL_SMEM, c.stream));
    count_launch();
    if (trace) {
        std::vector<unsigned long long> h((size_t)ntasks_all * 8);
        SGP_CUDA(cudaStreamSynchronize(c.stream));
        SGP_CUDA(cudaMemcpy(h.data(), tracebuf.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        const char* fn = getenv("SGP_LL_TRACE_FILE");
        FILE* f = fopen(fn ? fn : "potrf_ll_trace.txt", "w");
        if (f) {
            unsigned long long t0 = ~0ull;
            for (long t = 0; t < ntasks_all; t++) if (h[t * 8] && h[t * 8] < t0) t0 = h[t * 8];
            for (long t = 0; t < ntasks_all; t++) {
                fprintf(f, "%ld %d %d", t, (int)(h[t * 8 + 7] >> 32), (int)(h[t * 8 + 7] & 0xffffffffu));
                for (int k = 0; k < 6; k++) fprintf(f, " %lld", h[t * 8 + k] ? (long long)(h[t * 8 + k] - t0) : -1ll);
                fprintf(f, " %lld\n", (long long)h[t * 8 + 6]);       // accumulated ns in the 32 x 32 block factorisations
            }
            fclose(f);
        }
    }
    return ST_OK;
}

}  // namespace sgp
