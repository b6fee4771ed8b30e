// Blocked FP64 Cholesky (potrf), triangular solves (potrs), triangular inverse (trtri) and
// X^T X (lauum) on one B200, built from the structured DMMA GEMM of dmma_gemm.cuh.
//
// Replaces the LAPACK calls the reference makes from Python:
//   scipy.linalg.cholesky      python/05_tokamak/SympGPR/func.py:147       -> potrf
//   solve_triangular x2        python/02_pert_pendulum/func.py:173-177     -> potrs
//   np.linalg.inv (LU)         python/02_pert_pendulum/func.py:152         -> potri = trtri + lauum
//
// All matrices are column-major with order n_pad (a multiple of 128; rows/cols >= n carry an
// identity block so factor, log-determinant and inverse of the leading n x n block are
// unchanged).  Only the lower triangle is referenced.
#include "chol.cuh"

#include <limits.h>
#include <stdlib.h>

namespace sgp {

// ------------------------------------------------------------------------------------------
// 128 x 128 diagonal tile: factor in place, zero the strict upper part, emit inv(L) and
// sum(log(diag L)).  One CTA, tile resident in shared memory (row stride 129 doubles).
// ------------------------------------------------------------------------------------------
constexpr int PT_LD = TILE + 1;
constexpr int PT_THREADS = 512;                      // 4 threads per row
constexpr size_t PT_SMEM = (size_t)(TILE * PT_LD + 2 * PT_THREADS) * sizeof(double);

__global__ void __launch_bounds__(PT_THREADS, 1)
potrf_tile_kernel(double* __restrict__ A, long lda, double* __restrict__ Dinv, double* __restrict__ logpart,
                  int* __restrict__ info, int col0)
{
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                      // S[r*PT_LD + c]
    double* tmp = sm + TILE * PT_LD;     // 2*PT_THREADS doubles scratch
    const int tid = threadIdx.x;

    for (int idx = tid; idx < TILE * TILE; idx += PT_THREADS) {
        int r = idx & (TILE - 1), c = idx >> 7;
        S[r * PT_LD + c] = (r >= c) ? A[r + (long)c * lda] : 0.0;
    }
    __syncthreads();

    const int r = tid & (TILE - 1);      // row owned in the trailing update
    const int part = tid >> 7;           // columns c with (c - j - 1) % 4 == part
    double mylog = 0.0;
    double* Sr = S + r * PT_LD;

    for (int j = 0; j < TILE; j++) {
        double d = S[j * PT_LD + j];
        bool bad = !(d > 0.0);           // also true for NaN
        if (bad) {
            if (tid == 0) atomicCAS(info, 0, col0 + j + 1);
            d = 1.0;
        }
        const double dj = sqrt(d);
        const double inv = 1.0 / dj;
        __syncthreads();                 // everyone has read S[j][j]
        if (part == 0) {
            if (r > j) Sr[j] *= inv;
            else if (r == j) { Sr[j] = dj; mylog = log(dj); }
        }
        __syncthreads();
        if (r > j) {
            const double lrj = Sr[j];
            // columns c = j+1+part, j+5+part, ... <= r ; four independent updates per trip
            int c = j + 1 + part;
            for (; c + 12 <= r; c += 16) {
                const double l0 = S[c * PT_LD + j], l1 = S[(c + 4) * PT_LD + j], l2 = S[(c + 8) * PT_LD + j],
                             l3 = S[(c + 12) * PT_LD + j];
                const double s0 = Sr[c], s1 = Sr[c + 4], s2 = Sr[c + 8], s3 = Sr[c + 12];
                Sr[c] = s0 - lrj * l0; Sr[c + 4] = s1 - lrj * l1; Sr[c + 8] = s2 - lrj * l2; Sr[c + 12] = s3 - lrj * l3;
            }
            for (; c <= r; c += 4) Sr[c] -= lrj * S[c * PT_LD + j];
        }
        __syncthreads();                 // S[j+1][j+1] final before the next pivot is read
    }

    // L back to global (lower + explicit zeros above the diagonal)
    for (int idx = tid; idx < TILE * TILE; idx += PT_THREADS) {
        int rr = idx & (TILE - 1), c = idx >> 7;
        A[rr + (long)c * lda] = S[rr * PT_LD + c];
    }
    // sum of log(diag): fixed order
    if (part == 0) tmp[r] = mylog;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int k = 0; k < TILE; k++) s += tmp[k];
        *logpart = s;
    }
    __syncthreads();

    // in-place inverse of the lower-triangular tile (unblocked, last column first):
    //   X[j][j] = 1/L[j][j];  X[j+1:, j] = -X[j+1:, j+1:] * L[j+1:, j] * X[j][j]
    for (int j = TILE - 1; j >= 0; j--) {
        const double xjj = 1.0 / S[j * PT_LD + j];
        double p0 = 0.0, p1 = 0.0;
        if (r > j) {
            int k = j + 1 + part;
            for (; k + 4 <= r; k += 8) {
                p0 += Sr[k] * S[k * PT_LD + j];
                p1 += Sr[k + 4] * S[(k + 4) * PT_LD + j];
            }
            for (; k <= r; k += 4) p0 += Sr[k] * S[k * PT_LD + j];
        }
        tmp[tid] = p0 + p1;
        __syncthreads();                 // all reads of column j (still L) done
        if (part == 0) {
            if (r > j) Sr[j] = -((tmp[r] + tmp[r + TILE]) + (tmp[r + 2 * TILE] + tmp[r + 3 * TILE])) * xjj;
            else if (r == j) Sr[j] = xjj;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < TILE * TILE; idx += PT_THREADS) {
        int rr = idx & (TILE - 1), c = idx >> 7;
        Dinv[rr + c * TILE] = S[rr * PT_LD + c];
    }
}

// copy a 128x128 tile (src ld 128) into the matrix
__global__ void copy_tile_kernel(double* __restrict__ dst, long ldd, const double* __restrict__ src)
{
    for (int idx = threadIdx.x; idx < TILE * TILE; idx += blockDim.x) {
        int r = idx & (TILE - 1), c = idx >> 7;
        dst[r + (long)c * ldd] = src[idx];
    }
}

// ------------------------------------------------------------------------------------------
// potrs by blocked substitution with the tile inverses.
// forward step j:  w_j = Dinv_j * y_j ;  y[rows below] -= L[rows below, j] * w_j
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE)
trsv_fwd_step_kernel(const double* __restrict__ L, long ld, const double* __restrict__ Dinv, double* __restrict__ y,
                     double* __restrict__ w, int j)
{
    __shared__ double yj[TILE];
    __shared__ double wj[TILE];
    const int tid = threadIdx.x;
    const long base = (long)j * TILE;
    yj[tid] = y[base + tid];
    __syncthreads();
    {
        const double* D = Dinv + (long)j * TILE * TILE;
        double s = 0.0;
        for (int c = 0; c <= tid; c++) s += D[tid + c * TILE] * yj[c];
        wj[tid] = s;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        w[base + tid] = wj[tid];
        return;
    }
    const long row = base + (long)blockIdx.x * TILE + tid;     // blockIdx.x >= 1: rows below the tile
    const double* Lr = L + row + base * ld;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
    for (int c = 0; c < TILE; c += 2) {
        s0 += Lr[(long)c * ld] * wj[c];
        s1 += Lr[(long)(c + 1) * ld] * wj[c + 1];
    }
    y[row] -= (s0 + s1);
}

// backward step j:  a_j = Dinv_j^T * w_j ;  w[cols left] -= L[tile rows j, cols left]^T * a_j
__global__ void __launch_bounds__(256)
trsv_bwd_step_kernel(const double* __restrict__ L, long ld, const double* __restrict__ Dinv, double* __restrict__ w,
                     double* __restrict__ a, int j)
{
    __shared__ double wj[TILE];
    __shared__ double aj[TILE];
    const int tid = threadIdx.x;
    const long base = (long)j * TILE;
    if (tid < TILE) wj[tid] = w[base + tid];
    __syncthreads();
    if (tid < TILE) {
        const double* D = Dinv + (long)j * TILE * TILE;
        double s = 0.0;
        for (int rr = tid; rr < TILE; rr++) s += D[rr + tid * TILE] * wj[rr];   // (Dinv^T)[tid][rr]
        aj[tid] = s;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        if (tid < TILE) a[base + tid] = aj[tid];
        return;
    }
    // blocks 1..: 64 columns each, 8 warps x 8 columns
    const int warp = tid >> 5, lane = tid & 31;
    const long c0 = (long)(blockIdx.x - 1) * 64 + warp * 8;
    for (int cc = 0; cc < 8; cc++) {
        const long col = c0 + cc;
        const double* Lc = L + base + col * ld;
        double s = Lc[lane] * aj[lane] + Lc[lane + 32] * aj[lane + 32] + Lc[lane + 64] * aj[lane + 64] +
                   Lc[lane + 96] * aj[lane + 96];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) w[col] -= s;
    }
}

// ------------------------------------------------------------------------------------------
// drivers
// ------------------------------------------------------------------------------------------
static int g_potrf_cfg = 0;

static int launch_potrf_tile(Ctx& c, double* A, long lda, double* Dinv, double* logparts, int* info, int jt)
{
    if (!g_potrf_cfg) {
        SGP_CUDA(cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
        g_potrf_cfg = 1;
    }
    const long o = (long)jt * TILE;
    potrf_tile_kernel<<<1, PT_THREADS, PT_SMEM, c.stream>>>(A + o + o * lda, lda, Dinv + (long)jt * TILE * TILE, logparts + jt,
                                                      info, (int)o);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

static int gemm(Ctx& c, int al, int bl, const double* A, long lda, const double* B, long ldb, double* C, long ldc,
                int Mt, int Nt, long K, double alpha, double beta, int mode)
{
    GemmArgs g;
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    g.Mt = Mt; g.Nt = Nt; g.K = (int)K; g.alpha = alpha; g.beta = beta; g.mode = mode;
    return dmma_gemm(c, al, bl, g);
}

#define AT(A, lda, rt, ct) ((A) + (long)(rt) * TILE + (long)(ct) * TILE * (lda))

// X * L[l0..l0+kt)^T = B,  B = A[r0..r0+rt) x [l0..l0+kt)  (tile indices), in place
static int trsm_rec(Ctx& c, double* A, long lda, const double* Dinv, int l0, int kt, int r0, int rt)
{
    if (rt <= 0) return ST_OK;
    if (kt == 1) {
        double* B = AT(A, lda, r0, l0);
        return gemm(c, LAYOUT_MN, LAYOUT_MN, B, lda, Dinv + (long)l0 * TILE * TILE, TILE, B, lda, rt, 1, TILE, 1.0, 0.0,
                    TM_FULL);
    }
    const int k1 = kt / 2, k2 = kt - k1;
    SGP_TRY(trsm_rec(c, A, lda, Dinv, l0, k1, r0, rt));
    SGP_TRY(gemm(c, LAYOUT_MN, LAYOUT_MN, AT(A, lda, r0, l0), lda, AT(A, lda, l0 + k1, l0), lda, AT(A, lda, r0, l0 + k1),
                 lda, rt, k2, (long)k1 * TILE, -1.0, 1.0, TM_FULL));
    return trsm_rec(c, A, lda, Dinv, l0 + k1, k2, r0, rt);
}

static int potrf_rec(Ctx& c, double* A, long lda, double* Dinv, double* logparts, int* info, int j0, int mt)
{
    if (mt == 1) return launch_potrf_tile(c, A, lda, Dinv, logparts, info, j0);
    const int m1 = mt / 2, m2 = mt - m1;
    SGP_TRY(potrf_rec(c, A, lda, Dinv, logparts, info, j0, m1));
    SGP_TRY(trsm_rec(c, A, lda, Dinv, j0, m1, j0 + m1, m2));
    double* B = AT(A, lda, j0 + m1, j0);
    SGP_TRY(gemm(c, LAYOUT_MN, LAYOUT_MN, B, lda, B, lda, AT(A, lda, j0 + m1, j0 + m1), lda, m2, m2, (long)m1 * TILE,
                 -1.0, 1.0, TM_LOWER));
    return potrf_rec(c, A, lda, Dinv, logparts, info, j0 + m1, m2);
}

static int env_int(const char* name, int dflt, int lo, int hi)
{
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    int x = atoi(v);
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}

// Right-looking blocked factorisation: panels of NB = nb_tiles*128 columns.  The diagonal block is
// factored recursively (small, latency-bound launches), the panel below it is one recursive TRSM over
// all remaining rows and the trailing update one lower-triangular SYRK launch with K = NB, so almost
// all flops run in launches that fill the 148 SMs.
static int potrf_multi(Ctx& c, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info);

int potrf(Ctx& c, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, double* y, double* w)
{
    if (n_pad % TILE || lda % 2) { set_error("potrf: n_pad %ld / lda %ld not tile aligned", n_pad, lda); return ST_BADARG; }
    // default: the left-looking persistent kernel (potrf_ll.cu) with the forward substitution fused in.
    // SGP_POTRF=rec selects the recursive / blocked multi-launch drivers below (kept for comparison).
    static const int use_rec = [] { const char* v = getenv("SGP_POTRF"); return (v && v[0] == 'r') ? 1 : 0; }();
    if (!use_rec) {
        SGP_TRY(c.flags.reserve(potrf_ll_flag_bytes(n_pad)));
        return potrf_ll(c, A, n_pad, lda, Dinv, logparts, info, c.flags.as<int>(), y, w);
    }
    SGP_TRY(potrf_multi(c, A, n_pad, lda, Dinv, logparts, info));
    if (y && w) SGP_TRY(trsv_fwd(c, A, n_pad, lda, Dinv, y, w));
    return ST_OK;
}

static int potrf_multi(Ctx& c, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info)
{
    const int nt = (int)(n_pad / TILE);
    const int nb = env_int("SGP_POTRF_NB", 0, 0, 64);
    if (nb == 0) return potrf_rec(c, A, lda, Dinv, logparts, info, 0, nt);      // fully recursive
    const int lookahead = env_int("SGP_LOOKAHEAD", 1, 0, 1) && c.side != nullptr && nt > 2 * nb;
    cudaStream_t s0 = c.stream, s1 = c.side;

    auto panel = [&](int k0, int kb) -> int {            // diagonal block + everything below it
        SGP_TRY(potrf_rec(c, A, lda, Dinv, logparts, info, k0, kb));
        return trsm_rec(c, A, lda, Dinv, k0, kb, k0 + kb, nt - (k0 + kb));
    };

    if (!lookahead) {
        for (int k0 = 0; k0 < nt; k0 += nb) {
            const int kb = (nt - k0 < nb) ? (nt - k0) : nb;
            SGP_TRY(panel(k0, kb));
            const int rt = nt - (k0 + kb);
            if (rt > 0) {
                double* B = AT(A, lda, k0 + kb, k0);
                SGP_TRY(gemm(c, LAYOUT_MN, LAYOUT_MN, B, lda, B, lda, AT(A, lda, k0 + kb, k0 + kb), lda, rt, rt,
                             (long)kb * TILE, -1.0, 1.0, TM_LOWER));
            }
        }
        return ST_OK;
    }

    // Look-ahead of depth one: after panel k, first update only the next block column, then factor
    // panel k+1 on the high-priority side stream while the main stream applies panel k to the rest
    // of the trailing matrix.  The two touch disjoint block columns.
    SGP_TRY(panel(0, nb < nt ? nb : nt));
    for (int k0 = 0; k0 < nt; k0 += nb) {
        const int kb = (nt - k0 < nb) ? (nt - k0) : nb;
        const int n0 = k0 + kb;                          // first tile of the next panel
        const int rt = nt - n0;
        if (rt <= 0) break;
        const int kn = (rt < nb) ? rt : nb;              // width of the next panel
        double* B = AT(A, lda, n0, k0);
        // (a) next block column: rows n0.., cols n0..n0+kn
        SGP_TRY(gemm(c, LAYOUT_MN, LAYOUT_MN, B, lda, B, lda, AT(A, lda, n0, n0), lda, rt, kn, (long)kb * TILE, -1.0, 1.0,
                     TM_FULL));
        SGP_CUDA(cudaEventRecord(c.ev[0], s0));
        SGP_CUDA(cudaStreamWaitEvent(s1, c.ev[0], 0));
        c.stream = s1;
        int st = panel(n0, kn);
        c.stream = s0;
        SGP_TRY(st);
        SGP_CUDA(cudaEventRecord(c.ev[1], s1));
        // (b) rest of the trailing matrix: rows/cols from n0+kn
        const int r2 = rt - kn;
        if (r2 > 0) {
            double* B2 = AT(A, lda, n0 + kn, k0);
            SGP_TRY(gemm(c, LAYOUT_MN, LAYOUT_MN, B2, lda, B2, lda, AT(A, lda, n0 + kn, n0 + kn), lda, r2, r2, (long)kb * TILE,
                         -1.0, 1.0, TM_LOWER));
        }
        SGP_CUDA(cudaStreamWaitEvent(s0, c.ev[1], 0));
    }
    return ST_OK;
}

int trsv_fwd(Ctx& c, const double* L, long n_pad, long lda, const double* Dinv, double* y, double* w)
{
    const int nt = (int)(n_pad / TILE);
    for (int j = 0; j < nt; j++) {
        trsv_fwd_step_kernel<<<nt - j, TILE, 0, c.stream>>>(L, lda, Dinv, y, w, j);
    }
    SGP_CUDA(cudaGetLastError());
    count_launch((unsigned long long)nt);
    return ST_OK;
}

int trsv_bwd(Ctx& c, const double* L, long n_pad, long lda, const double* Dinv, double* w, double* alpha)
{
    const int nt = (int)(n_pad / TILE);
    for (int j = nt - 1; j >= 0; j--) {
        trsv_bwd_step_kernel<<<1 + 2 * j, 256, 0, c.stream>>>(L, lda, Dinv, w, alpha, j);
    }
    SGP_CUDA(cudaGetLastError());
    count_launch((unsigned long long)nt);
    return ST_OK;
}

// alpha = W y, W symmetric, lower triangle stored (tiles tm >= tn complete, as lauum writes them).  One CTA
// per tile row i: the tiles to the left (j <= i) contribute W(i,j) y_j, the tiles below (k > i) their
// transposes W(k,i)^T y_k.  Every row reads nt tiles, so the CTAs are balanced; each tile is read twice
// overall (8 n^2 bytes); the summation order is fixed.
__global__ void __launch_bounds__(256)
symv_lower_kernel(const double* __restrict__ W, long ld, const double* __restrict__ y, double* __restrict__ alpha, int nt)
{
    __shared__ double ys[TILE];
    __shared__ double part[2][TILE];
    __shared__ double colacc[TILE];
    const int i = blockIdx.x, tid = threadIdx.x;
    const int r = tid & (TILE - 1), half = tid >> 7;
    const int warp = tid >> 5, lane = tid & 31;
    double acc = 0.0;
    // left part: row r of W(i,j), columns [half*64, half*64+64)
    for (int j = 0; j <= i; j++) {
        __syncthreads();
        if (tid < TILE) ys[tid] = y[(long)j * TILE + tid];
        __syncthreads();
        const double* Wt = W + (long)i * TILE + r + ((long)j * TILE + half * 64) * ld;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
        for (int cc = 0; cc < 64; cc += 2) {
            a0 = fma(Wt[(long)cc * ld], ys[half * 64 + cc], a0);
            a1 = fma(Wt[(long)(cc + 1) * ld], ys[half * 64 + cc + 1], a1);
        }
        acc += a0 + a1;
    }
    part[half][r] = acc;
    if (tid < TILE) colacc[tid] = 0.0;
    // lower part: column c of W(k,i), reduced over its 128 rows by one warp (16 columns per warp)
    for (int k = i + 1; k < nt; k++) {
        __syncthreads();
        if (tid < TILE) ys[tid] = y[(long)k * TILE + tid];
        __syncthreads();
        for (int cc = 0; cc < 16; cc++) {
            const int col = warp * 16 + cc;
            const double* Wc = W + (long)k * TILE + ((long)i * TILE + col) * ld;
            double sacc = Wc[lane] * ys[lane] + Wc[lane + 32] * ys[lane + 32] + Wc[lane + 64] * ys[lane + 64] + Wc[lane + 96] * ys[lane + 96];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (lane == 0) colacc[col] += sacc;
        }
    }
    __syncthreads();
    if (tid < TILE) alpha[(long)i * TILE + tid] = (part[0][tid] + part[1][tid]) + colacc[tid];
}

// alpha = X^T w for the explicit lower-triangular X = L^-1 (alpha = L^-T w, the second half of potrs, once the
// inverse factor exists): one warp per column, the column's n - c entries are contiguous, 4 loads in flight per
// lane, fixed summation order.  4 n^2 bytes, HBM-read bound.
__global__ void __launch_bounds__(256)
gemv_t_lower_kernel(const double* __restrict__ X, long ld, const double* __restrict__ w, double* __restrict__ alpha, long n)
{
    const int lane = threadIdx.x & 31;
    const long c = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= n) return;
    const double* col = X + c * ld;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    long r = c + lane;
    for (; r + 96 < n; r += 128) {
        s0 = fma(col[r], w[r], s0);
        s1 = fma(col[r + 32], w[r + 32], s1);
        s2 = fma(col[r + 64], w[r + 64], s2);
        s3 = fma(col[r + 96], w[r + 96], s3);
    }
    for (; r < n; r += 32) s0 = fma(col[r], w[r], s0);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) alpha[c] = s;
}

int gemv_t_lower(Ctx& c, const double* X, long n_pad, long ldx, const double* w, double* alpha)
{
    gemv_t_lower_kernel<<<(unsigned)((n_pad + 7) / 8), 256, 0, c.stream>>>(X, ldx, w, alpha, n_pad);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

int symv_lower(Ctx& c, const double* W, long n_pad, long ldw, const double* y, double* alpha)
{
    const int nt = (int)(n_pad / TILE);
    symv_lower_kernel<<<nt, 256, 0, c.stream>>>(W, ldw, y, alpha, nt);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

// scratch (in 128x128 tiles) of trtri_rec on mt tiles: the m2 x m1 product T of this level, after the two
// halves, which may run concurrently and therefore get disjoint regions
static size_t trtri_tiles(int mt)
{
    if (mt <= 1) return 0;
    const int m1 = mt / 2, m2 = mt - m1;
    const size_t own = (size_t)m1 * m2, kids = trtri_tiles(m1) + trtri_tiles(m2);
    return own > kids ? own : kids;
}

constexpr int TRTRI_FORK_MAX = 32;     // blocks of up to 32 tiles run their two halves on different streams

// side streams / events for the forked sub-trees (created on first use, live as long as the process)
static cudaStream_t g_fork_streams[8];
static cudaEvent_t g_fork_events[64];
static int g_fork_ready = 0, g_fork_next_stream = 0, g_fork_next_event = 0;

static int fork_init()
{
    if (g_fork_ready) return ST_OK;
    for (auto& s : g_fork_streams) SGP_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (auto& e : g_fork_events) SGP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_fork_ready = 1;
    return ST_OK;
}

// X = L^-1 for the diagonal block of mt tiles starting at tile j0, in place, on c.stream.
// X11 and X22 are independent; X21 = -X22 L21 X11.  The small blocks at the bottom of the recursion are
// launches of a few CTAs each: their two halves are forked onto side streams so that they fill the GPU together.
static int trtri_rec(Ctx& c, double* A, long lda, const double* Dinv, double* T, int j0, int mt)
{
    if (mt == 1) {
        copy_tile_kernel<<<1, 256, 0, c.stream>>>(AT(A, lda, j0, j0), lda, Dinv + (long)j0 * TILE * TILE);
        count_launch();
        SGP_CUDA(cudaGetLastError());
        return ST_OK;
    }
    const int m1 = mt / 2, m2 = mt - m1;
    double* T2 = T + trtri_tiles(m1) * TILE * TILE;          // scratch of the second half
    if (mt <= TRTRI_FORK_MAX && mt >= 2) {
        SGP_TRY(fork_init());
        cudaStream_t s0 = c.stream;
        cudaStream_t s1 = g_fork_streams[g_fork_next_stream++ % 8];
        cudaEvent_t ef = g_fork_events[g_fork_next_event++ % 64];
        cudaEvent_t ej = g_fork_events[g_fork_next_event++ % 64];
        SGP_CUDA(cudaEventRecord(ef, s0));
        SGP_CUDA(cudaStreamWaitEvent(s1, ef, 0));
        SGP_TRY(trtri_rec(c, A, lda, Dinv, T, j0, m1));
        c.stream = s1;
        const int st = trtri_rec(c, A, lda, Dinv, T2, j0 + m1, m2);
        c.stream = s0;
        SGP_TRY(st);
        SGP_CUDA(cudaEventRecord(ej, s1));
        SGP_CUDA(cudaStreamWaitEvent(s0, ej, 0));
    } else {
        SGP_TRY(trtri_rec(c, A, lda, Dinv, T, j0, m1));
        SGP_TRY(trtri_rec(c, A, lda, Dinv, T2, j0 + m1, m2));
    }
    const long ldt = (long)m2 * TILE;
    // T = L21 * X11        (X11 lower: k >= tn*128)
    SGP_TRY(gemm(c, LAYOUT_MN, LAYOUT_K, AT(A, lda, j0 + m1, j0), lda, AT(A, lda, j0, j0), lda, T, ldt, m2, m1,
                 (long)m1 * TILE, 1.0, 0.0, TM_B_LOWER));
    // L21 = -X22 * T       (X22 lower: k < (tm+1)*128)
    SGP_TRY(gemm(c, LAYOUT_MN, LAYOUT_K, AT(A, lda, j0 + m1, j0 + m1), lda, T, ldt, AT(A, lda, j0 + m1, j0), lda, m2, m1,
                 (long)m2 * TILE, -1.0, 0.0, TM_A_LOWER));
    return ST_OK;
}

// Blocked lower-triangular inverse, last block column first (LAPACK dtrtri order):
//   X_jj = inv(L_jj) (recursive);  T = X[j+1:, j+1:] * L[j+1:, j];  X[j+1:, j] = -T * X_jj
int trtri(Ctx& c, double* A, long n_pad, long lda, const double* Dinv, double* T)
{
    const int nt = (int)(n_pad / TILE);
    const int nb = env_int("SGP_TRTRI_NB", 0, 0, 64);
    if (nb == 0) return trtri_rec(c, A, lda, Dinv, T, 0, nt);                    // fully recursive
    int j0 = ((nt - 1) / nb) * nb;
    for (; j0 >= 0; j0 -= nb) {
        const int jb = (nt - j0 < nb) ? (nt - j0) : nb;
        SGP_TRY(trtri_rec(c, A, lda, Dinv, T, j0, jb));          // needs jb/2 x jb/2 tiles of T
        const int rt = nt - (j0 + jb);
        if (rt > 0) {
            double* T2 = T + (trtri_tiles(nb) + 1) * TILE * TILE;   // past trtri_rec's scratch
            const long ldt = (long)rt * TILE;
            // T2 = X22 * L21          (X22 lower: k < (tm+1)*128)
            SGP_TRY(gemm(c, LAYOUT_MN, LAYOUT_K, AT(A, lda, j0 + jb, j0 + jb), lda, AT(A, lda, j0 + jb, j0), lda, T2, ldt, rt, jb,
                         (long)rt * TILE, 1.0, 0.0, TM_A_LOWER));
            // L21 = -T2 * X_jj        (X_jj lower: k >= tn*128)
            SGP_TRY(gemm(c, LAYOUT_MN, LAYOUT_K, T2, ldt, AT(A, lda, j0, j0), lda, AT(A, lda, j0 + jb, j0), lda, rt, jb,
                         (long)jb * TILE, -1.0, 0.0, TM_B_LOWER));
        }
    }
    return ST_OK;
}

size_t trtri_workspace_doubles(long n_pad)
{
    const size_t nb = 64;                                        // upper bound of SGP_TRTRI_NB
    const size_t rec = (trtri_tiles((int)nb) + 1) * TILE * TILE;
    const size_t blocked = rec + (size_t)n_pad * (nb * TILE);
    const size_t recursive = (trtri_tiles((int)(n_pad / TILE)) + 1) * TILE * TILE;
    return blocked > recursive ? blocked : recursive;
}

int lauum(Ctx& c, const double* X, long n_pad, long lda, double* W, long ldw)
{
    const int nt = (int)(n_pad / TILE);
    return gemm(c, LAYOUT_K, LAYOUT_K, X, lda, X, lda, W, ldw, nt, nt, n_pad, 1.0, 0.0, TM_LOWER_KGE);
}

// ------------------------------------------------------------------------------------------
// self test support: plain FP64 reference GEMM (one thread per output element)
// ------------------------------------------------------------------------------------------
__global__ void ref_gemm_kernel(GemmArgs p, int al, int bl, double* __restrict__ out)
{
    const long M = (long)p.Mt * TILE, N = (long)p.Nt * TILE;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const long m = idx % M, n = idx / M;
    const int tm = (int)(m / TILE), tn = (int)(n / TILE);
    long k0 = 0, k1 = p.K;
    bool skip = false;
    if (p.mode == TM_LOWER || p.mode == TM_LOWER_KGE) skip = tn > tm;
    if (p.mode == TM_LOWER_KGE) k0 = (long)tm * TILE;
    if (p.mode == TM_B_LOWER) k0 = (long)tn * TILE;
    if (p.mode == TM_A_LOWER) { k1 = (long)(tm + 1) * TILE; if (k1 > p.K) k1 = p.K; }
    double* o = out + m + n * p.ldc;
    if (skip) return;
    double s = 0.0;
    for (long k = k0; k < k1; k++) {
        const double a = (al == LAYOUT_MN) ? p.A[m + k * p.lda] : p.A[k + m * p.lda];
        const double b = (bl == LAYOUT_MN) ? p.B[n + k * p.ldb] : p.B[k + n * p.ldb];
        s += a * b;
    }
    *o = (p.beta == 0.0 ? 0.0 : p.beta * (*o)) + p.alpha * s;
}

int ref_gemm(Ctx& c, int al, int bl, const GemmArgs& g, double* out)
{
    const long tot = (long)g.Mt * TILE * (long)g.Nt * TILE;
    ref_gemm_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c.stream>>>(g, al, bl, out);
    SGP_CUDA(cudaGetLastError());
    return ST_OK;
}

// SGP_GEMM=v1 selects the cp.async ring kernel (dmma_gemm.cuh); default is the warp-specialised
// persistent kernel (dmma_gemm_ws.cuh).  Same arguments, same results up to summation order (identical).
int dmma_gemm(Ctx& c, int al, int bl, const GemmArgs& g)
{
    static const int use_v1 = [] { const char* v = getenv("SGP_GEMM"); return (v && v[0] == 'v' && v[1] == '1') ? 1 : 0; }();
    if (use_v1) SGP_CUDA(gemm_launch(al, bl, g, c.stream));
    else SGP_CUDA(gemm_ws_launch(al, bl, g, c.stream));
    count_launch();
    return ST_OK;
}

}  // namespace sgp
