// FP64 Cholesky driver (potrf -> potrf_ll.cu), backward substitution (potrs), triangular inverse (trtri) and
// X^T X (lauum) on one B200, built from the structured DMMA GEMM of dmma_gemm_ws.cuh.
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

#include <mutex>
#include <utility>
#include <vector>

namespace sgp {

// diagonal tile j of the matrix <- Dinv_j (ld 128): the leaves of the recursive triangular inverse, one CTA per tile
__global__ void copy_diag_tiles_kernel(double* __restrict__ A, long lda, const double* __restrict__ Dinv)
{
    const long j = blockIdx.x;
    double* dst = A + j * TILE + j * TILE * lda;
    const double* src = Dinv + j * TILE * TILE;
    for (int idx = threadIdx.x; idx < TILE * TILE; idx += blockDim.x) {
        int r = idx & (TILE - 1), c = idx >> 7;
        dst[r + (long)c * lda] = src[idx];
    }
}

// ------------------------------------------------------------------------------------------
// potrs, second half, by blocked substitution with the tile inverses (the forward substitution
// L w = y rides along in potrf_ll's diagonal tasks).
// ------------------------------------------------------------------------------------------
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
static int gemm(Ctx& c, int al, int bl, const double* A, long lda, const double* B, long ldb, double* C, long ldc,
                int Mt, int Nt, long K, double alpha, double beta, int mode)
{
    GemmArgs g;
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    g.Mt = Mt; g.Nt = Nt; g.K = (int)K; g.alpha = alpha; g.beta = beta; g.mode = mode;
    return dmma_gemm(c, al, bl, g);
}

#define AT(A, lda, rt, ct) ((A) + (long)(rt) * TILE + (long)(ct) * TILE * (lda))

// The whole factorisation is the left-looking persistent kernel of potrf_ll.cu, with the forward substitution fused in.
int potrf(Ctx& c, double* A, long n_pad, long lda, double* Dinv, double* logparts, int* info, double* y, double* w)
{
    if (n_pad % TILE || lda % 2) { set_error("potrf: n_pad %ld / lda %ld not tile aligned", n_pad, lda); return ST_BADARG; }
    SGP_TRY(c.flags.reserve(potrf_ll_flag_bytes(n_pad)));
    return potrf_ll(c, A, n_pad, lda, Dinv, logparts, info, c.flags.as<int>(), y, w);
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

// y = A x (mode 0) or y -= A x (mode 1) for a column-major block A (rows x cols, multiples of 128; tri != 0: square and lower
// triangular, the entries above the diagonal are not read).  With X = L^-1 this is the forward substitution w = X z of the
// INT8 route, which never holds L (ozaki_chol.cu); with A = L21 the update z2 -= L21 w1 of its factor-only variant.
// Block (i, j) = 128 rows x 512 columns, rows down the threads (coalesced), partial sums per column chunk written to `part`
// and added in chunk order by the second kernel: fixed summation order.  HBM-read bound.
constexpr int GEMV_CW = 512;
__global__ void __launch_bounds__(256)
gemv_part_kernel(const double* __restrict__ A, long ld, const double* __restrict__ x, double* __restrict__ part, long rows, long cols, int tri)
{
    __shared__ double xs[GEMV_CW];
    __shared__ double hi[TILE];
    const int i = blockIdx.x, j = blockIdx.y, tid = threadIdx.x;
    const long c0 = (long)j * GEMV_CW;
    const long cend = tri ? (long)(i + 1) * TILE : cols;          // columns this row tile reads at all
    if (c0 >= cend) return;
    const long c1 = c0 + GEMV_CW < cend ? c0 + GEMV_CW : cend;
    for (int k = tid; k < GEMV_CW; k += 256) xs[k] = (c0 + k < c1) ? x[c0 + k] : 0.0;
    __syncthreads();
    const int r = tid & (TILE - 1), half = tid >> 7;
    const long row = (long)i * TILE + r;
    const long hb = c0 + half * (GEMV_CW / 2);
    long he = hb + GEMV_CW / 2;
    if (he > c1) he = c1;
    if (tri && he > row + 1) he = row + 1;
    const double* Ar = A + row;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    long cc = hb;
    for (; cc + 3 < he; cc += 4) {
        a0 = fma(Ar[cc * ld], xs[cc - c0], a0);
        a1 = fma(Ar[(cc + 1) * ld], xs[cc + 1 - c0], a1);
        a2 = fma(Ar[(cc + 2) * ld], xs[cc + 2 - c0], a2);
        a3 = fma(Ar[(cc + 3) * ld], xs[cc + 3 - c0], a3);
    }
    for (; cc < he; cc++) a0 = fma(Ar[cc * ld], xs[cc - c0], a0);
    const double s = (a0 + a1) + (a2 + a3);
    if (half) hi[r] = s;
    __syncthreads();
    if (!half) part[(long)j * rows + row] = s + hi[r];
}

__global__ void gemv_sum_kernel(const double* __restrict__ part, double* __restrict__ y, long rows, long cols, int tri, int mode)
{
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const long cend = tri ? (row / TILE + 1) * TILE : cols;
    double s = 0.0;
    for (long j = 0; j * GEMV_CW < cend; j++) s += part[j * rows + row];
    y[row] = mode ? y[row] - s : s;
}

size_t gemv_scratch_doubles(long rows, long cols) { return (size_t)((cols + GEMV_CW - 1) / GEMV_CW) * (size_t)rows; }

int gemv_blocked(Ctx& c, const double* A, long ld, long rows, long cols, int tri, const double* x, double* y, int mode, double* part)
{
    if (rows % TILE || cols % TILE || (tri && rows != cols)) { set_error("gemv_blocked: block %ld x %ld not tile aligned", rows, cols); return ST_BADARG; }
    const dim3 grid((unsigned)(rows / TILE), (unsigned)((cols + GEMV_CW - 1) / GEMV_CW));
    gemv_part_kernel<<<grid, 256, 0, c.stream>>>(A, ld, x, part, rows, cols, tri);
    gemv_sum_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, c.stream>>>(part, y, rows, cols, tri, mode);
    SGP_CUDA(cudaGetLastError());
    count_launch(2);
    return ST_OK;
}

size_t trmv_lower_scratch_doubles(long n_pad) { return gemv_scratch_doubles(n_pad, n_pad); }

int trmv_lower(Ctx& c, const double* X, long n_pad, long ldx, const double* z, double* w, double* part)
{
    return gemv_blocked(c, X, ldx, n_pad, n_pad, 1, z, w, 0, part);
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

// ------------------------------------------------------------------------------------------
// Triangular inverse X = L^-1, in place, by the recursion
//     X = [X11 0; X21 X22],   X11 = inv(L11), X22 = inv(L22) (independent),   X21 = -X22 (L21 X11)
// down to single tiles, whose inverses potrf_ll left in Dinv.  All nodes of one recursion depth are independent,
// so a depth is TWO grouped launches of the persistent DMMA GEMM (T = L21 X11 for every node, then L21 = -X22 T
// for every node) however many nodes it has: 2 ceil(log2 nt) + 1 launches for the whole inverse (17 at
// n = 32768, 13 at n = 8192; the node-by-node recursion took 510 launches and left the GPU mostly idle at the
// bottom of the tree, where a node is a handful of tiles of depth 128).
// ------------------------------------------------------------------------------------------
struct TrtriPlan {
    // key
    const double* A = nullptr; long lda = 0; int nt = 0; const double* T = nullptr;
    // per depth: [first, last) into the entry arrays, tile counts
    std::vector<int> first;          // size depths + 1
    std::vector<long> tiles1, tiles2;
    DBuf d_e1, d_e2;                 // GemmGroupEntry arrays in device memory (step 1 / step 2 of every depth)
};

static void trtri_collect(double* A, long lda, double* T, int j0, int mt, int depth, std::vector<std::vector<std::pair<int, int>>>& lv)
{
    if (mt <= 1) return;
    if ((int)lv.size() <= depth) lv.resize((size_t)depth + 1);
    lv[(size_t)depth].push_back({j0, mt});
    const int m1 = mt / 2;
    trtri_collect(A, lda, T, j0, m1, depth + 1, lv);
    trtri_collect(A, lda, T, j0 + m1, mt - m1, depth + 1, lv);
}

static int trtri_build_plan(Ctx& c, TrtriPlan& pl, double* A, long lda, int nt, double* T)
{
    pl.A = nullptr;
    std::vector<std::vector<std::pair<int, int>>> lv;
    trtri_collect(A, lda, T, 0, nt, 0, lv);
    std::vector<GemmGroupEntry> e1, e2;
    pl.first.assign(1, 0);
    pl.tiles1.clear(); pl.tiles2.clear();
    for (size_t d = 0; d < lv.size(); d++) {
        long t1 = 0, t2 = 0;
        size_t toff = 0;                                         // nodes of one depth get disjoint scratch regions
        for (auto& nd : lv[d]) {
            const int j0 = nd.first, mt = nd.second, m1 = mt / 2, m2 = mt - m1;
            double* Tn = T + toff * TILE * TILE;
            toff += (size_t)m1 * m2;
            const long ldt = (long)m2 * TILE;
            GemmGroupEntry a{}, b{};
            // T = L21 * X11        (X11 lower: k >= tn*128)
            a.a.A = AT(A, lda, j0 + m1, j0); a.a.lda = lda; a.a.B = AT(A, lda, j0, j0); a.a.ldb = lda; a.a.C = Tn; a.a.ldc = ldt;
            a.a.Mt = m2; a.a.Nt = m1; a.a.K = m1 * TILE; a.a.alpha = 1.0; a.a.beta = 0.0; a.a.mode = TM_B_LOWER; a.tile0 = t1;
            // L21 = -X22 * T       (X22 lower: k < (tm+1)*128)
            b.a.A = AT(A, lda, j0 + m1, j0 + m1); b.a.lda = lda; b.a.B = Tn; b.a.ldb = ldt; b.a.C = AT(A, lda, j0 + m1, j0); b.a.ldc = lda;
            b.a.Mt = m2; b.a.Nt = m1; b.a.K = m2 * TILE; b.a.alpha = -1.0; b.a.beta = 0.0; b.a.mode = TM_A_LOWER; b.tile0 = t2;
            t1 += (long)m2 * m1; t2 += (long)m2 * m1;
            e1.push_back(a); e2.push_back(b);
        }
        pl.first.push_back((int)e1.size());
        pl.tiles1.push_back(t1); pl.tiles2.push_back(t2);
    }
    if (!e1.empty()) {
        SGP_TRY(pl.d_e1.reserve(e1.size() * sizeof(GemmGroupEntry)));
        SGP_TRY(pl.d_e2.reserve(e2.size() * sizeof(GemmGroupEntry)));
        // pageable source: the copy is staged before the call returns, the vectors may go out of scope
        SGP_CUDA(cudaMemcpyAsync(pl.d_e1.p, e1.data(), e1.size() * sizeof(GemmGroupEntry), cudaMemcpyHostToDevice, c.stream));
        SGP_CUDA(cudaMemcpyAsync(pl.d_e2.p, e2.data(), e2.size() * sizeof(GemmGroupEntry), cudaMemcpyHostToDevice, c.stream));
    }
    pl.A = A; pl.lda = lda; pl.nt = nt; pl.T = T;
    return ST_OK;
}

// scratch (in 128x128 tiles): the products T of all nodes of one depth at the same time; the root alone needs
// floor(nt/2) ceil(nt/2) tiles and no deeper level needs more
static size_t trtri_tiles(int nt)
{
    return (size_t)(nt / 2) * (size_t)(nt - nt / 2);
}

int trtri(Ctx& c, double* A, long n_pad, long lda, const double* Dinv, double* T)
{
    const int nt = (int)(n_pad / TILE);
    copy_diag_tiles_kernel<<<nt, 256, 0, c.stream>>>(A, lda, Dinv);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    if (nt == 1) return ST_OK;
    // the plan (per-depth problem lists in device memory) depends only on the pointers and the order; a few are kept (the
    // INT8 route inverts its leaf blocks one after the other, ozaki_chol.cu)
    constexpr int NPLAN = 48;
    static TrtriPlan plans[NPLAN];
    static int plan_dev[NPLAN];
    static int next_plan = 0;
    static std::mutex mu;                                 // contexts of several host threads share the cache
    std::lock_guard<std::mutex> lock(mu);
    TrtriPlan* found = nullptr;
    for (int i = 0; i < NPLAN && !found; i++)
        if (plans[i].A == A && plans[i].lda == lda && plans[i].nt == nt && plans[i].T == T && plan_dev[i] == c.device) found = &plans[i];
    if (!found) {
        found = &plans[next_plan];
        plan_dev[next_plan] = c.device;
        next_plan = (next_plan + 1) % NPLAN;
        if (found->d_e1.p) SGP_CUDA(cudaDeviceSynchronize());        // the entry being replaced may still be in use
        SGP_TRY(trtri_build_plan(c, *found, A, lda, nt, T));
    }
    TrtriPlan& pl = *found;
    const GemmGroupEntry* e1 = pl.d_e1.as<GemmGroupEntry>();
    const GemmGroupEntry* e2 = pl.d_e2.as<GemmGroupEntry>();
    for (int d = (int)pl.tiles1.size() - 1; d >= 0; d--) {          // deepest nodes first
        const int f = pl.first[(size_t)d], np = pl.first[(size_t)d + 1] - f;
        SGP_CUDA((gemm_ws_launch_grouped_t<LAYOUT_MN, LAYOUT_K>(e1 + f, np, pl.tiles1[(size_t)d], c.stream)));
        SGP_CUDA((gemm_ws_launch_grouped_t<LAYOUT_MN, LAYOUT_K>(e2 + f, np, pl.tiles2[(size_t)d], c.stream)));
        count_launch(2);
    }
    return ST_OK;
}

size_t trtri_workspace_doubles(long n_pad)
{
    return (trtri_tiles((int)(n_pad / TILE)) + 1) * TILE * TILE;
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

int dmma_gemm(Ctx& c, int al, int bl, const GemmArgs& g)
{
    SGP_CUDA(gemm_ws_launch(al, bl, g, c.stream));
    count_launch();
    return ST_OK;
}

}  // namespace sgp
