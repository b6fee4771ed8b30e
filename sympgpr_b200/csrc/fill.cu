// Kernel-matrix fills.
//
// Replaces sympgpr.f90::build_K (:12-38) and ::buildKreg (:40-60) and their Python-loop twins
// (python/04_standard_map/func.py:32-68).  K is column-major; rows are the "x" points (b),
// columns the "x0" points (a).  Block layout of build_K:
//     K(i, j) = kxx   K(i, N0+j) = kxy
//     K(N+i,j) = kxy  K(N+i,N0+j) = kyy        all times hyp(3)
//
// HBM-write bound: every thread owns two consecutive rows and walks over a strip of columns,
// so each block value leaves as one 128-bit store and a warp writes 512 contiguous bytes per
// store instruction.  One exp per (i,j) pair serves all four blocks; sin/cos never appear
// per pair (forms.cuh).
#include "fill.cuh"

namespace sgp {

constexpr int FILL_THREADS = 128;      // rows per block = 256
constexpr int FILL_COLS = 32;          // columns per block

template <int FAM>
__global__ void points_kernel(const double* __restrict__ x, const double* __restrict__ y, long n, double p,
                              Pt* __restrict__ out)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_pt<FAM>(x[i], y[i], p);
}

int make_points(Ctx& c, int fam, double p, const double* x, const double* y, long n, Pt* out)
{
    if (n <= 0) return ST_OK;
    const unsigned g = (unsigned)((n + 255) / 256);
    if (fam == FAM_SQ) points_kernel<FAM_SQ><<<g, 256, 0, c.stream>>>(x, y, n, p, out);
    else points_kernel<FAM_PRODUCT><<<g, 256, 0, c.stream>>>(x, y, n, p, out);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

__device__ __forceinline__ void st2(double* p, double a, double b, bool vec)
{
    if (vec) *reinterpret_cast<double2*>(p) = make_double2(a, b);
    else { p[0] = a; p[1] = b; }
}

// General rectangular fill (build_K): rows 2N (points pb), cols 2N0 (points pa).
template <int FAM>
__global__ void __launch_bounds__(FILL_THREADS)
fill_hess_kernel(const Pt* __restrict__ pb, long N, const Pt* __restrict__ pa, long N0, HypC h, double* __restrict__ K,
                 long ld, int vec)
{
    __shared__ Pt sa[FILL_COLS];
    const long j0 = (long)blockIdx.y * FILL_COLS;
    if (threadIdx.x < FILL_COLS) {
        long j = j0 + threadIdx.x;
        Pt z; z.u = 0; z.v = 1; z.y = 0;
        sa[threadIdx.x] = (j < N0) ? pa[j] : z;
    }
    __syncthreads();
    const long i = ((long)blockIdx.x * FILL_THREADS + threadIdx.x) * 2;
    if (i >= N) return;
    const bool two = (i + 1 < N);
    const Pt b0 = pb[i];
    const Pt b1 = two ? pb[i + 1] : b0;
    const int jn = (int)((N0 - j0 < FILL_COLS) ? (N0 - j0) : FILL_COLS);
    const double sig = h.sig;
    const bool v2 = vec && two;
#pragma unroll 2
    for (int jj = 0; jj < jn; jj++) {
        const Pt a = sa[jj];
        const Pair<FAM> q0(a, b0, h), q1(a, b1, h);
        const double xx0 = sig * q0.kxx(h), xy0 = sig * q0.kxy(h), yy0 = sig * q0.kyy(h);
        const double xx1 = sig * q1.kxx(h), xy1 = sig * q1.kxy(h), yy1 = sig * q1.kyy(h);
        const long j = j0 + jj;
        double* c0 = K + i + j * ld;
        double* c1 = K + i + (N0 + j) * ld;
        if (v2) {
            st2(c0, xx0, xx1, true);
            st2(c0 + N, xy0, xy1, true);
            st2(c1, xy0, xy1, true);
            st2(c1 + N, yy0, yy1, true);
        } else {
            c0[0] = xx0; c0[N] = xy0; c1[0] = xy0; c1[N] = yy0;
            if (two) { c0[1] = xx1; c0[N + 1] = xy1; c1[1] = xy1; c1[N + 1] = yy1; }
        }
    }
}

int fill_hess(Ctx& c, int fam, const Pt* pb, long N, const Pt* pa, long N0, const HypC& h, double* K, long ld)
{
    if (N <= 0 || N0 <= 0) return ST_OK;
    dim3 grid((unsigned)((N + 2 * FILL_THREADS - 1) / (2 * FILL_THREADS)), (unsigned)((N0 + FILL_COLS - 1) / FILL_COLS));
    const int vec = (N % 2 == 0) && (ld % 2 == 0) && (((size_t)K & 15) == 0);
    switch (fam) {
    case FAM_PRODUCT: fill_hess_kernel<FAM_PRODUCT><<<grid, FILL_THREADS, 0, c.stream>>>(pb, N, pa, N0, h, K, ld, vec); break;
    case FAM_SQ: fill_hess_kernel<FAM_SQ><<<grid, FILL_THREADS, 0, c.stream>>>(pb, N, pa, N0, h, K, ld, vec); break;
    case FAM_SUM: fill_hess_kernel<FAM_SUM><<<grid, FILL_THREADS, 0, c.stream>>>(pb, N, pa, N0, h, K, ld, vec); break;
    default: set_error("unknown kernel family %d", fam); return ST_BADARG;
    }
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

// Symmetric training fill for the Cholesky path: x == x0, Ky = K + noise*I, only what potrf
// reads is written: the yx quadrant completely, xx and yy on and below the diagonal (at the
// granularity of this kernel's 256 x 32 strips), i.e. half the bytes of the full matrix.
template <int FAM>
__global__ void __launch_bounds__(FILL_THREADS)
fill_hess_sym_kernel(const Pt* __restrict__ pts, long N, HypC h, double noise, double* __restrict__ K, long ld, int vec)
{
    __shared__ Pt sa[FILL_COLS];
    const long j0 = (long)blockIdx.y * FILL_COLS;
    if (threadIdx.x < FILL_COLS) {
        long j = j0 + threadIdx.x;
        Pt z; z.u = 0; z.v = 1; z.y = 0;
        sa[threadIdx.x] = (j < N) ? pts[j] : z;
    }
    __syncthreads();
    const long i = ((long)blockIdx.x * FILL_THREADS + threadIdx.x) * 2;
    if (i >= N) return;
    const bool two = (i + 1 < N);
    const Pt b0 = pts[i];
    const Pt b1 = two ? pts[i + 1] : b0;
    const int jn = (int)((N - j0 < FILL_COLS) ? (N - j0) : FILL_COLS);
    const double sig = h.sig;
    const bool v2 = vec && two;
    // the whole strip of this thread lies strictly above the diagonal -> only the yx block
    const bool diag_blocks = (i + 1 >= j0);
#pragma unroll 2
    for (int jj = 0; jj < jn; jj++) {
        const Pt a = sa[jj];
        const Pair<FAM> q0(a, b0, h), q1(a, b1, h);
        const long j = j0 + jj;
        const double xy0 = sig * q0.kxy(h), xy1 = sig * q1.kxy(h);
        double* c0 = K + i + j * ld;
        if (v2) st2(c0 + N, xy0, xy1, true);
        else { c0[N] = xy0; if (two) c0[N + 1] = xy1; }
        if (diag_blocks) {
            double xx0 = sig * q0.kxx(h), yy0 = sig * q0.kyy(h);
            double xx1 = sig * q1.kxx(h), yy1 = sig * q1.kyy(h);
            if (i == j) { xx0 += noise; yy0 += noise; }
            if (i + 1 == j) { xx1 += noise; yy1 += noise; }
            double* c1 = K + (N + i) + (N + j) * ld;
            if (v2) { st2(c0, xx0, xx1, true); st2(c1, yy0, yy1, true); }
            else { c0[0] = xx0; c1[0] = yy0; if (two) { c0[1] = xx1; c1[1] = yy1; } }
        }
    }
}

// identity on the padding rows/cols [n, n_pad) (lower part only)
__global__ void pad_identity_kernel(double* __restrict__ K, long ld, long n, long n_pad)
{
    const long npadrows = n_pad - n;
    const long tot = npadrows * n_pad;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (long)gridDim.x * blockDim.x) {
        const long r = n + idx % npadrows, col = idx / npadrows;
        K[r + col * ld] = (r == col) ? 1.0 : 0.0;
    }
}

int fill_hess_sym(Ctx& c, int fam, const Pt* pts, long N, const HypC& h, double noise, double* K, long ld, long n_pad)
{
    dim3 grid((unsigned)((N + 2 * FILL_THREADS - 1) / (2 * FILL_THREADS)), (unsigned)((N + FILL_COLS - 1) / FILL_COLS));
    const int vec = (N % 2 == 0) && (ld % 2 == 0) && (((size_t)K & 15) == 0);
    switch (fam) {
    case FAM_PRODUCT: fill_hess_sym_kernel<FAM_PRODUCT><<<grid, FILL_THREADS, 0, c.stream>>>(pts, N, h, noise, K, ld, vec); break;
    case FAM_SQ: fill_hess_sym_kernel<FAM_SQ><<<grid, FILL_THREADS, 0, c.stream>>>(pts, N, h, noise, K, ld, vec); break;
    case FAM_SUM: fill_hess_sym_kernel<FAM_SUM><<<grid, FILL_THREADS, 0, c.stream>>>(pts, N, h, noise, K, ld, vec); break;
    default: set_error("unknown kernel family %d", fam); return ST_BADARG;
    }
    SGP_CUDA(cudaGetLastError());
    count_launch();
    if (n_pad > 2 * N) {
        pad_identity_kernel<<<256, 256, 0, c.stream>>>(K, ld, 2 * N, n_pad);
        count_launch();
        SGP_CUDA(cudaGetLastError());
    }
    return ST_OK;
}

// Plain kernel matrix (buildKreg): K(i,j) = sig * k(a_j, b_i); rows N, cols N0.
template <int FAM>
__global__ void __launch_bounds__(FILL_THREADS)
fill_reg_kernel(const Pt* __restrict__ pb, long N, const Pt* __restrict__ pa, long N0, HypC h, double noise, int sym,
                double* __restrict__ K, long ld, int vec, int which)
{
    __shared__ Pt sa[FILL_COLS];
    const long j0 = (long)blockIdx.y * FILL_COLS;
    if (threadIdx.x < FILL_COLS) {
        long j = j0 + threadIdx.x;
        Pt z; z.u = 0; z.v = 1; z.y = 0;
        sa[threadIdx.x] = (j < N0) ? pa[j] : z;
    }
    __syncthreads();
    const long i = ((long)blockIdx.x * FILL_THREADS + threadIdx.x) * 2;
    if (i >= N) return;
    if (sym && i + 1 < j0) return;           // strictly upper strip
    const bool two = (i + 1 < N);
    const Pt b0 = pb[i];
    const Pt b1 = two ? pb[i + 1] : b0;
    const int jn = (int)((N0 - j0 < FILL_COLS) ? (N0 - j0) : FILL_COLS);
    const bool v2 = vec && two;
    for (int jj = 0; jj < jn; jj++) {
        const Pt a = sa[jj];
        const Pair<FAM> q0(a, b0, h), q1(a, b1, h);
        const long j = j0 + jj;
        // which: 0 kernel value (buildKreg); 1 / 2 the (q,q) / (P,P) Hessian block alone (nll_expl,
        // python/04_standard_map/func.py:126-141: the sum kernel's matrix is block diagonal)
        double k0, k1;
        if (which == 0) { k0 = h.sig * q0.k(); k1 = h.sig * q1.k(); }
        else if (which == 1) { k0 = h.sig * q0.kxx(h); k1 = h.sig * q1.kxx(h); }
        else { k0 = h.sig * q0.kyy(h); k1 = h.sig * q1.kyy(h); }
        if (sym) { if (i == j) k0 += noise; if (i + 1 == j) k1 += noise; }
        double* c0 = K + i + j * ld;
        if (v2) st2(c0, k0, k1, true);
        else { c0[0] = k0; if (two) c0[1] = k1; }
    }
}

int fill_reg(Ctx& c, int fam, const Pt* pb, long N, const Pt* pa, long N0, const HypC& h, double* K, long ld)
{
    if (N <= 0 || N0 <= 0) return ST_OK;
    dim3 grid((unsigned)((N + 2 * FILL_THREADS - 1) / (2 * FILL_THREADS)), (unsigned)((N0 + FILL_COLS - 1) / FILL_COLS));
    const int vec = (ld % 2 == 0) && (((size_t)K & 15) == 0);
    switch (fam) {
    case FAM_PRODUCT: fill_reg_kernel<FAM_PRODUCT><<<grid, FILL_THREADS, 0, c.stream>>>(pb, N, pa, N0, h, 0.0, 0, K, ld, vec, 0); break;
    case FAM_SQ: fill_reg_kernel<FAM_SQ><<<grid, FILL_THREADS, 0, c.stream>>>(pb, N, pa, N0, h, 0.0, 0, K, ld, vec, 0); break;
    case FAM_SUM: fill_reg_kernel<FAM_SUM><<<grid, FILL_THREADS, 0, c.stream>>>(pb, N, pa, N0, h, 0.0, 0, K, ld, vec, 0); break;
    default: set_error("unknown kernel family %d", fam); return ST_BADARG;
    }
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

int fill_reg_sym(Ctx& c, int fam, const Pt* pts, long N, const HypC& h, double noise, double* K, long ld, long n_pad, int which)
{
    dim3 grid((unsigned)((N + 2 * FILL_THREADS - 1) / (2 * FILL_THREADS)), (unsigned)((N + FILL_COLS - 1) / FILL_COLS));
    const int vec = (ld % 2 == 0) && (((size_t)K & 15) == 0);
    switch (fam) {
    case FAM_PRODUCT: fill_reg_kernel<FAM_PRODUCT><<<grid, FILL_THREADS, 0, c.stream>>>(pts, N, pts, N, h, noise, 1, K, ld, vec, which); break;
    case FAM_SQ: fill_reg_kernel<FAM_SQ><<<grid, FILL_THREADS, 0, c.stream>>>(pts, N, pts, N, h, noise, 1, K, ld, vec, which); break;
    case FAM_SUM: fill_reg_kernel<FAM_SUM><<<grid, FILL_THREADS, 0, c.stream>>>(pts, N, pts, N, h, noise, 1, K, ld, vec, which); break;
    default: set_error("unknown kernel family %d", fam); return ST_BADARG;
    }
    SGP_CUDA(cudaGetLastError());
    count_launch();
    if (n_pad > N) {
        pad_identity_kernel<<<256, 256, 0, c.stream>>>(K, ld, N, n_pad);
        count_launch();
        SGP_CUDA(cudaGetLastError());
    }
    return ST_OK;
}

}  // namespace sgp
