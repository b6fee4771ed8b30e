// Hyper-parameter gradient of the negative log marginal likelihood without materialising dK.
//
// Reference: nll_grad python/02_pert_pendulum/func.py:148-162 builds dense dK/dlx, dK/dly
// (build_dK :80-129, 8 N^2 scalar f2py calls) and takes trace(Kyinv @ dK) through a full n^3
// product.  Here one sweep over the lower triangle of Ky^-1 regenerates each dK entry from the
// closed forms and accumulates, per hyper-parameter theta in {lx, ly, sig},
//      A_theta = sum_ij alpha_i alpha_j dK_theta,ij      (= alpha' dK alpha)
//      B_theta = sum_ij Kyinv_ij dK_theta,ij             (= trace(Kyinv dK))
// so that  dNLL/dtheta = -0.5 A_theta + 0.5 B_theta.  HBM-read bound: 8 n^2 / 2 bytes.
//
// Per-block partial sums are written out and reduced in a fixed order by finalize (nll.cu),
// so results are bit-reproducible run to run.
#include "grad.cuh"

namespace sgp {

constexpr int GR_THREADS = 128;
constexpr int GR_COLS = 32;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// acc[0..2] = A_lx, A_ly, A_sig ; acc[3..5] = B_lx, B_ly, B_sig
__device__ __forceinline__ void block_store(double* acc, double* __restrict__ partial)
{
    __shared__ double red[GR_THREADS / 32][6];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0.0;
        for (int w = 0; w < GR_THREADS / 32; w++) s += red[w][threadIdx.x];
        const long blk = (long)blockIdx.y * gridDim.x + blockIdx.x;
        partial[blk * 6 + threadIdx.x] = s;
    }
}

template <int FAM>
__global__ void __launch_bounds__(GR_THREADS)
grad_hess_kernel(const Pt* __restrict__ pts, long N, HypC h, const double* __restrict__ Kinv, long ld,
                 const double* __restrict__ alpha, double* __restrict__ partial)
{
    __shared__ Pt sa[GR_COLS];
    __shared__ double saq[GR_COLS], saP[GR_COLS];
    const long j0 = (long)blockIdx.y * GR_COLS;
    if (threadIdx.x < GR_COLS) {
        long j = j0 + threadIdx.x;
        Pt z; z.u = 0; z.v = 1; z.y = 0;
        sa[threadIdx.x] = (j < N) ? pts[j] : z;
        saq[threadIdx.x] = (j < N) ? alpha[j] : 0.0;
        saP[threadIdx.x] = (j < N) ? alpha[N + j] : 0.0;
    }
    __syncthreads();
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const long i = (long)blockIdx.x * GR_THREADS + threadIdx.x;
    if (i < N) {
        const Pt b = pts[i];
        const double aqi = alpha[i], aPi = alpha[N + i];
        const int jn = (int)((N - j0 < GR_COLS) ? (N - j0) : GR_COLS);
        for (int jj = 0; jj < jn; jj++) {
            const long j = j0 + jj;
            const Pair<FAM> q(sa[jj], b, h);
            // yx block entry (N+i, j) and its mirror (j, N+i): weight 2
            {
                const double kinv = Kinv[(N + i) + j * ld];
                const double aa = aPi * saq[jj];
                const double dlx = 2.0 * q.kxy_lx(h), dly = 2.0 * q.kxy_ly(h), ds = 2.0 * q.kxy(h);
                acc[0] += aa * dlx; acc[1] += aa * dly; acc[2] += aa * ds;
                acc[3] += kinv * dlx; acc[4] += kinv * dly; acc[5] += kinv * ds;
            }
            if (i >= j) {
                const double wgt = (i == j) ? 1.0 : 2.0;
                const double kxx = Kinv[i + j * ld], kyy = Kinv[(N + i) + (N + j) * ld];
                const double axx = aqi * saq[jj], ayy = aPi * saP[jj];
                const double xlx = wgt * q.kxx_lx(h), xly = wgt * q.kxx_ly(h), xs = wgt * q.kxx(h);
                const double ylx = wgt * q.kyy_lx(h), yly = wgt * q.kyy_ly(h), ys = wgt * q.kyy(h);
                acc[0] += axx * xlx + ayy * ylx; acc[1] += axx * xly + ayy * yly; acc[2] += axx * xs + ayy * ys;
                acc[3] += kxx * xlx + kyy * ylx; acc[4] += kxx * xly + kyy * yly; acc[5] += kxx * xs + kyy * ys;
            }
        }
    }
    block_store(acc, partial);
}

template <int FAM>
__global__ void __launch_bounds__(GR_THREADS)
grad_reg_kernel(const Pt* __restrict__ pts, long N, HypC h, const double* __restrict__ Kinv, long ld,
                const double* __restrict__ alpha, double* __restrict__ partial)
{
    __shared__ Pt sa[GR_COLS];
    __shared__ double sal[GR_COLS];
    const long j0 = (long)blockIdx.y * GR_COLS;
    if (threadIdx.x < GR_COLS) {
        long j = j0 + threadIdx.x;
        Pt z; z.u = 0; z.v = 1; z.y = 0;
        sa[threadIdx.x] = (j < N) ? pts[j] : z;
        sal[threadIdx.x] = (j < N) ? alpha[j] : 0.0;
    }
    __syncthreads();
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const long i = (long)blockIdx.x * GR_THREADS + threadIdx.x;
    if (i < N && i + 1 > j0) {
        const Pt b = pts[i];
        const double ai = alpha[i];
        const int jn = (int)((N - j0 < GR_COLS) ? (N - j0) : GR_COLS);
        for (int jj = 0; jj < jn; jj++) {
            const long j = j0 + jj;
            if (i < j) break;
            const Pair<FAM> q(sa[jj], b, h);
            const double wgt = (i == j) ? 1.0 : 2.0;
            const double kinv = Kinv[i + j * ld];
            const double aa = ai * sal[jj];
            const double dlx = wgt * q.k_lx(h), dly = wgt * q.k_ly(h), ds = wgt * q.k();
            acc[0] += aa * dlx; acc[1] += aa * dly; acc[2] += aa * ds;
            acc[3] += kinv * dlx; acc[4] += kinv * dly; acc[5] += kinv * ds;
        }
    }
    block_store(acc, partial);
}

long grad_num_partials(long N)
{
    const long gx = (N + GR_THREADS - 1) / GR_THREADS, gy = (N + GR_COLS - 1) / GR_COLS;
    return gx * gy;
}

int grad_contract(Ctx& c, int fam, int reg, const Pt* pts, long N, const HypC& h, const double* Kinv, long ld,
                  const double* alpha, double* partial)
{
    dim3 grid((unsigned)((N + GR_THREADS - 1) / GR_THREADS), (unsigned)((N + GR_COLS - 1) / GR_COLS));
#define LAUNCH(KERN, F) KERN<F><<<grid, GR_THREADS, 0, c.stream>>>(pts, N, h, Kinv, ld, alpha, partial)
    if (!reg) {
        switch (fam) {
        case FAM_PRODUCT: LAUNCH(grad_hess_kernel, FAM_PRODUCT); break;
        case FAM_SQ: LAUNCH(grad_hess_kernel, FAM_SQ); break;
        case FAM_SUM: LAUNCH(grad_hess_kernel, FAM_SUM); break;
        default: set_error("unknown kernel family %d", fam); return ST_BADARG;
        }
    } else {
        switch (fam) {
        case FAM_PRODUCT: LAUNCH(grad_reg_kernel, FAM_PRODUCT); break;
        case FAM_SQ: LAUNCH(grad_reg_kernel, FAM_SQ); break;
        case FAM_SUM: LAUNCH(grad_reg_kernel, FAM_SUM); break;
        default: set_error("unknown kernel family %d", fam); return ST_BADARG;
        }
    }
#undef LAUNCH
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

}  // namespace sgp
