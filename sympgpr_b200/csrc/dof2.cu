// 2-degree-of-freedom generalisation of the Hessian-block kernel matrix: 4 x 4 blocks.
//
// NOT IN THE REFERENCE.  BASELINE.json's config 3 names a "2-DOF (4D phase space) 4x4-block Hessian kernel";
// the reference reduces Henon-Heiles to a 2-D section map with 2 x 2 blocks (python/03_henon_heiles/
// main.py:91-106).  The construction below is the direct generalisation of build_K (sympgpr.f90:12-38,
// "Eq. (38)") to the generating function F(q1, q2, P1, P2) with the squared-exponential kernel of
// python/03_henon_heiles/init_func.py:24-28 in all four mixed variables u = (q1, q2, P1, P2):
//     k(u, u') = exp(-sum_c (u_c - u'_c)^2 / (2 l_c^2)),   l = (lq, lq, lP, lP)
//     K[(a N + i), (b N0 + j)] = sig d^2 k / du_a du'_b = sig (delta_ab / l_a^2 - D_a D_b / (l_a^2 l_b^2)) k,
//     D_c = u_c(x_i) - u_c(x0_j)
// with observations z = [p1 - P1; p2 - P2; Q1 - q1; Q2 - q2].  Parity is unpinned (no reference code): the oracle
// twin (oracle/oracle.py build_k4 / nll_grad4) is validated by finite differences and by its 2 x 2 sub-blocks
// reproducing the reference's SE x SE matrix.  This file is the training side (fill, NLL, gradient); the prediction
// side (2 x 2 Newton map kernel) is map4_kernel in map.cu.
#include "dof2.cuh"

namespace sgp {

constexpr int D4_THREADS = 128;
constexpr int D4_COLS = 16;

struct Hyp4 {
    double g[4];        // 1 / l_a^2
    double sig;
    double ilq3, ilP3;  // 1 / lq^3, 1 / lP^3
    double ilq, ilP;    // 1 / lq, 1 / lP
};

static Hyp4 make_hyp4(double lq, double lP, double sig)
{
    Hyp4 h;
    h.g[0] = h.g[1] = 1.0 / (lq * lq);
    h.g[2] = h.g[3] = 1.0 / (lP * lP);
    h.sig = sig;
    h.ilq3 = 1.0 / (lq * lq * lq); h.ilP3 = 1.0 / (lP * lP * lP);
    h.ilq = 1.0 / lq; h.ilP = 1.0 / lP;
    return h;
}

// rows: points x (4 arrays of N), cols: points x0 (4 arrays of N0).  sym: x == x0, write only what a lower
// Cholesky reads (blocks a > b completely, diagonal blocks for i >= j at strip granularity), add noise on the diagonal.
__global__ void __launch_bounds__(D4_THREADS)
fill4_kernel(const double* __restrict__ x, long N, const double* __restrict__ x0, long N0, Hyp4 h, double noise, int sym,
             double* __restrict__ K, long ld)
{
    __shared__ double sa[4][D4_COLS];
    const long j0 = (long)blockIdx.y * D4_COLS;
    if (threadIdx.x < 4 * D4_COLS) {
        const int c = threadIdx.x / D4_COLS, jj = threadIdx.x % D4_COLS;
        const long j = j0 + jj;
        sa[c][jj] = (j < N0) ? x0[c * N0 + j] : 0.0;
    }
    __syncthreads();
    const long i = (long)blockIdx.x * D4_THREADS + threadIdx.x;
    if (i >= N) return;
    double u[4];
#pragma unroll
    for (int c = 0; c < 4; c++) u[c] = x[c * N + i];
    const int jn = (int)((N0 - j0 < D4_COLS) ? (N0 - j0) : D4_COLS);
    const bool diag_blocks = !sym || (i >= j0);
    for (int jj = 0; jj < jn; jj++) {
        const long j = j0 + jj;
        double D[4], e = 0.0;
#pragma unroll
        for (int c = 0; c < 4; c++) { D[c] = u[c] - sa[c][jj]; e = fma(D[c] * D[c], h.g[c], e); }
        const double E = h.sig * exp_neg(-0.5 * e);
        double w[4];
#pragma unroll
        for (int c = 0; c < 4; c++) w[c] = D[c] * h.g[c];
#pragma unroll
        for (int a = 0; a < 4; a++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                if (sym && (b > a || (a == b && !diag_blocks))) continue;
                if (sym && a == b && i < j) continue;
                double v = -(w[a] * w[b]) * E;
                if (a == b) { v += h.g[a] * E; if (sym && i == j) v += noise; }
                K[(a * N + i) + (b * N0 + j) * ld] = v;
            }
        }
    }
}

// identity on the padding rows/cols [n, n_pad) (lower part only)
__global__ void pad_identity4_kernel(double* __restrict__ K, long ld, long n, long n_pad)
{
    const long npadrows = n_pad - n;
    const long tot = npadrows * n_pad;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (long)gridDim.x * blockDim.x) {
        const long r = n + idx % npadrows, col = idx / npadrows;
        K[r + col * ld] = (r == col) ? 1.0 : 0.0;
    }
}

int fill4(Ctx& c, const double* x, long N, const double* x0, long N0, double lq, double lP, double sig, double* K, long ld)
{
    if (N <= 0 || N0 <= 0) return ST_OK;
    dim3 grid((unsigned)((N + D4_THREADS - 1) / D4_THREADS), (unsigned)((N0 + D4_COLS - 1) / D4_COLS));
    fill4_kernel<<<grid, D4_THREADS, 0, c.stream>>>(x, N, x0, N0, make_hyp4(lq, lP, sig), 0.0, 0, K, ld);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

int fill4_sym(Ctx& c, const double* x, long N, double lq, double lP, double sig, double noise, double* K, long ld, long n_pad)
{
    dim3 grid((unsigned)((N + D4_THREADS - 1) / D4_THREADS), (unsigned)((N + D4_COLS - 1) / D4_COLS));
    fill4_kernel<<<grid, D4_THREADS, 0, c.stream>>>(x, N, x, N, make_hyp4(lq, lP, sig), noise, 1, K, ld);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    if (n_pad > 4 * N) {
        pad_identity4_kernel<<<256, 256, 0, c.stream>>>(K, ld, 4 * N, n_pad);
        SGP_CUDA(cudaGetLastError());
        count_launch();
    }
    return ST_OK;
}

// Gradient contraction (the 2-DOF twin of grad_hess_kernel): one sweep over the lower triangle of Kyinv,
//   A_theta = alpha' dK_theta alpha,  B_theta = trace(Kyinv dK_theta),  theta in {lq, lP, sig}, with
//   d/dl_g of block (a,b) = sig k [ -2 [a in g] delta_ab g_a / l_g + 2 ([a in g] + [b in g]) D_a D_b g_a g_b / l_g
//                                   + (delta_ab g_a - D_a D_b g_a g_b) S_g / l_g^3 ],   S_g = sum_{c in g} D_c^2.
// partial[(blk * 6) + k]: k = 0..2 A_lq, A_lP, A_sig (A_sig, B_sig as K/sig like grad.cu); 3..5 B.
__global__ void __launch_bounds__(D4_THREADS)
grad4_kernel(const double* __restrict__ x, long N, Hyp4 h, const double* __restrict__ Kinv, long ld,
             const double* __restrict__ alpha, double* __restrict__ partial)
{
    __shared__ double sa[4][D4_COLS];
    __shared__ double sal[4][D4_COLS];
    __shared__ double red[D4_THREADS / 32][6];
    const long j0 = (long)blockIdx.y * D4_COLS;
    if (threadIdx.x < 4 * D4_COLS) {
        const int c = threadIdx.x / D4_COLS, jj = threadIdx.x % D4_COLS;
        const long j = j0 + jj;
        sa[c][jj] = (j < N) ? x[c * N + j] : 0.0;
        sal[c][jj] = (j < N) ? alpha[c * N + j] : 0.0;
    }
    __syncthreads();
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const long i = (long)blockIdx.x * D4_THREADS + threadIdx.x;
    if (i < N) {
        double u[4], al[4];
#pragma unroll
        for (int c = 0; c < 4; c++) { u[c] = x[c * N + i]; al[c] = alpha[c * N + i]; }
        const int jn = (int)((N - j0 < D4_COLS) ? (N - j0) : D4_COLS);
        for (int jj = 0; jj < jn; jj++) {
            const long j = j0 + jj;
            double D[4], w[4], e = 0.0;
#pragma unroll
            for (int c = 0; c < 4; c++) { D[c] = u[c] - sa[c][jj]; w[c] = D[c] * h.g[c]; e = fma(D[c], w[c], e); }
            const double E = exp_neg(-0.5 * e);
            const double Sq = (D[0] * D[0] + D[1] * D[1]) * h.ilq3, SP = (D[2] * D[2] + D[3] * D[3]) * h.ilP3;
#pragma unroll
            for (int a = 0; a < 4; a++) {
#pragma unroll
                for (int b = 0; b <= a; b++) {
                    // lower triangle of the full matrix: block a > b all (i,j); diagonal blocks i >= j
                    if (a == b && i < j) continue;
                    const double wgt = (a == b && i == j) ? 1.0 : 2.0;
                    const double ww = w[a] * w[b];
                    const double base = ((a == b) ? h.g[a] : 0.0) - ww;            // block / (sig k)
                    const int qa = a < 2, qb = b < 2;
                    double dq = base * Sq, dP = base * SP;
                    dq += (2.0 * (qa + qb) * ww - ((a == b && qa) ? 2.0 * h.g[a] : 0.0)) * h.ilq;
                    dP += (2.0 * ((1 - qa) + (1 - qb)) * ww - ((a == b && !qa) ? 2.0 * h.g[a] : 0.0)) * h.ilP;
                    const double kinv = Kinv[(a * N + i) + (b * N + j) * ld];
                    const double aa = al[a] * sal[b][jj];
                    const double f = wgt * E;
                    acc[0] += aa * (f * dq); acc[1] += aa * (f * dP); acc[2] += aa * (f * base);
                    acc[3] += kinv * (f * dq); acc[4] += kinv * (f * dP); acc[5] += kinv * (f * base);
                }
            }
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double s = 0.0;
        for (int w2 = 0; w2 < D4_THREADS / 32; w2++) s += red[w2][threadIdx.x];
        const long blk = (long)blockIdx.y * gridDim.x + blockIdx.x;
        partial[blk * 6 + threadIdx.x] = s;
    }
}

long grad4_num_partials(long N)
{
    return ((N + D4_THREADS - 1) / D4_THREADS) * ((N + D4_COLS - 1) / D4_COLS);
}

// the finalisation of nll.cu multiplies slots 0,1,3,4 by sig and takes 2,5 as they are (dK_sig = K / sig)
int grad4_contract(Ctx& c, const double* x, long N, double lq, double lP, double sig, const double* Kinv, long ld,
                   const double* alpha, double* partial)
{
    dim3 grid((unsigned)((N + D4_THREADS - 1) / D4_THREADS), (unsigned)((N + D4_COLS - 1) / D4_COLS));
    grad4_kernel<<<grid, D4_THREADS, 0, c.stream>>>(x, N, make_hyp4(lq, lP, sig), Kinv, ld, alpha, partial);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

}  // namespace sgp
