// NLL, NLL+gradient and model fit on one GPU.
//
// Reference call sites this pipeline replaces as a whole:
//   nll_chol      python/05_tokamak/SympGPR/func.py:143-150     (fill, + noise, potrf, potrs, value)
//   nll_chol_reg  python/05_tokamak/SympGPR/func.py:134-141
//   nll_grad      python/02_pert_pendulum/func.py:148-162       (+ inverse, dK, traces)
//   nll_grad_reg  python/02_pert_pendulum/func.py:132-146
//   Kyinv = scipy.linalg.inv(K + sig2_n I)  python/01_pendulum/implicit/main.py:138-140,159-161
#include "nll.cuh"

#include <float.h>

#include "chol.cuh"
#include "fill.cuh"
#include "dof2.cuh"
#include "grad.cuh"
#include "ozaki.cuh"

namespace sgp {

__global__ void finalize_kernel(const double* __restrict__ z, const double* __restrict__ alpha, long n,
                                const double* __restrict__ logparts, int nt, const int* __restrict__ info,
                                const double* __restrict__ partial, long npart, double sig, int ngrad,
                                double* __restrict__ res)
{
    __shared__ double red[256];
    const int tid = threadIdx.x;
    // 0.5 * z' alpha, fixed summation order
    double s = 0.0;
    for (long i = tid; i < n; i += 256) s += z[i] * alpha[i];
    red[tid] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    const double quad = 0.5 * red[0];
    __syncthreads();
    double g[6] = {0, 0, 0, 0, 0, 0};
    if (ngrad > 0) {
        for (int k = 0; k < 6; k++) {
            double a = 0.0;
            for (long b = tid; b < npart; b += 256) a += partial[b * 6 + k];
            red[tid] = a;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if (tid < o) red[tid] += red[tid + o];
                __syncthreads();
            }
            g[k] = red[0];
            __syncthreads();
        }
    }
    if (tid == 0) {
        double ld = 0.0;
        for (int k = 0; k < nt; k++) ld += logparts[k];
        res[0] = quad + ld;
        res[5] = quad;
        res[6] = ld;
        res[4] = (double)(*info);
        res[7] = 0.0;
        // dK_theta = sig * d3k..., dK_sig = K / sig
        const double A[3] = {sig * g[0], sig * g[1], g[2]};
        const double B[3] = {sig * g[3], sig * g[4], g[5]};
        for (int k = 0; k < 3; k++) {
            res[1 + k] = (ngrad > k) ? (-0.5 * A[k] + 0.5 * B[k]) : 0.0;
            res[8 + k] = A[k];
            res[11 + k] = B[k];
        }
        res[14] = res[15] = 0.0;
    }
}

// dst (n x n, ld n) = symmetric completion of the lower triangle of src (ld lds)
__global__ void sym_out_kernel(const double* __restrict__ src, long lds, double* __restrict__ dst, long n)
{
    const long tot = n * n;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx % n, c = idx / n;
        dst[idx] = (r >= c) ? src[r + c * lds] : src[c + r * lds];
    }
}

// dst (n x n, ld n) = lower triangle of src, zeros above
__global__ void tril_out_kernel(const double* __restrict__ src, long lds, double* __restrict__ dst, long n)
{
    const long tot = n * n;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < tot; idx += (long)gridDim.x * blockDim.x) {
        const long r = idx % n, c = idx / n;
        dst[idx] = (r >= c) ? src[r + c * lds] : 0.0;
    }
}

int nll_enqueue(Ctx& c, const NllJob& job)
{
    const long n = job.n;
    if (n <= 0) { set_error("nll: empty problem (n = %ld)", n); return ST_BADARG; }
    if (!job.reg && (n % 2)) { set_error("nll: derivative-kernel order must be even, got %ld", n); return ST_BADARG; }
    if (job.reg == 4 && (n % 4)) { set_error("nll: the order of the 2-DOF derivative kernel must be a multiple of 4, got %ld", n); return ST_BADARG; }
    const bool dof2 = job.reg == 4;
    const long N = dof2 ? n / 4 : (job.reg ? n : n / 2);
    const long n_pad = round_up(n, TILE);
    const int nt = (int)(n_pad / TILE);
    const int fam = job.fam;
    if (fam < 0 || fam > 2) { set_error("unknown kernel family %d", fam); return ST_BADARG; }
    if (job.reg < 0 || job.reg > 4 || ((job.reg == 2 || job.reg == 3) && job.ngrad > 0)) { set_error("nll: reg must be 0..4 (no gradient for 2 and 3)"); return ST_BADARG; }
    // the kernels depend on l^2 only (kernels.f90:9-10) and their l-derivatives are odd in l, so a negative length scale
    // is a valid argument exactly as in the reference (an optimiser working on raw hyper-parameters may pass one)
    if (job.hyp[0] == 0.0 || job.hyp[1] == 0.0 || !(fabs(job.hyp[0]) <= DBL_MAX) || !(fabs(job.hyp[1]) <= DBL_MAX)) {
        set_error("nll: length scales must be finite and non-zero"); return ST_BADARG;
    }
    const bool need_inv = job.ngrad > 0 || job.d_kinv != nullptr;

    SGP_TRY(c.Kmat.reserve((size_t)n_pad * n_pad * sizeof(double)));
    SGP_TRY(c.Dinv.reserve((size_t)nt * TILE * TILE * sizeof(double)));
    SGP_TRY(c.vecs.reserve((size_t)4 * n_pad * sizeof(double)));
    SGP_TRY(c.pts.reserve((size_t)N * sizeof(Pt)));
    const long npart = need_inv ? (dof2 ? grad4_num_partials(N) : grad_num_partials(N)) : 0;
    SGP_TRY(c.partial.reserve((size_t)(npart * 6 + 8) * sizeof(double)));
    SGP_TRY(c.small.reserve((size_t)(nt + 8) * sizeof(double)));
    if (need_inv) {
        SGP_TRY(c.Wmat.reserve((size_t)n_pad * n_pad * sizeof(double)));
        SGP_TRY(c.Tmat.reserve((trtri_workspace_doubles(n_pad) + 2) * sizeof(double)));
    }
    double* K = c.Kmat.as<double>();
    double* Dinv = c.Dinv.as<double>();
    double* yv = c.vecs.as<double>();
    double* wv = yv + n_pad;
    double* av = wv + n_pad;
    Pt* pts = c.pts.as<Pt>();
    double* logparts = c.small.as<double>();
    int* info = (int*)(logparts + nt);
    cudaStream_t st = c.stream;

    const HypC h = make_hypc(fam, job.hyp[0], job.hyp[1], job.hyp[2], job.per);
    const double noise = fabs(job.hyp[3]);

    c.pev_valid = false;
    SGP_TRY(c.mark(0));
    if (dof2) {
        // 2-DOF 4x4-block kernel (dof2.cu; not in the reference): coordinates are used as they are
        SGP_TRY(fill4_sym(c, job.d_x, N, job.hyp[0], job.hyp[1], job.hyp[2], noise, K, n_pad, n_pad));
    } else {
        SGP_TRY(make_points(c, fam, job.per, job.d_x, job.d_x + N, N, pts));
        if (job.reg) SGP_TRY(fill_reg_sym(c, fam, pts, N, h, noise, K, n_pad, n_pad, job.reg - 1));
        else SGP_TRY(fill_hess_sym(c, fam, pts, N, h, noise, K, n_pad, n_pad));
    }

    SGP_CUDA(cudaMemsetAsync(yv, 0, (size_t)n_pad * sizeof(double), st));
    SGP_CUDA(cudaMemcpyAsync(yv, job.d_z, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    SGP_CUDA(cudaMemsetAsync(info, 0, sizeof(double), st));
    SGP_TRY(c.mark(1));

    // opt-in (sgp_set_ozaki_ex): factor and inverse factor in one recursion whose products run on the INT8 tensor pipe
    // (ozaki_chol.cu).  It leaves X = L^-1 and never L, so it serves the evaluations that form the inverse and do not hand L out.
    const bool oz_ok = c.ozaki_slices > 0;
    const bool oz_fact = oz_ok && (c.ozaki_stages & 2) && need_inv && !job.d_L && n_pad > c.ozaki_leaf;
    // ... and its factor-only variant for the evaluations that need the value alone (nll_chol, the objective of the scripts'
    // L-BFGS / CMA loops): 0.38 n^3 flop on the INT8 pipe instead of n^3 / 3 on DMMA
    // (from 3 leaf sizes up: below that the two leaf factorisations + one leaf inversion cost more than one potrf_ll of the whole)
    const bool oz_value = oz_ok && (c.ozaki_stages & 2) && !need_inv && !job.d_L && !job.d_alpha && job.ngrad == 0 && n_pad >= 3 * c.ozaki_leaf;
    if (oz_value) {
        const size_t wb = ozaki_factinv_workspace_bytes(n_pad, c.ozaki_slices);
        SGP_TRY(c.ozbuf.reserve(wb));
        SGP_TRY(c.Tmat.reserve((trtri_workspace_doubles(n_pad / 2 + TILE) + ozaki_factor_solve_scratch_doubles(n_pad) + 2) * sizeof(double)));
        double* T = c.Tmat.as<double>();
        SGP_TRY(ozaki_factor_solve(c, c.ozaki_slices, c.ozaki_leaf, K, n_pad, n_pad, Dinv, logparts, info, T, yv, wv,
                                   T + trtri_workspace_doubles(n_pad / 2 + TILE), c.ozbuf.p, wb));
    } else if (!oz_fact) {
        // factor + forward substitution L w = z in one kernel (potrf_ll.cu).  0.5 z'alpha = 0.5 w'w, so the value
        // needs nothing else; alpha itself comes from the backward substitution, or -- when the inverse is formed
        // anyway -- from one transposed matrix-vector product with the explicit inverse factor.
        SGP_TRY(potrf(c, K, n_pad, n_pad, Dinv, logparts, info, yv, wv));
    }
    SGP_TRY(c.mark(2));
    if (job.d_L) {
        tril_out_kernel<<<1024, 256, 0, st>>>(K, n_pad, job.d_L, n);
        SGP_CUDA(cudaGetLastError());
        count_launch();
    }
    const bool need_alpha = job.d_alpha != nullptr || job.ngrad > 0;
    const double* dot_a = wv;                      // finalize: 0.5 * dot_a . dot_b
    const double* dot_b = wv;
    if (need_alpha && !need_inv) {
        SGP_CUDA(cudaMemcpyAsync(yv, wv, (size_t)n_pad * sizeof(double), cudaMemcpyDeviceToDevice, st));   // trsv_bwd eats its input
        SGP_TRY(trsv_bwd(c, K, n_pad, n_pad, Dinv, yv, av));
    }
    SGP_TRY(c.mark(3));

    double* partial = c.partial.as<double>();
    if (need_inv) {
        double* W = c.Wmat.as<double>();
        if (oz_fact) {
            const size_t wb = ozaki_factinv_workspace_bytes(n_pad, c.ozaki_slices);
            SGP_TRY(c.ozbuf.reserve(wb > trmv_lower_scratch_doubles(n_pad) * sizeof(double) ? wb : trmv_lower_scratch_doubles(n_pad) * sizeof(double)));
            SGP_TRY(ozaki_factinv(c, c.ozaki_slices, c.ozaki_leaf, K, n_pad, n_pad, Dinv, logparts, info, c.Tmat.as<double>(), c.ozbuf.p, wb));
            SGP_TRY(trmv_lower(c, K, n_pad, n_pad, yv, wv, c.ozbuf.as<double>()));          // w = L^-1 z
        } else {
            SGP_TRY(trtri(c, K, n_pad, n_pad, Dinv, c.Tmat.as<double>()));
        }
        // K now holds X = L^-1: alpha = L^-T w = X^T w is one pass over its columns
        if (need_alpha) SGP_TRY(gemv_t_lower(c, K, n_pad, n_pad, wv, av));
        SGP_TRY(c.mark(4));
        if (oz_ok && (c.ozaki_stages & 1)) {
            // opt-in: W = X^T X from INT8 slice products on the 5th-generation tensor core (sgp_set_ozaki)
            const size_t wb = ozaki_lauum_workspace_bytes(n_pad, c.ozaki_slices);
            SGP_TRY(c.ozbuf.reserve(wb));
            SGP_TRY(ozaki_lauum(c, c.ozaki_slices, K, n_pad, n_pad, W, n_pad, c.ozbuf.p, wb));
        } else {
            SGP_TRY(lauum(c, K, n_pad, n_pad, W, n_pad));
        }
        SGP_TRY(c.mark(5));
        if (job.ngrad > 0) {
            if (dof2) SGP_TRY(grad4_contract(c, job.d_x, N, job.hyp[0], job.hyp[1], job.hyp[2], W, n_pad, av, partial));
            else SGP_TRY(grad_contract(c, fam, job.reg, pts, N, h, W, n_pad, av, partial));
        }
        if (job.d_kinv) {
            sym_out_kernel<<<1024, 256, 0, st>>>(W, n_pad, job.d_kinv, n);
            SGP_CUDA(cudaGetLastError());
            count_launch();
        }
    } else {
        SGP_TRY(c.mark(4));
        SGP_TRY(c.mark(5));
    }
    if (job.d_alpha) SGP_CUDA(cudaMemcpyAsync(job.d_alpha, av, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    SGP_TRY(c.mark(6));
    finalize_kernel<<<1, 256, 0, st>>>(dot_a, dot_b, n, logparts, nt, info, partial, job.ngrad > 0 ? npart : 0, job.hyp[2],
                                       job.ngrad, job.d_res);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    SGP_TRY(c.mark(7));
    c.pev_valid = c.prof;
    return ST_OK;
}

}  // namespace sgp
