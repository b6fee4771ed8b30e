// Structured FP64 tensor-core GEMM for the blocked Cholesky / inverse.
//
//   C(m,n) = beta*C(m,n) + alpha * sum_{k in [k0,k1)} A(m,k) * Bt(n,k)
//
// C is column-major (ldc).  Each operand is given in one of two layouts:
//   LAYOUT_MN : element (m,k) at ptr[m + k*ld]   (contiguous along the tile's 128-dimension)
//   LAYOUT_K  : element (m,k) at ptr[k + m*ld]   (contiguous along k)
// so  L*L^T (syrk/trailing update), B*inv(L)^T (trsm by block inverse), L*X and X*T
// (triangular inverse) and X^T*X (lauum) are all the same kernel.
//
// The k-range of a 128x128 output tile depends on the tile index (triangular operands)
// and the tile set can be restricted to the lower triangle; see TileMode.
//
// Hardware mapping (sm_100a): FP64 has no tcgen05 kind; the FP64 tensor path is the
// warp-synchronous DMMA (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4; the m16n8k* PTX shapes
// decompose into it).  CTA tile 128x128x16, 8 warps as 2(M) x 4(N), warp tile 64x32 =
// 8x4 DMMA tiles per k4 step (32 DMMA per 12 64-bit shared loads).  Operands are staged
// global->shared with 16-byte cp.async in a STAGES-deep ring; shared rows are padded so
// the 64-bit fragment loads are bank-conflict free for both layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgp {

constexpr int GT = 128;        // CTA tile (M and N)
constexpr int GK = 16;         // k per pipeline stage
constexpr int GSTAGES = 4;
constexpr int GTHREADS = 256;
constexpr int LDMN = GT + 4;   // [k][m] rows of 132 doubles: (4*t + g) mod 16 distinct
constexpr int LDKK = GK + 4;   // [m][k] rows of 20 doubles:  (4*g + t) mod 16 distinct
constexpr int STAGE_DOUBLES = (GT * LDKK > GK * LDMN) ? GT * LDKK : GK * LDMN;   // 2560
constexpr size_t GEMM_SMEM = (size_t)GSTAGES * 2 * STAGE_DOUBLES * sizeof(double);  // 163840 B

enum Layout : int { LAYOUT_MN = 0, LAYOUT_K = 1 };

enum TileMode : int {
    TM_FULL = 0,      // all Mt x Nt tiles, k in [0,K)
    TM_LOWER = 1,     // tiles with tm >= tn only (square), k in [0,K)              (syrk)
    TM_LOWER_KGE = 2, // tiles with tm >= tn only, k in [tm*128, K)                  (X^T X, X lower)
    TM_B_LOWER = 3,   // all tiles, k in [tn*128, K)           (A * X, X lower triangular, K == N)
    TM_A_LOWER = 4,   // all tiles, k in [0, (tm+1)*128)       (X * T, X lower triangular, K == M)
};

struct GemmArgs {
    const double* A; long lda;
    const double* B; long ldb;
    double* C; long ldc;
    int Mt, Nt;       // tiles
    int K;            // multiple of GK
    double alpha, beta;
    int mode;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Load one 128 x GK operand tile (rows r0.., k from kk..) into a stage buffer.
template <int LAY>
__device__ __forceinline__ void load_tile(double* sm, const double* __restrict__ g, long ld, long r0, long kk, int tid)
{
    if (LAY == LAYOUT_MN) {
        // 16 k-rows of 128 contiguous doubles = 64 chunks of 16 B each -> 1024 chunks
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int ch = tid + i * GTHREADS;
            int k = ch >> 6, c = ch & 63;
            cp_async16(sm + k * LDMN + c * 2, g + (kk + k) * ld + r0 + c * 2);
        }
    } else {
        // 128 m-rows of 16 contiguous doubles = 8 chunks each -> 1024 chunks
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int ch = tid + i * GTHREADS;
            int m = ch >> 3, c = ch & 7;
            cp_async16(sm + m * LDKK + c * 2, g + (r0 + m) * ld + kk + c * 2);
        }
    }
}

template <int LAY>
__device__ __forceinline__ double frag(const double* sm, int row, int k)
{
    return LAY == LAYOUT_MN ? sm[k * LDMN + row] : sm[row * LDKK + k];
}

template <int AL, int BL>
__global__ void __launch_bounds__(GTHREADS, 1) gemm_f64_kernel(GemmArgs p)
{
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64;    // warp row offset in the tile
    const int wn = (warp >> 1) * 32;   // warp col offset

    // ---- tile index and k-range from the mode ------------------------------------
    int tm, tn;
    long k0 = 0, k1 = p.K;
    const long b = blockIdx.x;
    if (p.mode == TM_LOWER || p.mode == TM_LOWER_KGE) {
        // row-major enumeration of the lower triangle: b = tm(tm+1)/2 + tn
        long r = (long)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
        while ((r + 1) * (r + 2) / 2 <= b) r++;
        while (r * (r + 1) / 2 > b) r--;
        tm = (int)r;
        tn = (int)(b - r * (r + 1) / 2);
        if (p.mode == TM_LOWER_KGE) k0 = (long)tm * GT;
    } else if (p.mode == TM_A_LOWER) {
        tm = p.Mt - 1 - (int)(b / p.Nt);   // longest k-range first
        tn = (int)(b % p.Nt);
        k1 = (long)(tm + 1) * GT;
        if (k1 > p.K) k1 = p.K;
    } else {
        tm = (int)(b % p.Mt);
        tn = (int)(b / p.Mt);
        if (p.mode == TM_B_LOWER) k0 = (long)tn * GT;
    }
    const long m0 = (long)tm * GT, n0 = (long)tn * GT;
    const int nk = (int)((k1 - k0) / GK);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // ---- prologue: fill STAGES-1 stages --------------------------------------------
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; s++) {
        if (s < nk) {
            double* sa = smem + (size_t)s * 2 * STAGE_DOUBLES;
            load_tile<AL>(sa, p.A, p.lda, m0, k0 + (long)s * GK, tid);
            load_tile<BL>(sa + STAGE_DOUBLES, p.B, p.ldb, n0, k0 + (long)s * GK, tid);
        }
        cp_async_commit();
    }

    for (int it = 0; it < nk; it++) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        // prefetch stage it + STAGES-1 into the buffer freed by iteration it-1
        {
            int nx = it + GSTAGES - 1;
            if (nx < nk) {
                double* sa = smem + (size_t)(nx % GSTAGES) * 2 * STAGE_DOUBLES;
                load_tile<AL>(sa, p.A, p.lda, m0, k0 + (long)nx * GK, tid);
                load_tile<BL>(sa + STAGE_DOUBLES, p.B, p.ldb, n0, k0 + (long)nx * GK, tid);
            }
            cp_async_commit();
        }
        const double* sa = smem + (size_t)(it % GSTAGES) * 2 * STAGE_DOUBLES;
        const double* sb = sa + STAGE_DOUBLES;
#pragma unroll
        for (int kk = 0; kk < GK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; i++) af[i] = frag<AL>(sa, wm + i * 8 + g, kk + t);
#pragma unroll
            for (int j = 0; j < 4; j++) bf[j] = frag<BL>(sb, wn + j * 8 + g, kk + t);
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ------------------------------------------------------------------
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long row = m0 + wm + i * 8 + g;
            const long col = n0 + wn + j * 8 + 2 * t;
            double* c0 = p.C + row + col * p.ldc;
            double* c1 = c0 + p.ldc;
            if (beta == 0.0) {
                *c0 = alpha * acc[i][j][0];
                *c1 = alpha * acc[i][j][1];
            } else {
                *c0 = beta * (*c0) + alpha * acc[i][j][0];
                *c1 = beta * (*c1) + alpha * acc[i][j][1];
            }
        }
    }
}

inline long gemm_num_tiles(const GemmArgs& a)
{
    if (a.mode == TM_LOWER || a.mode == TM_LOWER_KGE) return (long)a.Mt * (a.Mt + 1) / 2;
    return (long)a.Mt * a.Nt;
}

template <int AL, int BL>
inline cudaError_t gemm_launch_t(const GemmArgs& a, cudaStream_t st)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_f64_kernel<AL, BL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)GEMM_SMEM);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    long nt = gemm_num_tiles(a);
    if (nt <= 0 || a.K <= 0) return cudaSuccess;
    gemm_f64_kernel<AL, BL><<<(unsigned)nt, GTHREADS, GEMM_SMEM, st>>>(a);
    return cudaGetLastError();
}

inline cudaError_t gemm_launch(int al, int bl, const GemmArgs& a, cudaStream_t st)
{
    if (al == LAYOUT_MN && bl == LAYOUT_MN) return gemm_launch_t<LAYOUT_MN, LAYOUT_MN>(a, st);
    if (al == LAYOUT_MN && bl == LAYOUT_K) return gemm_launch_t<LAYOUT_MN, LAYOUT_K>(a, st);
    if (al == LAYOUT_K && bl == LAYOUT_K) return gemm_launch_t<LAYOUT_K, LAYOUT_K>(a, st);
    return gemm_launch_t<LAYOUT_K, LAYOUT_MN>(a, st);
}

}  // namespace sgp
