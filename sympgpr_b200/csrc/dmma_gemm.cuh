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
// 8x4 DMMA tiles per k4 step (32 DMMA per 12 64-bit shared loads).  Shared rows are padded so
// the 64-bit fragment loads are bank-conflict free for both layouts.  This header holds the
// contract (layouts, tile modes, GemmArgs, fragment addressing); the kernel is the warp-specialised
// persistent one of dmma_gemm_ws.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgp {

constexpr int GT = 128;        // CTA tile (M and N)
constexpr int GK = 16;         // k per pipeline stage
constexpr int LDMN = GT + 4;   // [k][m] rows of 132 doubles: (4*t + g) mod 16 distinct
constexpr int LDKK = GK + 4;   // [m][k] rows of 20 doubles:  (4*g + t) mod 16 distinct
constexpr int STAGE_DOUBLES = (GT * LDKK > GK * LDMN) ? GT * LDKK : GK * LDMN;   // 2560

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

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int LAY>
__device__ __forceinline__ double frag(const double* sm, int row, int k)
{
    return LAY == LAYOUT_MN ? sm[k * LDMN + row] : sm[row * LDKK + k];
}

inline long gemm_num_tiles(const GemmArgs& a)
{
    if (a.mode == TM_LOWER || a.mode == TM_LOWER_KGE) return (long)a.Mt * (a.Mt + 1) / 2;
    return (long)a.Mt * a.Nt;
}

}  // namespace sgp
