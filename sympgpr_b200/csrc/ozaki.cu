// FP64 GEMM from the INT8 tensor pipe of sm_100a (Ozaki splitting): an OPT-IN building block, not on any default path.
//
// north_star asks for DMMA in the Cholesky (done: potrf_ll.cu / dmma_gemm_ws.cuh run at 91-97 % of the DMMA issue limit),
// which makes that limit -- 37 TFLOP/s -- the ceiling of the NLL+gradient evaluation.  The only way past it on Blackwell is
// the 5th-generation tensor core, which has no f64 kind: tcgen05.mma kind::i8 (INT8 x INT8 -> INT32, accumulators in tensor
// memory).  An FP64 product is recovered EXACTLY up to the slices kept by splitting both operands into signed 8-bit digits
// of their row-scaled mantissas (Ozaki scheme, balanced radix 256):
//     A(m,:) = 2^(ea(m)-7) sum_s As(m,:) 2^(-8 s),   B(n,:) = 2^(eb(n)-7) sum_t Bt(n,:) 2^(-8 t),   -128 <= As, Bt <= 127
//     C(m,n) = 2^(ea(m)+eb(n)-14) sum_d 2^(-8 d) sum_{s+t=d} As(m,:) . Bt(n,:)          (d < NS: the pairs that matter)
// Every slice product is an integer GEMM without rounding (NS K 128^2 < 2^31 per launch: deeper products are split along K
// on the host side), pairs with equal s + t share one INT32 accumulator in tensor memory, and the only floating-point
// roundings are the final NS-term sum per output element.  Balanced digits (every digit signed, taken from the bottom of the
// two's-complement mantissa with a carry) keep the dropped pairs zero-mean, so their sum over k grows like sqrt(K), not K.
// NS digits keep 8 NS - 1 bits of every ROW'S scale: 6 -> 47 (21 slice pairs), 7 -> 55 (28 pairs: FP64-grade), 8 -> 63.
// (Round-2 states up to "i" used 7-bit digits from round-to-nearest of every stage: 49 bits for 28 pairs.)
//
// This file: (1) the tcgen05 plumbing (TMEM allocation, shared-memory matrix descriptors for the K-major 128-byte-swizzled
// canonical layout, the kind::i8 instruction descriptor, tcgen05.commit -> mbarrier, tcgen05.ld) with a self test of one
// 128 x 64 x K INT8 product against a plain integer kernel; (2) the slicing kernel; (3) the Ozaki GEMM
// C = alpha A B^T + beta C on FP64 operands, validated against gemm_f64_ws_kernel (tests/test_gpu_ozaki.py).
#include "ozaki.cuh"

#include <cuda.h>
#include <float.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "mbar.cuh"

namespace sgp {

namespace {

// ---------------------------------------------------------------------------------------------------------
// tcgen05 helpers (PTX as CUTLASS 4.x spells it: cute/arch/mma_sm100_umma.hpp, tmem_allocator_sm100.hpp, copy_sm100.hpp)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] B[smem]^T, INT8 x INT8 -> INT32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
// the same with the A operand kept in the tensor core's collector buffer for the next MMA (KEEP = 1: fill and keep) or taken
// from it (KEEP = 2: last use): two MMAs that share A read it from shared memory once (SASS UTCIMMA .A_KEEP / .A_REUSE)
template <int KEEP>
__device__ __forceinline__ void umma_i8_a(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    if (KEEP == 1)
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
            : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
            : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
__device__ __forceinline__ void umma_commit(unsigned long long* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread t of warp w (w = warp index % 4) receives columns c .. c+31 of TMEM lane 32 w + t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp, SmemDescriptor) of a K-major operand tile in the canonical
// 128-byte-swizzled layout: rows of 128 bytes (128 INT8 along K), 8-row atoms of 1024 bytes, the 16-byte chunk index of a
// row XORed with the row index inside its atom (Swizzle<3,4,3>), tile base 1024-byte aligned.
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (1: unused for swizzled K-major)
//   bits [32,46) stride byte offset >> 4 (1024 B between 8-row atoms)     bits [46,48) version = 1 (Blackwell)
//   bits [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(const void* smem_tile)
{
    const uint64_t addr = (uint64_t)((smem_u32(smem_tile) & 0x3FFFFu) >> 4);
    return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// Instruction descriptor (InstrDescriptor) for kind::i8: dense, no saturation, C = S32 (2 at bits [4,6)), A and B signed
// 8 bit (1 at bits [7,10) and [10,13)), both K-major (0 at bits 15 and 16), N >> 3 at bits [17,23), M >> 4 at bits [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of (row r, byte k < 128) inside a 128-byte-swizzled K-major tile
__host__ __device__ inline uint32_t sw128_offset(int r, int k)
{
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 4) ^ (r & 7)) & 7) << 4) + (k & 15));
}

constexpr int OZ_KC_ = 128;                            // k-chunk (bytes of one swizzled row)
constexpr int ST_M = 128, ST_N = 64, ST_KC = 128;     // self test: one 128 x 64 tile, k-chunks of 128 bytes

// ---------------------------------------------------------------------------------------------------------
// (1) self test: D (128 x 64, INT32) = A (128 x K, INT8, K contiguous) B (64 x K)^T on the INT8 tensor pipe.  One CTA of four
// warps; every k-chunk is copied to shared memory by all threads in the swizzled layout, four UMMAs (K = 32 each) are issued
// by thread 0 and committed to an mbarrier that everybody waits on before the chunk is overwritten: the simplest correct
// pipeline (no overlap), only there to pin descriptors, TMEM addressing and the tcgen05.ld mapping against a plain kernel.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) i8mma_selftest_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int K,
                                                                 int32_t* __restrict__ D)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sA = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = sA + ST_M * ST_KC;
    __shared__ unsigned long long bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = umma_idesc_i8(ST_M, ST_N);
    uint32_t phase = 0;
    for (int k0 = 0; k0 < K; k0 += ST_KC) {
        for (int c = tid; c < ST_M * 8; c += 128) {
            const int r = c >> 3, ch = c & 7;
            *reinterpret_cast<uint4*>(sA + sw128_offset(r, ch * 16)) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + k0 + ch * 16);
        }
        for (int c = tid; c < ST_N * 8; c += 128) {
            const int r = c >> 3, ch = c & 7;
            *reinterpret_cast<uint4*>(sB + sw128_offset(r, ch * 16)) = *reinterpret_cast<const uint4*>(B + (size_t)r * K + k0 + ch * 16);
        }
        fence_async_smem();                      // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < ST_KC / 32; ks++)
                umma_i8(tmem, umma_desc_k128(sA + ks * 32), umma_desc_k128(sB + ks * 32), idesc, (k0 > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        tc_fence_after();
    }
    uint32_t v[32];
#pragma unroll
    for (int half = 0; half < 2; half++) {
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + half * 32, v);
#pragma unroll
        for (int i = 0; i < 32; i++) D[(size_t)(32 * warp + lane) * ST_N + half * 32 + i] = (int32_t)v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

__global__ void i8_ref_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int M, int N, int K, int32_t* __restrict__ D)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const int m = idx / N, n = idx % N;
    int s = 0;
    for (int k = 0; k < K; k++) s += (int)A[(size_t)m * K + k] * (int)B[(size_t)n * K + k];
    D[idx] = s;
}

__global__ void i8_fill_kernel(int8_t* __restrict__ a, long n, unsigned long long seed)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        a[i] = (int8_t)((int)((z >> 40) % 129ull) - 64);        // [-64, 64]: the range of the Ozaki slices
    }
}

__global__ void i32_diff_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int n, int* __restrict__ nbad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) atomicAdd(nbad, 1);
}

// ---------------------------------------------------------------------------------------------------------
// (2) slicing: FP64 operand X (R rows x K) -> row exponents e(r) with max_k |X(r,k)| < 2^e(r) and NS INT8 slices
//     slices[s][r][k] (k contiguous, row pitch Kp, slice pitch Rp Kp) with
//     X(r,k) = 2^(e-7) sum_s slices[s][r][k] 2^(-8 s) + O(2^(e - 8 NS)),  -128 <= slice <= 127.
// Every step is exact (scaling by a power of two, one rounding to a 64-bit integer, integer digit extraction).  Two storage orders of X:
//     OZ_MN: element (r, k) at X[r + k ld]   (operand rows run down the columns of a column-major matrix)
//     OZ_K:  element (r, k) at X[k + r ld]   (operand rows ARE the columns of a column-major matrix: the operand is X^T)
// and triangular operands whose other half holds unrelated numbers (the Cholesky works in the lower triangle only):
//     tri = 1: only k <= r is valid (OZ_MN view of a lower-triangular matrix),  tri = 2: only k >= r (OZ_K view of one);
// invalid entries become zeros, and k-chunks that no GEMM tile of a triangular product loads are not written at all.
// Rows / columns beyond R / K are zero.
// ---------------------------------------------------------------------------------------------------------
// 2^e as a double for -1022 <= e <= 1023
__device__ __forceinline__ double oz_pow2(int e) { return __hiloint2double((1023 + e) << 20, 0); }

// x (|x| < 0.995 * 2^e, e its row's exponent; anything else -> treated as 0) -> NS balanced radix-256 digits through
// `put(s, digit)`: v = rint(x 2^(8 NS - 1 - e)) as a 64-bit integer (exact: a power-of-two scaling of a 53-bit mantissa, at most
// 2 bits of left shift), digits from the bottom with the carry of the sign extension
template <int NS, class Put>
__device__ __forceinline__ void oz_split(double x, int e, Put put)
{
    const int k = 8 * NS - 1 - e, k1 = k > 1000 ? 1000 : k;       // -969 <= k <= 1063
    const double y = (x * oz_pow2(k1)) * oz_pow2(k - k1);
    long long v = (fabs(y) < 0.996 * oz_pow2(8 * NS - 1)) ? __double2ll_rn(y) : 0ll;       // NaN / Inf / out of range: not representable
#pragma unroll
    for (int s = NS - 1; s >= 0; s--) {
        const int8_t d = (int8_t)(v & 0xff);
        put(s, d);
        v = (v - d) >> 8;
    }
}

// row exponent: max|x| = f 2^e with 0.5 <= f < 0.995 (one more when the mantissa is at the top of its binade, so that the leading
// digit stays below 128); rows below 1e-290 count as zero rows
__device__ __forceinline__ int oz_exp_of(double mx)
{
    int e = 0;
    if (mx >= 1e-290 && mx <= DBL_MAX) {
        const double f = frexp(mx, &e);
        if (f >= 0.995) e++;
    }
    return e;
}

constexpr int OZ_RMK = 512;                            // columns per block of the row-maximum pass
__global__ void __launch_bounds__(128) oz_rowmax_mn_kernel(const double* __restrict__ X, long ld, long R, long K, int tri,
                                                           unsigned long long* __restrict__ rowmax)
{
    const long r = (long)blockIdx.x * 128 + threadIdx.x;
    const long k0 = (long)blockIdx.y * OZ_RMK;
    if (r >= R) return;
    long k1 = k0 + OZ_RMK < K ? k0 + OZ_RMK : K;
    if (tri == 1 && k1 > r + 1) k1 = r + 1;
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
    long k = k0;
    for (; k + 3 < k1; k += 4) {
        m0 = fmax(m0, fabs(X[r + k * ld]));
        m1 = fmax(m1, fabs(X[r + (k + 1) * ld]));
        m2 = fmax(m2, fabs(X[r + (k + 2) * ld]));
        m3 = fmax(m3, fabs(X[r + (k + 3) * ld]));
    }
    for (; k < k1; k++) m0 = fmax(m0, fabs(X[r + k * ld]));
    const double mx = fmax(fmax(m0, m1), fmax(m2, m3));
    if (mx > 0.0) atomicMax(rowmax + r, (unsigned long long)__double_as_longlong(mx));     // non-negative doubles order like integers
}

__global__ void oz_exp_kernel(const unsigned long long* __restrict__ rowmax, long R, long Rp, int* __restrict__ exps)
{
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Rp) return;
    exps[r] = (r < R) ? oz_exp_of(__longlong_as_double((long long)rowmax[r])) : 0;
}

// OZ_MN slicing: block = 32 rows x 128 k.  Phase 1: thread (row tx, k = ty + 8 j) loads coalesced along the rows, splits, and
// stores the bytes into a [slice][row][k] tile (pitch 132: conflict-free both ways); phase 2: 8 threads per row write 128
// contiguous bytes per slice with 16-byte stores.
template <int NS>
__global__ void __launch_bounds__(256) oz_slice_mn_kernel(const double* __restrict__ X, long ld, long R, long K, int tri,
                                                           const int* __restrict__ exps, int8_t* __restrict__ slices, long Rp, long Kp)
{
    constexpr int PITCH = 132;
    __shared__ __align__(16) int8_t tile[NS][32][PITCH];
    const long r0 = (long)blockIdx.x * 32, k0 = (long)blockIdx.y * 128;
    if (tri == 1 && k0 >= (r0 / 128 + 1) * 128) return;             // beyond every k-range that reads these rows
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long r = r0 + tx;
    const int e = (r < R) ? exps[r] : 0;
    double x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const long k = k0 + ty + 8 * j;
        x[j] = (r < R && k < K && (tri != 1 || k <= r)) ? X[r + k * ld] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 16; j++) oz_split<NS>(x[j], e, [&](int s, int8_t b) { tile[s][tx][ty + 8 * j] = b; });
    __syncthreads();
    const int row = threadIdx.x >> 3, kq = threadIdx.x & 7;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&tile[s][row][kq * 16]);
        const uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
        *reinterpret_cast<uint4*>(slices + (size_t)s * Rp * Kp + (size_t)(r0 + row) * Kp + k0 + kq * 16) = v;
    }
}

// OZ_K input, one warp per operand row (a contiguous column of the matrix)
__global__ void __launch_bounds__(256) oz_rowexp_k_kernel(const double* __restrict__ X, long ld, long R, long Rp, long K, int tri, int* __restrict__ exps)
{
    const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= Rp) return;
    double mx = 0.0;
    if (r < R) {
        const double* row = X + r * ld;
        for (long k = (tri == 2 ? r : 0) + lane; k < K; k += 32) mx = fmax(mx, fabs(row[k]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) exps[r] = oz_exp_of(mx);
}

// OZ_K slicing: one thread = 16 consecutive k of one row: 128 bytes in, NS x 16 bytes out
template <int NS>
__global__ void __launch_bounds__(256) oz_slice_k_kernel(const double* __restrict__ X, long ld, long R, long K, int tri,
                                                          const int* __restrict__ exps, int8_t* __restrict__ slices, long Rp, long Kp)
{
    const long kb = (long)blockIdx.y * blockDim.x + threadIdx.x;       // 16-element block index along k
    const long r = blockIdx.x;                                          // (rows in x: more than 65 535 of them at n = 65 536)
    if (kb * 16 >= Kp) return;
    const long k0 = kb * 16;
    if (tri == 2 && k0 + 16 <= (r / 128) * 128) return;                 // never loaded
    const int e = (r < R) ? exps[r] : 0;
    const double* row = X + r * ld;
    double x[16];
    if (r < R && k0 + 16 <= K && (tri != 2 || k0 >= r)) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const double2 v = *reinterpret_cast<const double2*>(row + k0 + i);          // ld even, k0 % 16 == 0
            x[i] = v.x; x[i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const long k = k0 + i;
            x[i] = (r < R && k < K && (tri != 2 || k >= r)) ? row[k] : 0.0;
        }
    }
    union { int8_t b[NS][16]; uint4 v[NS]; } out;
#pragma unroll
    for (int i = 0; i < 16; i++) oz_split<NS>(x[i], e, [&](int s, int8_t b) { out.b[s][i] = b; });
#pragma unroll
    for (int s = 0; s < NS; s++) *reinterpret_cast<uint4*>(slices + (size_t)s * Rp * Kp + (size_t)r * Kp + k0) = out.v[s];
}

// ---------------------------------------------------------------------------------------------------------
// (3) the GEMM  C = alpha A B^T + beta C  on sliced operands.
// One CTA (four warps) per 128 x 64 output tile, persistent, tiles handed out in list order by an atomic ticket (the k-ranges
// of triangular products differ per tile; the lists put the longest first).  Warp-specialised main loop on a two-stage ring:
//   warp 0, one lane: TMA producer -- per k-chunk of 64 it waits for the stage to be free and issues 2 NS bulk tensor copies
//                     (cp.async.bulk.tensor.3d, 64-byte swizzle: one box per slice of A and of B) onto the stage's full barrier;
//   warp 1, one lane: MMA issuer -- waits for the bytes and issues tcgen05.mma kind::i8 (M = 128, K = 32).  The slice pairs
//                     (s, t) with s + t = d accumulate into TMEM columns [64 d, 64 d + 64), so for a fixed s the pairs
//                     t = 0 .. NS-1-s are ONE product of A_s with the row-stacked slices [B_0; B_1; ...] (contiguous in shared
//                     memory) into the contiguous columns [64 s, 64 NS): issued as MMAs of N <= 256 -- 10 instead of 28
//                     per k-step at NS = 7, and 4 KB of A read from shared memory per 256 instead of per 64 columns (a
//                     128 x 64 x 32 MMA reads 6 KB per 32 clocks of tensor time and is bound by shared-memory bandwidth);
//                     then tcgen05.commit to the stage's empty barrier (after the last chunk: to the accumulator barrier);
//   all four warps:   epilogue -- tcgen05.ld of their 32 TMEM lanes, the NS integer accumulators combined in FP64 from the
//                     smallest term up, 2^(ea + eb - 12), C written.
// 12 KB per slice and stage: 192 KB of shared memory at NS = 8; all 512 TMEM columns at NS = 8.
// ---------------------------------------------------------------------------------------------------------
constexpr int OZ_M = 128, OZ_N = 64, OZ_KC = 64, OZ_STAGES = 2;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, unsigned long long* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// K-major operand tile with 64-byte rows in the canonical 64-byte-swizzled layout (8-row atoms of 512 bytes): layout type 4,
// stride byte offset 512; written by the tensor-map copies with CU_TENSOR_MAP_SWIZZLE_64B, tile base 512-byte aligned
__device__ __forceinline__ uint64_t umma_desc_k64(const void* smem_tile)
{
    const uint64_t addr = (uint64_t)((smem_u32(smem_tile) & 0x3FFFFu) >> 4);
    return addr | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

struct OzArgs {
    const int* ea; const int* eb;       // row exponents of A (M) and B (N)
    double* C; long ldc;                // column-major output
    long M, N;                          // logical sizes (rows beyond them exist, zero, inside the slices)
    int nk;                             // k-chunks of OZ_KC in the sliced operands
    int kc_lo, kc_hi;                   // this launch covers the chunks [kc_lo, kc_hi) (deep products are split along K)
    int kmode;                          // OZ_KLO_* / OZ_KHI_* bits: the k-range of a tile
    const int2* tiles; int ntiles;      // (tm, tn) of every tile in processing order (oz_tile_list)
    int* ticket;                        // zeroed before the launch
    double alpha, beta;
};

constexpr int OZ_THREADS = 192;     // warps 0-3: epilogue (TMEM lanes 32 w ..), warp 4: TMA producer, warp 5: MMA issuer

template <int NS>
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, OzArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sbase = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int A_BYTES = OZ_M * OZ_KC, B_BYTES = OZ_N * OZ_KC, STAGE_BYTES = NS * (A_BYTES + B_BYTES);
    // ring of operand slices: 2 k-chunks x NS slots for the A slices (8 KB each, handed back one by one as soon as the MMAs of
    // their s are done, so the refill of a slot has ~2 chunk times to land), 2 buffers for the NS B slices of a chunk
    __shared__ unsigned long long fullA[OZ_STAGES * NS], emptyA[OZ_STAGES * NS], fullB[OZ_STAGES], emptyB[OZ_STAGES];
    __shared__ unsigned long long tq_full[2], tq_empty[2], tmem_full, tmem_empty;
    __shared__ int tileq[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_eb[OZ_N];
    constexpr uint32_t TCOLS = (NS * OZ_N <= 64) ? 64 : (NS * OZ_N <= 128) ? 128 : (NS * OZ_N <= 256) ? 256 : 512;
    static_assert(NS * OZ_N <= 512, "accumulators of all slice-sum classes must fit the 512 TMEM columns");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
    if (tid == 32) {
        for (int i = 0; i < OZ_STAGES * NS; i++) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < OZ_STAGES; i++) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tq_full[i], 1); mbar_init(&tq_empty[i], 129); }
        mbar_init(&tmem_full, 1);
        mbar_init(&tmem_empty, 128);
        mbar_fence_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // k-chunk range of a tile
    auto krange = [&](int tm, int tn, int& kc0, int& kc1) {
        kc0 = a.kc_lo; kc1 = a.kc_hi;
        if (a.kmode & OZ_KLO_TM) kc0 = max(kc0, tm * (OZ_M / OZ_KC));
        if (a.kmode & OZ_KLO_TN) kc0 = max(kc0, tn * (OZ_N / OZ_KC));
        if (a.kmode & OZ_KHI_TM) kc1 = min(kc1, (tm + 1) * (OZ_M / OZ_KC));
        if (a.kmode & OZ_KHI_TN) kc1 = min(kc1, (tn + 1) * (OZ_N / OZ_KC));
    };
    if (warp == 4) {
        // ---- TMA producer: takes the tickets, announces every tile to the other roles and runs ahead of them by the depth of
        // the ring (the first chunks of the next tile are in flight while the epilogue of the current one runs)
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t nt = 0;; nt++) {
                const int tile = atomicAdd(a.ticket, 1);
                const uint32_t q = nt & 1u;
                mbar_wait(&tq_empty[q], ((nt >> 1) & 1u) ^ 1u);
                tileq[q] = tile;
                mbar_arrive(&tq_full[q]);
                if (tile >= a.ntiles) break;
                const int2 tt = a.tiles[tile];
                const int tm = tt.x, tn = tt.y;
                int kc0, kc1;
                krange(tm, tn, kc0, kc1);
                for (int kc = kc0; kc < kc1; kc++, it++) {
                    const uint32_t st = it % OZ_STAGES, ph = (it / OZ_STAGES) & 1u;
                    uint8_t* sA = sbase + st * STAGE_BYTES;
                    uint8_t* sB = sA + NS * A_BYTES;
                    mbar_wait(&emptyB[st], ph ^ 1u);
                    mbar_arrive_expect_tx(&fullB[st], (uint32_t)(NS * B_BYTES));
#pragma unroll
                    for (int s = 0; s < NS; s++) tma_load_3d(sB + s * B_BYTES, &tmB, kc * OZ_KC, tn * OZ_N, s, &fullB[st]);
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        mbar_wait(&emptyA[st * NS + s], ph ^ 1u);
                        mbar_arrive_expect_tx(&fullA[st * NS + s], (uint32_t)A_BYTES);
                        tma_load_3d(sA + s * A_BYTES, &tmA, kc * OZ_KC, tm * OZ_M, s, &fullA[st * NS + s]);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ---- MMA issuer
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t nt = 0;; nt++) {
                const uint32_t q = nt & 1u;
                mbar_wait(&tq_full[q], (nt >> 1) & 1u);
                const int tile = tileq[q];
                mbar_arrive(&tq_empty[q]);
                if (tile >= a.ntiles) break;
                const int2 tt = a.tiles[tile];
                int kc0, kc1;
                krange(tt.x, tt.y, kc0, kc1);
                mbar_wait(&tmem_empty, (nt & 1u) ^ 1u);      // the epilogue of the previous tile has read the accumulators
                tc_fence_after();
                for (int kc = kc0; kc < kc1; kc++, it++) {
                    const uint32_t st = it % OZ_STAGES, ph = (it / OZ_STAGES) & 1u;
                    mbar_wait(&fullB[st], ph);
                    const uint8_t* sA = sbase + st * STAGE_BYTES;
                    const uint8_t* sB = sA + NS * A_BYTES;
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        // A_s x [B_0; ...; B_{NS-1-s}] -> accumulators s .. NS-1; the first write of every accumulator of this tile
                        // is chunk kc0, k-step 0 of s = 0
                        mbar_wait(&fullA[st * NS + s], ph);
                        tc_fence_after();
#pragma unroll
                        for (int ks = 0; ks < OZ_KC / 32; ks++) {
                            const uint64_t adesc = umma_desc_k64(sA + s * A_BYTES + ks * 32);
                            const uint32_t accf = (kc > kc0 || ks > 0 || s > 0) ? 1u : 0u;
                            const int ncols = (NS - s) * OZ_N;
                            if (ncols > 256) {
                                // two MMAs share A_s: the first keeps it in the collector buffer, the second takes it from there
                                umma_i8_a<1>(tmem + s * OZ_N, adesc, umma_desc_k64(sB + ks * 32), umma_idesc_i8(OZ_M, 256), accf);
                                umma_i8_a<2>(tmem + s * OZ_N + 256, adesc, umma_desc_k64(sB + 4 * B_BYTES + ks * 32),
                                             umma_idesc_i8(OZ_M, ncols - 256), accf);
                            } else {
                                umma_i8(tmem + s * OZ_N, adesc, umma_desc_k64(sB + ks * 32), umma_idesc_i8(OZ_M, ncols), accf);
                            }
                        }
                        umma_commit(&emptyA[st * NS + s]);   // this A slot is free once the MMAs issued so far have read it
                    }
                    umma_commit(&emptyB[st]);                // and the chunk's B slices
                }
                umma_commit(&tmem_full);                     // all accumulators of this tile final
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue warps: thread = row m of the tile (TMEM lane 32 warp + lane)
        for (uint32_t nt = 0;; nt++) {
            const uint32_t q = nt & 1u;
            mbar_wait(&tq_full[q], (nt >> 1) & 1u);
            const int tile = tileq[q];
            mbar_arrive(&tq_empty[q]);
            if (tile >= a.ntiles) break;
            const int2 tt = a.tiles[tile];
            const int tm = tt.x, tn = tt.y;
            int kc0, kc1;
            krange(tm, tn, kc0, kc1);
            const bool empty = kc0 >= kc1;
            if (tid < OZ_N) {
                const long n = (long)tn * OZ_N + tid;
                s_eb[tid] = (n < a.N) ? a.eb[n] : 0;
            }
            const long m = (long)tm * OZ_M + 32 * warp + lane;
            const int eam = ((m < a.M) ? a.ea[m] : 0) - 7;
            asm volatile("bar.sync 1, 128;" ::: "memory");           // s_eb visible to the four epilogue warps
            // the accumulators take the whole main loop of the tile (~10^5 clocks): wait with a back-off instead of polling
            while (!mbar_try_wait(&tmem_full, nt & 1u)) __nanosleep(256);
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < OZ_N / 32; half++) {
                double acc[32];
#pragma unroll
                for (int i = 0; i < 32; i++) acc[i] = 0.0;
                if (!empty) {
#pragma unroll 1
                    for (int d = NS - 1; d >= 0; d--) {           // smallest contributions first
                        uint32_t v[32];
                        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + d * OZ_N + half * 32, v);
                        const double w = oz_pow2(-8 * d);
#pragma unroll
                        for (int i = 0; i < 32; i++) acc[i] = fma((double)(int32_t)v[i], w, acc[i]);
                    }
                }
                if (half == OZ_N / 32 - 1) {                      // last TMEM read of this tile done: the next tile's MMAs may start
                    tc_fence_before();
                    mbar_arrive(&tmem_empty);
                }
                if (m < a.M && !(empty && a.beta == 1.0)) {       // (a later k-block of a tile whose range ended earlier: nothing to add)
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const long n = (long)tn * OZ_N + half * 32 + i;
                        if (n < a.N) {
                            // alpha acc 2^(ea + eb - 14) as two exact scalings, the smaller exponent first (no spurious overflow)
                            const int eb6 = s_eb[half * 32 + i] - 7;
                            const double val = a.alpha * ((acc[i] * oz_pow2(min(eam, eb6))) * oz_pow2(max(eam, eb6)));
                            double* cp = a.C + m + n * a.ldc;
                            *cp = (a.beta == 0.0) ? val : fma(a.beta, *cp, val);
                        }
                    }
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");           // everybody is done with s_eb
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// Processing order of the output tiles.  Operand bytes per tile and k-chunk are 2 NS x 12 KB against 2 NS (NS + 1) MMA units,
// i.e. ~340 INT8 ops per byte -- below the ~730 ops/byte the tensor pipe needs from HBM -- so the tiles that run at the same
// time must share their operand panels in L2: super-tiles of 8 (tm) x 16 (tn) = 128 tiles (about one wave of 148 CTAs) touch
// 8 A panels and 16 B panels instead of 64 + 2.  `order` puts the tiles with the longest k-range first (the ticket hands them
// out in list order): 0 = block rows top down, 1 = bottom up, 2 = block columns left to right, 3 = right to left.
struct OzTileList { int device = -1, mt = 0, nt = 0, lower = 0, order = 0; int n = 0; DBuf buf; };

int oz_tile_list(Ctx& c, int mt, int nt, int lower, int order, const int2** d_tiles, int* ntiles)
{
    constexpr int NLIST = 64;
    static OzTileList lists[NLIST];
    static int next = 0;
    static std::mutex mu;                                 // contexts of several host threads share the cache
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < NLIST; i++) {
        OzTileList& L = lists[i];
        if (L.device == c.device && L.mt == mt && L.nt == nt && L.lower == lower && L.order == order && L.buf.p) {
            *d_tiles = L.buf.as<int2>(); *ntiles = L.n;
            return ST_OK;
        }
    }
    std::vector<int2> v;
    constexpr int BM = 8, BN = 16;
    const int nTM = (mt + BM - 1) / BM, nTN = (nt + BN - 1) / BN;
    auto block = [&](int TM, int TN) {
        for (int j = 0; j < BN; j++)
            for (int i = 0; i < BM; i++) {
                const int tm = TM * BM + i, tn = TN * BN + j;
                if (tm >= mt || tn >= nt) continue;
                if (lower && tn > 2 * tm + 1) continue;
                v.push_back(make_int2(tm, tn));
            }
    };
    if (order == 0) { for (int TM = 0; TM < nTM; TM++) for (int TN = 0; TN < nTN; TN++) block(TM, TN); }
    else if (order == 1) { for (int TM = nTM - 1; TM >= 0; TM--) for (int TN = 0; TN < nTN; TN++) block(TM, TN); }
    else if (order == 2) { for (int TN = 0; TN < nTN; TN++) for (int TM = 0; TM < nTM; TM++) block(TM, TN); }
    else { for (int TN = nTN - 1; TN >= 0; TN--) for (int TM = 0; TM < nTM; TM++) block(TM, TN); }
    OzTileList& L = lists[next];
    next = (next + 1) % NLIST;
    // a list that is replaced may still be read by a kernel in flight on this or another device's stream
    if (L.buf.p) SGP_CUDA(cudaDeviceSynchronize());
    SGP_TRY(L.buf.reserve((v.size() + 1) * sizeof(int2)));
    SGP_CUDA(cudaMemcpyAsync(L.buf.p, v.data(), v.size() * sizeof(int2), cudaMemcpyHostToDevice, c.stream));
    SGP_CUDA(cudaStreamSynchronize(c.stream));
    L.device = c.device; L.mt = mt; L.nt = nt; L.lower = lower; L.order = order; L.n = (int)v.size();
    *d_tiles = L.buf.as<int2>(); *ntiles = L.n;
    return ST_OK;
}

int oz_ticket(Ctx& c, int** ticket)
{
    // one ticket word per (device, stream slot): launches on ONE stream are ordered, so they may share a word; contexts on
    // different streams of a device get different words (8 slots, hashed by the stream handle)
    static DBuf tickets[64];
    static std::mutex mu;
    if (c.device < 0 || c.device >= 64) { set_error("ozaki: device index out of range"); return ST_BADARG; }
    {
        std::lock_guard<std::mutex> lock(mu);
        SGP_TRY(tickets[c.device].reserve(8 * 256));
    }
    const size_t slot = ((uintptr_t)c.stream >> 4) % 8;
    *ticket = tickets[c.device].as<int>() + slot * 64;
    SGP_CUDA(cudaMemsetAsync(*ticket, 0, sizeof(int), c.stream));
    return ST_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_slice_tmap(CUtensorMap* tm, const int8_t* slices, long Rp, long Kp, int ns, int box_rows)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SGP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return ST_CUDA; }
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)Rp, (cuuint64_t)ns};
    const cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * (cuuint64_t)Rp};             // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)OZ_KC, (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)slices, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return ST_CUDA; }
    return ST_OK;
}

template <int NS>
int oz_launch(Ctx& c, const CUtensorMap& tmA, const CUtensorMap& tmB, const OzArgs& a)
{
    const size_t smem = (size_t)OZ_STAGES * NS * (OZ_M + OZ_N) * OZ_KC + 1024;
    static bool configured[64] = {};
    if (first_use_on_current_device(configured))
        SGP_CUDA(cudaFuncSetAttribute(oz_gemm_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = c.sm_count > 0 ? c.sm_count : 148;
    oz_gemm_kernel<NS><<<(unsigned)(a.ntiles < sms ? a.ntiles : sms), OZ_THREADS, smem, c.stream>>>(tmA, tmB, a);
    SGP_CUDA(cudaGetLastError());
    count_launch();
    return ST_OK;
}

template <int NS>
int oz_slice_t(Ctx& c, const double* X, long ld, long R, long K, int layout, int tri, const OzSliced& o)
{
    if (layout == OZ_MN) {
        SGP_CUDA(cudaMemsetAsync(o.rowmax, 0, (size_t)o.Rp * sizeof(unsigned long long), c.stream));
        oz_rowmax_mn_kernel<<<dim3((unsigned)((R + 127) / 128), (unsigned)((K + OZ_RMK - 1) / OZ_RMK)), 128, 0, c.stream>>>(X, ld, R, K, tri, o.rowmax);
        oz_exp_kernel<<<(unsigned)((o.Rp + 255) / 256), 256, 0, c.stream>>>(o.rowmax, R, o.Rp, o.ex);
        oz_slice_mn_kernel<NS><<<dim3((unsigned)(o.Rp / 32), (unsigned)(o.Kp / 128)), 256, 0, c.stream>>>(X, ld, R, K, tri, o.ex, o.sl, o.Rp, o.Kp);
        count_launch(3);
    } else {
        oz_rowexp_k_kernel<<<(unsigned)((o.Rp + 7) / 8), 256, 0, c.stream>>>(X, ld, R, o.Rp, K, tri, o.ex);
        oz_slice_k_kernel<NS><<<dim3((unsigned)o.Rp, (unsigned)((o.Kp / 16 + 255) / 256)), 256, 0, c.stream>>>(X, ld, R, K, tri, o.ex, o.sl, o.Rp, o.Kp);
        count_launch(2);
    }
    SGP_CUDA(cudaGetLastError());
    return ST_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// host interface
// ---------------------------------------------------------------------------------------------------------
size_t ozaki_sliced_bytes(long R, long K, int ns)
{
    const size_t Rp = (size_t)round_up(R, OZ_M), Kp = (size_t)round_up(K, OZ_KC_);
    return (size_t)ns * Rp * Kp + Rp * (sizeof(int) + sizeof(unsigned long long)) + 256;
}

OzSliced ozaki_carve(void* buf, long R, long K, int ns)
{
    OzSliced o;
    o.Rp = round_up(R, OZ_M); o.Kp = round_up(K, OZ_KC_);
    o.sl = reinterpret_cast<int8_t*>(((uintptr_t)buf + 255) & ~(uintptr_t)255);
    o.rowmax = reinterpret_cast<unsigned long long*>(o.sl + (size_t)ns * o.Rp * o.Kp);
    o.ex = reinterpret_cast<int*>(o.rowmax + o.Rp);
    return o;
}

static int oz_check_ns(int ns)
{
    if (ns < 4 || ns > 8) { set_error("ozaki: 4..8 slices per operand, got %d", ns); return ST_BADARG; }
    return ST_OK;
}

int ozaki_slice(Ctx& c, int ns, const double* X, long ld, long R, long K, int layout, int tri, const OzSliced& out)
{
    SGP_TRY(oz_check_ns(ns));
    if (R <= 0 || K <= 0 || (layout != OZ_MN && layout != OZ_K) || (layout == OZ_MN && tri == 2) || (layout == OZ_K && tri == 1) || (layout == OZ_K && ((ld & 1) || ((uintptr_t)X & 15)))) {
        set_error("ozaki_slice: bad arguments"); return ST_BADARG;
    }
    switch (ns) {
    case 4: return oz_slice_t<4>(c, X, ld, R, K, layout, tri, out);
    case 5: return oz_slice_t<5>(c, X, ld, R, K, layout, tri, out);
    case 6: return oz_slice_t<6>(c, X, ld, R, K, layout, tri, out);
    case 7: return oz_slice_t<7>(c, X, ld, R, K, layout, tri, out);
    default: return oz_slice_t<8>(c, X, ld, R, K, layout, tri, out);
    }
}

int ozaki_gemm_sliced(Ctx& c, int ns, const OzSliced& A, const OzSliced& B, long M, long N, double alpha, double beta, double* C, long ldc,
                      int kmode, int lower)
{
    SGP_TRY(oz_check_ns(ns));
    if (M <= 0 || N <= 0 || A.Kp != B.Kp || M > A.Rp || N > B.Rp) { set_error("ozaki_gemm_sliced: operand shapes do not match"); return ST_BADARG; }
    CUtensorMap tmA, tmB;
    SGP_TRY(make_slice_tmap(&tmA, A.sl, A.Rp, A.Kp, ns, OZ_M));
    SGP_TRY(make_slice_tmap(&tmB, B.sl, B.Rp, B.Kp, ns, OZ_N));
    OzArgs a;
    a.ea = A.ex; a.eb = B.ex; a.C = C; a.ldc = ldc; a.M = M; a.N = N;
    a.nk = (int)(A.Kp / OZ_KC); a.kmode = kmode; a.alpha = alpha;
    const int mt = (int)((M + OZ_M - 1) / OZ_M), nt = (int)((N + OZ_N - 1) / OZ_N);
    const int order = (kmode & OZ_KHI_TM) ? 1 : (kmode & OZ_KLO_TN) ? 2 : (kmode & OZ_KHI_TN) ? 3 : 0;
    SGP_TRY(oz_tile_list(c, mt, nt, lower, order, &a.tiles, &a.ntiles));
    // exact INT32 accumulation: at most ns pairs of |digit products| <= 2^14 per k and accumulator -> k-blocks of
    // <= (2^31 - 1) / (ns 2^14) per launch, the blocks added in FP64 (beta = 1 after the first)
    const int kmax = (int)((2147483647L / ((long)ns * 16384L)) / OZ_KC);          // in chunks
    // The limit is on the depth a TILE accumulates, so the blocks are counted from the side its k-range is anchored at: from
    // the end for the k >= ... modes (lauum: only the long rows of the triangle take part in a second launch).
    const int nblk = (a.nk + kmax - 1) / kmax, per = (a.nk + nblk - 1) / nblk;
    const bool from_end = (kmode & (OZ_KLO_TM | OZ_KLO_TN)) && !(kmode & (OZ_KHI_TM | OZ_KHI_TN));
    for (int b = 0; b < nblk; b++) {
        if (from_end) { a.kc_hi = a.nk - b * per; a.kc_lo = a.nk - (b + 1) * per > 0 ? a.nk - (b + 1) * per : 0; }
        else { a.kc_lo = b * per; a.kc_hi = (b + 1) * per < a.nk ? (b + 1) * per : a.nk; }
        a.beta = b == 0 ? beta : 1.0;
        SGP_TRY(oz_ticket(c, &a.ticket));
        switch (ns) {
        case 4: SGP_TRY(oz_launch<4>(c, tmA, tmB, a)); break;
        case 5: SGP_TRY(oz_launch<5>(c, tmA, tmB, a)); break;
        case 6: SGP_TRY(oz_launch<6>(c, tmA, tmB, a)); break;
        case 7: SGP_TRY(oz_launch<7>(c, tmA, tmB, a)); break;
        default: SGP_TRY(oz_launch<8>(c, tmA, tmB, a)); break;
        }
    }
    return ST_OK;
}

size_t ozaki_workspace_bytes(long M, long N, long K, int ns)
{
    return ozaki_sliced_bytes(M, K, ns) + ozaki_sliced_bytes(N, K, ns) + 512;
}

static int oz_check(int ns, long M, long N, long K, size_t work_bytes)
{
    if (M <= 0 || N <= 0 || K <= 0) { set_error("ozaki_gemm: bad arguments"); return ST_BADARG; }
    SGP_TRY(oz_check_ns(ns));
    if (work_bytes < ozaki_workspace_bytes(M, N, K, ns)) { set_error("ozaki_gemm: workspace too small"); return ST_BADARG; }
    return ST_OK;
}

// the GEMM on operands that ozaki_slice_operands() has already split into `work`
int ozaki_gemm_presliced(Ctx& c, int ns, long M, long N, long K, double alpha, double beta, double* C, long ldc, void* work, size_t work_bytes)
{
    SGP_TRY(oz_check(ns, M, N, K, work_bytes));
    const OzSliced A = ozaki_carve(work, M, K, ns);
    const OzSliced B = ozaki_carve((char*)work + ozaki_sliced_bytes(M, K, ns), N, K, ns);
    return ozaki_gemm_sliced(c, ns, A, B, M, N, alpha, beta, C, ldc, 0, 0);
}

// split A (M x K) and B (N x K) (element (r, k) at ptr[r + k ld]) into ns INT8 slices + row exponents in `work`
int ozaki_slice_operands(Ctx& c, int ns, long M, long N, long K, const double* A, long lda, const double* B, long ldb, void* work, size_t work_bytes)
{
    SGP_TRY(oz_check(ns, M, N, K, work_bytes));
    const OzSliced sA = ozaki_carve(work, M, K, ns);
    const OzSliced sB = ozaki_carve((char*)work + ozaki_sliced_bytes(M, K, ns), N, K, ns);
    SGP_TRY(ozaki_slice(c, ns, A, lda, M, K, OZ_MN, 0, sA));
    return ozaki_slice(c, ns, B, ldb, N, K, OZ_MN, 0, sB);
}

size_t ozaki_lauum_workspace_bytes(long n_pad, int ns) { return ozaki_sliced_bytes(n_pad, n_pad, ns) + 256; }

// W (lower tiles, column-major, ldw) = X^T X for the lower-triangular X (n_pad x n_pad, column-major, ldx; n_pad a multiple of
// 128) -- the lauum stage of the inverse -- on the INT8 tensor pipe: ONE operand is sliced (A = B = X^T, whose rows are the
// contiguous columns of X), the tile set is the lower triangle and every tile skips the k-chunks above its rows.
int ozaki_lauum(Ctx& c, int ns, const double* X, long n_pad, long ldx, double* W, long ldw, void* work, size_t work_bytes)
{
    if (n_pad <= 0 || n_pad % OZ_M) { set_error("ozaki_lauum: bad arguments"); return ST_BADARG; }
    SGP_TRY(oz_check_ns(ns));
    if (work_bytes < ozaki_lauum_workspace_bytes(n_pad, ns)) { set_error("ozaki_lauum: workspace too small"); return ST_BADARG; }
    const OzSliced S = ozaki_carve(work, n_pad, n_pad, ns);
    SGP_TRY(ozaki_slice(c, ns, X, ldx, n_pad, n_pad, OZ_K, 2, S));
    return ozaki_gemm_sliced(c, ns, S, S, n_pad, n_pad, 1.0, 0.0, W, ldw, OZ_KLO_TM, 1);
}

// C (M x N, column-major, ldc) = alpha A B^T + beta C with A (M x K) and B (N x K) given as element (r, k) at ptr[r + k ld];
// ns = 4..8 digits per operand (6: 2^-47 of the row scales; 7: 2^-55).
int ozaki_gemm(Ctx& c, int ns, long M, long N, long K, double alpha, const double* A, long lda, const double* B, long ldb, double beta,
               double* C, long ldc, void* work, size_t work_bytes)
{
    SGP_TRY(ozaki_slice_operands(c, ns, M, N, K, A, lda, B, ldb, work, work_bytes));
    return ozaki_gemm_presliced(c, ns, M, N, K, alpha, beta, C, ldc, work, work_bytes);
}

int i8mma_selftest(Ctx& c, int K, int* mismatches, int* first_bad_ref, int* first_bad_got)
{
    if (K <= 0 || K % ST_KC) { set_error("i8mma_selftest: K must be a positive multiple of %d", ST_KC); return ST_BADARG; }
    const size_t szA = (size_t)ST_M * K, szB = (size_t)ST_N * K, szD = (size_t)ST_M * ST_N;
    SGP_TRY(c.io.reserve(szA + szB + 2 * szD * sizeof(int32_t) + 64));
    int8_t* dA = c.io.as<int8_t>();
    int8_t* dB = dA + szA;
    int32_t* dD = reinterpret_cast<int32_t*>(c.io.as<char>() + ((szA + szB + 15) & ~(size_t)15));
    int32_t* dR = dD + szD;
    int* dbad = reinterpret_cast<int*>(dR + szD);
    i8_fill_kernel<<<64, 256, 0, c.stream>>>(dA, (long)szA, 11ull);
    i8_fill_kernel<<<64, 256, 0, c.stream>>>(dB, (long)szB, 23ull);
    SGP_CUDA(cudaMemsetAsync(dD, 0xff, szD * sizeof(int32_t), c.stream));
    SGP_CUDA(cudaMemsetAsync(dbad, 0, sizeof(int), c.stream));
    const size_t smem = (size_t)(ST_M + ST_N) * ST_KC + 1024;
    i8mma_selftest_kernel<<<1, 128, smem, c.stream>>>(dA, dB, K, dD);
    SGP_CUDA(cudaGetLastError());
    i8_ref_kernel<<<(unsigned)((szD + 127) / 128), 128, 0, c.stream>>>(dA, dB, ST_M, ST_N, K, dR);
    i32_diff_kernel<<<(unsigned)((szD + 127) / 128), 128, 0, c.stream>>>(dD, dR, (int)szD, dbad);
    SGP_CUDA(cudaGetLastError());
    count_launch(5);
    int bad = -1;
    SGP_CUDA(cudaMemcpyAsync(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    int32_t r0[2] = {0, 0};
    SGP_CUDA(cudaMemcpyAsync(&r0[0], dR + 5 * ST_N + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
    SGP_CUDA(cudaMemcpyAsync(&r0[1], dD + 5 * ST_N + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
    SGP_CUDA(cudaStreamSynchronize(c.stream));
    if (mismatches) *mismatches = bad;
    if (first_bad_ref) *first_bad_ref = r0[0];
    if (first_bad_got) *first_bad_got = r0[1];
    return ST_OK;
}

}  // namespace sgp
