// Warp-specialised, persistent version of the structured FP64 DMMA GEMM (dmma_gemm.cuh has the
// operand layouts, tile modes and the GemmArgs contract; this file only changes HOW a CTA runs).
//
// One CTA per SM, 10 warps:
//   warps 8, 9  producers: stream the A and B k-slabs of the CTA's tiles global -> shared into the
//               same padded, bank-conflict-free layouts as before; LAYOUT_MN slabs by cp.async.bulk
//               (SASS UBLKCP, the TMA engine's 1-D path, one 1 KB copy per k-row, completion by
//               expect_tx / complete_tx), LAYOUT_K slabs by 16-byte LDGSTS (completion by
//               cp.async.mbarrier.arrive.noinc, rows split over both producer warps); both kinds land
//               on one "full" mbarrier per stage.
//   warps 0..7  consumers: 2(M) x 4(N) warp grid, 64x32 warp tile, 32 DMMA.8x8x4 per k4 step fed by
//               12 LDS.64; fragments are double-buffered in registers so the shared loads of step
//               k+1 are in flight while the DMMAs of step k issue.
// A ring of WS_STAGES stages is handed back and forth with full/empty mbarriers: there is no
// __syncthreads in the main loop, so the two consumer warps of each SM sub-partition drift out of
// phase and one keeps the DMMA pipe busy while the other waits, loads or stores.  The CTA is
// persistent (tiles blockIdx.x, blockIdx.x + gridDim.x, ...; long k-ranges first), so the producer
// is already filling the ring for the next tile while the consumers write the current one back:
// prologue latency is paid once per CTA, not once per tile.
//
// (v1 measured on B200: 82 % DMMA pipe utilisation at K = 8192, 35-75 % for K <= 1024, with the
// block-wide barrier + per-stage address arithmetic of the cp.async ring as the top stalls;
// profiles/r01_gemm_v1_ncu.txt.)
#pragma once
#include "dmma_gemm.cuh"
#include "mbar.cuh"

namespace sgp {

constexpr int WS_STAGES = 5;
constexpr int WS_CONSUMER_WARPS = 8;
constexpr int WS_THREADS = (WS_CONSUMER_WARPS + 1) * 32;       // potrf_ll.cu: one producer warp (all operands LAYOUT_MN)
// the GEMM kernel runs two producer warps: LAYOUT_K slabs take 32 LDGSTS per lane with one warp, and ncu
// showed the consumers waiting on the full barriers then (DMMA pipe 86 % against 95 % for bulk-copied slabs)
constexpr int WS_PRODUCER_WARPS = 2;
constexpr int WS_GEMM_THREADS = (WS_CONSUMER_WARPS + WS_PRODUCER_WARPS) * 32;
constexpr uint32_t WS_SLAB_BYTES = GT * GK * sizeof(double);               // bytes of one operand slab (128 x GK)
constexpr size_t WS_SMEM = (size_t)WS_STAGES * 2 * STAGE_DOUBLES * sizeof(double) + 2 * WS_STAGES * sizeof(unsigned long long);

struct TileInfo {
    int tm, tn;
    long k0;
    int nk;       // k-slabs of GK
};

// One problem of a GROUPED launch: independent GEMMs (the nodes of one level of the recursive triangular
// inverse, chol.cu) share one persistent launch; global tile index t belongs to the entry with
// tile0 <= t < tile0 + gemm_num_tiles(a).
struct GemmGroupEntry {
    GemmArgs a;
    long tile0;
};

// global tile index -> (problem, tile index inside the problem); entries are sorted by tile0
__device__ __forceinline__ void group_lookup(const GemmGroupEntry* __restrict__ grp, int nprob, long tile, GemmArgs& p, long& b)
{
    int lo = 0, hi = nprob - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (grp[mid].tile0 <= tile) lo = mid; else hi = mid - 1;
    }
    p = grp[lo].a;
    b = tile - grp[lo].tile0;
}

// tile index -> position and k-range (same enumeration as gemm_f64_kernel: long k-ranges first)
__device__ __forceinline__ TileInfo tile_info(const GemmArgs& p, long b)
{
    TileInfo t;
    long k0 = 0, k1 = p.K;
    if (p.mode == TM_LOWER || p.mode == TM_LOWER_KGE) {
        long r = (long)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
        while ((r + 1) * (r + 2) / 2 <= b) r++;
        while (r * (r + 1) / 2 > b) r--;
        t.tm = (int)r;
        t.tn = (int)(b - r * (r + 1) / 2);
        if (p.mode == TM_LOWER_KGE) k0 = (long)t.tm * GT;
    } else if (p.mode == TM_A_LOWER) {
        t.tm = p.Mt - 1 - (int)(b / p.Nt);
        t.tn = (int)(b % p.Nt);
        k1 = (long)(t.tm + 1) * GT;
        if (k1 > p.K) k1 = p.K;
    } else {
        t.tm = (int)(b % p.Mt);
        t.tn = (int)(b / p.Mt);
        if (p.mode == TM_B_LOWER) k0 = (long)t.tn * GT;
    }
    t.k0 = k0;
    t.nk = (int)((k1 - k0) / GK);
    if (t.nk < 0) t.nk = 0;
    return t;
}

// producer side: one operand slab (128 x GK) of one stage, spread over the 32 lanes.
//  LAYOUT_MN: GK k-rows of 128 contiguous doubles -> one 1 KB cp.async.bulk per row (lanes 0..GK-1;
//             UBLKCP takes uniform operands, the warp issues the rows one after the other).
//  LAYOUT_K : 128 m-rows of GK contiguous doubles (128 B) -> 16-byte LDGSTS, 32 per lane; a bulk copy
//             per 128 B row would cost 128 serialised UBLKCP issues per slab and starve the consumers.
// `part` of `nparts` producer warps: bulk copies are issued by part 0 only, LDGSTS rows are split evenly
template <int LAY>
__device__ __forceinline__ void produce_slab(double* sm, const double* __restrict__ g, long ld, long r0, long kk, int lane,
                                             unsigned long long* bar, int part = 0, int nparts = 1)
{
    if (LAY == LAYOUT_MN) {
        if (part == 0 && lane < GK) bulk_g2s(sm + lane * LDMN, g + (kk + lane) * ld + r0, GT * sizeof(double), bar);
    } else {
        const int rows = GT / nparts, row0 = part * rows;
        const double* src = g + (r0 + row0 + (lane >> 3)) * ld + kk + (lane & 7) * 2;
        double* dst = sm + (row0 + (lane >> 3)) * LDKK + (lane & 7) * 2;
        const long step = 4 * ld;
#pragma unroll 8
        for (int i = 0; i < rows / 4; i++) cp_async16(dst + i * 4 * LDKK, src + i * step);
    }
}

template <int LAY>
__device__ __forceinline__ void load_frags_a(const double* sa, int wm, int g, int t, int kk, double (&af)[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) af[i] = frag<LAY>(sa, wm + i * 8 + g, kk + t);
}
template <int LAY>
__device__ __forceinline__ void load_frags_b(const double* sb, int wn, int g, int t, int kk, double (&bf)[4])
{
#pragma unroll
    for (int j = 0; j < 4; j++) bf[j] = frag<LAY>(sb, wn + j * 8 + g, kk + t);
}

// GROUPED = false: one problem `p0`.  GROUPED = true: `nprob` problems `grp` (device memory), `ntiles` tiles in all.
template <int AL, int BL, bool GROUPED>
__global__ void __launch_bounds__(WS_GEMM_THREADS, 1) gemm_f64_ws_kernel(GemmArgs p0, long ntiles, const GemmGroupEntry* __restrict__ grp, int nprob)
{
    extern __shared__ __align__(16) double smem[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)WS_STAGES * 2 * STAGE_DOUBLES);
    unsigned long long* empty = full + WS_STAGES;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WS_STAGES; s++) {
            // producer arrivals per phase: lane 0's arrive.expect_tx when an operand comes by bulk copy,
            // plus one cp.async.mbarrier.arrive.noinc per lane when an operand comes by LDGSTS
            mbar_init(full + s, (AL == LAYOUT_MN || BL == LAYOUT_MN ? 1 : 0) + (AL == LAYOUT_K || BL == LAYOUT_K ? 32 * WS_PRODUCER_WARPS : 0));
            mbar_init(empty + s, WS_CONSUMER_WARPS);      // one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= WS_CONSUMER_WARPS) {
        // ------------------------------------------------------------------ producer warps
        const int part = warp - WS_CONSUMER_WARPS;
        if (part > 0 && AL == LAYOUT_MN && BL == LAYOUT_MN) return;      // nothing to do for bulk-only operands
        uint32_t it = 0;
        for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            GemmArgs p = p0;
            long tb = tile;
            if (GROUPED) group_lookup(grp, nprob, tile, p, tb);
            const TileInfo ti = tile_info(p, tb);
            const long m0 = (long)ti.tm * GT, n0 = (long)ti.tn * GT;
            for (int kb = 0; kb < ti.nk; kb++, it++) {
                const int s = (int)(it % WS_STAGES);
                const uint32_t ph = (it / WS_STAGES) & 1u;
                mbar_wait(empty + s, ph ^ 1u);
                constexpr uint32_t tx = (AL == LAYOUT_MN ? WS_SLAB_BYTES : 0u) + (BL == LAYOUT_MN ? WS_SLAB_BYTES : 0u);
                if (tx != 0u && part == 0) {
                    if (lane == 0) mbar_arrive_expect_tx(full + s, tx);
                    __syncwarp();
                }
                double* sa = smem + (size_t)s * 2 * STAGE_DOUBLES;
                const long kk = ti.k0 + (long)kb * GK;
                produce_slab<AL>(sa, p.A, p.lda, m0, kk, lane, full + s, part, WS_PRODUCER_WARPS);
                produce_slab<BL>(sa + STAGE_DOUBLES, p.B, p.ldb, n0, kk, lane, full + s, part, WS_PRODUCER_WARPS);
                if (AL == LAYOUT_K || BL == LAYOUT_K) cp_async_mbar_arrive_noinc(full + s);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64;
    const int wn = (warp >> 1) * 32;
    uint32_t it = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        GemmArgs p = p0;
        long tb = tile;
        if (GROUPED) group_lookup(grp, nprob, tile, p, tb);
        const double alpha = p.alpha, beta = p.beta;
        const TileInfo ti = tile_info(p, tb);
        const long m0 = (long)ti.tm * GT, n0 = (long)ti.tn * GT;
        double acc[8][4][2];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int kb = 0; kb < ti.nk; kb++, it++) {
            const int s = (int)(it % WS_STAGES);
            const uint32_t ph = (it / WS_STAGES) & 1u;
            mbar_wait(full + s, ph);
            const double* sa = smem + (size_t)s * 2 * STAGE_DOUBLES;
            const double* sb = sa + STAGE_DOUBLES;
            double af[2][8], bf[2][4];
            load_frags_a<AL>(sa, wm, g, t, 0, af[0]);
            load_frags_b<BL>(sb, wn, g, t, 0, bf[0]);
#pragma unroll
            for (int k4 = 0; k4 < GK / 4; k4++) {
                const int cur = k4 & 1, nxt = cur ^ 1;
                if (k4 + 1 < GK / 4) {
                    load_frags_a<AL>(sa, wm, g, t, (k4 + 1) * 4, af[nxt]);
                    load_frags_b<BL>(sb, wn, g, t, (k4 + 1) * 4, bf[nxt]);
                }
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }

        // epilogue: C = beta*C + alpha*acc
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
}

inline int ws_sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int AL, int BL>
inline cudaError_t gemm_ws_launch_t(const GemmArgs& a, cudaStream_t st)
{
    static bool configured[64] = {};
    if (first_use_on_current_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_f64_ws_kernel<AL, BL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM);
        if (e != cudaSuccess) return e;
    }
    const long nt = gemm_num_tiles(a);
    if (nt <= 0 || a.K <= 0) return cudaSuccess;
    const int sms = ws_sm_count();
    const unsigned grid = (unsigned)(nt < sms ? nt : sms);
    gemm_f64_ws_kernel<AL, BL, false><<<grid, WS_GEMM_THREADS, WS_SMEM, st>>>(a, nt, nullptr, 0);
    return cudaGetLastError();
}

// grouped launch: `d_grp` = nprob entries in device memory, `ntiles` = sum of their tile counts
template <int AL, int BL>
inline cudaError_t gemm_ws_launch_grouped_t(const GemmGroupEntry* d_grp, int nprob, long ntiles, cudaStream_t st)
{
    static bool configured[64] = {};
    if (first_use_on_current_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_f64_ws_kernel<AL, BL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM);
        if (e != cudaSuccess) return e;
    }
    if (nprob <= 0 || ntiles <= 0) return cudaSuccess;
    const int sms = ws_sm_count();
    const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
    GemmArgs none{};
    gemm_f64_ws_kernel<AL, BL, true><<<grid, WS_GEMM_THREADS, WS_SMEM, st>>>(none, ntiles, d_grp, nprob);
    return cudaGetLastError();
}

inline cudaError_t gemm_ws_launch(int al, int bl, const GemmArgs& a, cudaStream_t st)
{
    if (al == LAYOUT_MN && bl == LAYOUT_MN) return gemm_ws_launch_t<LAYOUT_MN, LAYOUT_MN>(a, st);
    if (al == LAYOUT_MN && bl == LAYOUT_K) return gemm_ws_launch_t<LAYOUT_MN, LAYOUT_K>(a, st);
    if (al == LAYOUT_K && bl == LAYOUT_K) return gemm_ws_launch_t<LAYOUT_K, LAYOUT_K>(a, st);
    return gemm_ws_launch_t<LAYOUT_K, LAYOUT_MN>(a, st);
}

}  // namespace sgp
