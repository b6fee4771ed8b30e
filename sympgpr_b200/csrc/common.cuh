// Shared host-side plumbing of libsympgpr_b200: context, workspace, error codes.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "forms.cuh"

namespace sgp {

// status codes of the C ABI (include/sympgpr_b200.h)
constexpr int ST_OK = 0;
constexpr int ST_BADARG = -1;
constexpr int ST_CUDA = -2;
constexpr int ST_NOMEM = -3;
constexpr int ST_NODEV = -4;

void set_error(const char* fmt, ...);
void count_launch(unsigned long long n = 1ull);   // kernels launched by this library (bench.py gpu_launches)
unsigned long long launch_count();
const char* get_error();

#define SGP_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            sgp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (e__ == cudaErrorMemoryAllocation) ? sgp::ST_NOMEM : sgp::ST_CUDA;        \
        }                                                                                     \
    } while (0)

#define SGP_TRY(call)            \
    do {                         \
        int s__ = (call);        \
        if (s__ != 0) return s__; \
    } while (0)

// A growable device buffer.
struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return ST_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            cudaGetLastError();
            return ST_NOMEM;
        }
        cap = bytes;
        return ST_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// cudaFuncSetAttribute applies to the CURRENT device: a process that drives several devices (one context each) must opt every
// kernel into its large dynamic shared memory once per device.  Returns true the first time it is called for (flags, device).
inline bool first_use_on_current_device(bool (&flags)[64])
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;        // unknown: configure again (harmless)
    if (flags[dev]) return false;
    flags[dev] = true;
    return true;
}

constexpr int TILE = 128;
inline long round_up(long v, long m) { return (v + m - 1) / m * m; }

// Result block written by the device-side finalisation of an NLL evaluation.
//   [0] nll  [1] dlx  [2] dly  [3] dsig  [4] info (0 ok, >0 first non-positive pivot, 1-based)
//   [5] 0.5*y'alpha  [6] sum log diag(L)  [7] reserved
//   [8..10] A_lx, A_ly, A_sig = alpha' dK alpha   [11..13] B_lx, B_ly, B_sig = trace(Kyinv dK)
constexpr int RES_DOUBLES = 16;

// stage boundaries of one NLL(+gradient) evaluation: start | fill | potrf | potrs | trtri | lauum | grad | finalize
constexpr int NSTAGE_EV = 8;

// alpha = Kyinv z of one (Kyinv, z) pair the f2py-signature entry points were handed, kept on the device: the
// reference recomputes this matvec inside every guessP / calcq / target call (sympgpr.f90:72,85,121); an unchanged
// Python loop makes 2 E S such calls with the same arrays, so re-uploading n^2 doubles per call would dominate.
struct AlphaEntry {
    const double* kyinv = nullptr;    // host addresses + order + sampled checksum identify the pair
    const double* z = nullptr;
    long n = 0;
    unsigned long long sum = 0ull, stamp = 0ull;
    DBuf buf;                         // [alpha (n) | z staging (n)]
};
constexpr int ALPHA_CACHE = 8;

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;    // main stream (borrowed or owned)
    bool own_stream = false;
    DBuf Kmat, Wmat, Tmat, Dinv, vecs, pts, partial, small, mapbuf, io, flags, ozbuf;
    int ozaki_slices = 0;             // > 0: OPT-IN, stages of the inverse run on the INT8 tensor pipe (ozaki.cu, ozaki_chol.cu)
    int ozaki_stages = 1;             // bit 0: lauum (W = X^T X); bit 1: factor + triangular inverse (ozaki_factinv)
    long ozaki_leaf = 4096;           // blocks of at most this many rows stay on the DMMA kernels
    AlphaEntry acache[ALPHA_CACHE];
    unsigned long long aclock = 0ull, ahits = 0ull, amisses = 0ull;
    double* h_res = nullptr;          // pinned host staging (RES_DOUBLES + spare)
    int sm_count = 148;
    // optional stage timers of one NLL evaluation (sgp_set_profiling / sgp_stage_times)
    bool prof = false;
    cudaEvent_t pev[NSTAGE_EV] = {};
    bool pev_valid = false;           // pev[] recorded by the last nll_enqueue
    int mark(int k)
    {
        if (!prof) return ST_OK;
        cudaError_t e = cudaEventRecord(pev[k], stream);
        if (e != cudaSuccess) { set_error("cudaEventRecord: %s", cudaGetErrorString(e)); return ST_CUDA; }
        return ST_OK;
    }
};

}  // namespace sgp
