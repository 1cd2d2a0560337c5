// common.cuh — process-wide context, error plumbing and warp helpers of libbpk.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/bpk.h"

struct BpkCtx {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;
    int *d_flag = nullptr;          // device status word set by kernels (non-SPD, domain)
    int *h_flag = nullptr;          // pinned mirror
    double *scratch = nullptr;      // reduction partials
    size_t scratch_bytes = 0;
    void *l2buf = nullptr;          // > L2 sized buffer for bpk_flush_l2
    size_t l2bytes = 0;
    uint64_t launches = 0;
    char err[512] = {0};
};

extern BpkCtx g_bpk;

// peer-memory exchange window (runtime.cu), one per rank, mapped into every peer through CUDA IPC.
// Layout in 8-byte words:
//   [0]                               exchanges completed by this rank (written by its own kernels only)
//   [BPK_XCHG_DATA + ((par*R + r)*CAP + e)*2 .. +2)
//                                     rank r's deposit of element e with sequence parity par, as two
//                                     "LL" packets {seq32 : low half} {seq32 : high half} of the double.
//   [BPK_XCHG_TOTALS + (par*CAP + e)*2 .. +2)
//                                     the sum over ranks of element e, same packet format (local hand-off from
//                                     the CTA that owns e to the CTA that runs the sweep's small ops).
// Every 8-byte packet carries its own sequence tag and is written with ONE atomic store, so a reader
// that sees the expected tag in both packets has the whole value: no fence, no separate flag, one NVLink
// one-way latency per exchange (the protocol of NCCL's LL mode, widened to fp64 payloads).
#define BPK_XCHG_MAXRANKS 8
#define BPK_XCHG_CAP 2048
#define BPK_XCHG_DATA 64
#define BPK_XCHG_TOTALS (BPK_XCHG_DATA + 2 * 2 * BPK_XCHG_MAXRANKS * BPK_XCHG_CAP)
#define BPK_XCHG_BYTES ((BPK_XCHG_TOTALS + 2 * 2 * BPK_XCHG_CAP) * sizeof(double))
struct BpkXchg {
    bool ready = false;
    int nranks = 1, rank = 0;
    double *own = nullptr;
    double *win[BPK_XCHG_MAXRANKS] = {nullptr};
};
extern BpkXchg g_xchg;
int bpk_xchg_local(void);            // make sure this rank's own window exists (single-GPU hand-off)

int bpk_set_error(int code, const char *fmt, ...);
int bpk_check_flag(int what_if_set);    // sync + read d_flag, clear it
// gmm.cu: bpk_gmm_sweep with a stop word (a raised word turns the launches into no-ops) — for gmm_vb.cu
int bpk_gmm_sweep_resident(const double *Y, int64_t N, int D, int K, const double *c, const double *h, const double *Lam,
                           const double *logpi, double *P, double *g, double *stats, const int *stop);
double *bpk_scratch(size_t bytes);      // stream-ordered scratch (grown on demand)

#define BPK_REQUIRE_INIT()                                                      \
    do {                                                                        \
        if (!g_bpk.ready)                                                       \
            return bpk_set_error(BPK_ENOGPU, "bpk_init() has not succeeded: no CUDA device bound"); \
    } while (0)

#define BPK_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess)                                                  \
            return bpk_set_error(BPK_ECUDA, "%s failed: %s (%s:%d)", #call,     \
                                 cudaGetErrorString(e_), __FILE__, __LINE__);   \
    } while (0)

// every kernel launch goes through this so bpk_launch_count() is honest
#define BPK_LAUNCH(kernel, grid, block, smem, ...)                              \
    do {                                                                        \
        kernel<<<(grid), (block), (smem), g_bpk.stream>>>(__VA_ARGS__);         \
        g_bpk.launches++;                                                       \
        cudaError_t e_ = cudaPeekAtLastError();                                 \
        if (e_ != cudaSuccess)                                                  \
            return bpk_set_error(BPK_ECUDA, "launch of %s failed: %s (%s:%d)",  \
                                 #kernel, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define BPK_FLAG_NOTSPD 1
#define BPK_FLAG_DOMAIN 2

// ---- device helpers -------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// psi(x), fp64, |err| ~ 1e-15 for x > 0 (recurrence up to x >= 10, then the
// asymptotic series); reflection for x <= 0.  Matches scipy.special.psi to
// ~1e-14 relative away from the root at 1.4616.
__device__ __forceinline__ double bpk_digamma(double x) {
    double r = 0.0;
    if (x <= 0.0) {
        if (x == floor(x)) return nan("");
        // psi(1-x) - psi(x) = pi cot(pi x)
        r = -M_PI / tan(M_PI * x);
        x = 1.0 - x;
    }
    while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
    double xi = 1.0 / x, x2 = xi * xi;
    // B2k/(2k): 1/12, -1/120, 1/252, -1/240, 1/132, -691/32760, 1/12
    double s = x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0 - x2 * (1.0 / 240.0
               - x2 * (1.0 / 132.0 - x2 * (691.0 / 32760.0 - x2 * (1.0 / 12.0)))))));
    return r + log(x) - 0.5 * xi - s;
}
// psi'(x) for x > 0: recurrence psi'(x) = psi'(x+1) + 1/x^2 up to x >= 10, then the asymptotic series
// 1/x + 1/(2x^2) + sum_k B_2k / x^(2k+1); 9e-16 relative against scipy.special.polygamma(1, x) on [1e-3, 1e8].
__device__ __forceinline__ double bpk_trigamma(double x) {
    if (!(x > 0.0)) return nan("");
    double r = 0.0;
    while (x < 10.0) { r += 1.0 / (x * x); x += 1.0; }
    const double xi = 1.0 / x, x2 = xi * xi;
    const double s = xi * x2 * (1.0 / 6.0 - x2 * (1.0 / 30.0 - x2 * (1.0 / 42.0 - x2 * (1.0 / 30.0
                     - x2 * (5.0 / 66.0 - x2 * (691.0 / 2730.0 - x2 * (7.0 / 6.0)))))));
    return r + xi + 0.5 * x2 + s;
}
__device__ __forceinline__ double bpk_mvdigamma(double a, int d) {
    double s = 0.0;
    for (int i = 0; i < d; ++i) s += bpk_digamma(a - 0.5 * i);
    return s;
}
__device__ __forceinline__ double bpk_mvlgamma(double a, int d) {
    // scipy.special.multigammaln: d(d-1)/4 log(pi) + sum_j lgamma(a - j/2)
    double s = d * (d - 1) * 0.25 * 1.1447298858494001741434;
    for (int j = 0; j < d; ++j) s += lgamma(a - 0.5 * j);
    return s;
}
