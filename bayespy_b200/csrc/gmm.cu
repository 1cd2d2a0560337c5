// gmm.cu — fused E-step + sufficient statistics of a Gaussian mixture
//   Y = Mixture(Z, Gaussian, mu, Lambda)       
// Reference path being replaced, per VB sweep:
//   Z.update():      mixture.py:53-106 (index 0: L[n,k] = g_k + <phi_k, u_n>, via
//                    expfamily.py:45-61 with (N,K,D,D) temporaries) + softmax
//                    multinomial.py:101-121 / misc.py:1366-1401
//   mu/Lambda/alpha: mixture.py:108-160 (p-weighted [y, yy^T]) plate-summed over N
//                    by node.py:650 — recomputed once per parent
// Here: one pass over y [N][D]:  L = c_k + y.h_k - 1/2 y^T Lam_k y + logpi_k,
// p = softmax_k(L) (with the reference's second renormalisation), and
//   stats = [ sum_n p | sum_n p y | sum_n p y y^T | sum_n logsumexp ].
// Algorithmic traffic: D*8 B read + K*8 B written per row (576 B at D=8, K=64).
//
// v0 mapping: persistent CTAs of 128 threads; a tile of 128 rows per step;
// thread r computes row r's responsibilities with the K parameter blocks
// staged in shared memory (all lanes read the same parameter -> broadcast);
// then the CTA accumulates the K*(1+D+D^2) statistics in registers, each
// thread owning a fixed subset of (k, feature) pairs across all tiles.
// Deterministic: per-CTA partials reduced in a fixed order by a second kernel.
// TODO(next): move both contractions onto the fp64 tensor pipe (DMMA) — the
// quadratic form is a (rows x F) x (F x K) GEMM in the monomial features of y.
#include "common.cuh"

#define GMM_ROWS 128
#define GMM_MAXK 128
#define GMM_MAXD 16

struct GmmArgs {
    const double *Y;
    int64_t N;
    int D, K;
    const double *c, *h, *Lam, *logpi;
    double *P, *g, *partial;
    int given_p;        // 1: P is an input (statistics only), 0: E-step + statistics
    int nfeat;          // 1 + D + D*D
    int per_thread;     // ceil(K*nfeat / 128)
};

template <int PT>
__global__ void __launch_bounds__(GMM_ROWS, 1) gmm_sweep_kernel(GmmArgs a) {
    extern __shared__ double sm[];
    const int D = a.D, K = a.K, F = a.nfeat;
    const int ldp = K | 1;
    double *sc = sm;                       // [K]     c + logpi
    double *sh = sc + K;                   // [K][D]
    double *sL = sh + K * D;               // [K][D][D]
    double *sP = sL + K * D * D;           // [ROWS][ldp]
    double *sZ = sP + GMM_ROWS * ldp;      // [ROWS][F]   features 1, y_i, y_i y_j
    __shared__ double red[GMM_ROWS / 32];
    const int t = threadIdx.x;
    if (!a.given_p) {
        for (int e = t; e < K; e += GMM_ROWS) sc[e] = a.c[e] + a.logpi[e];
        for (int e = t; e < K * D; e += GMM_ROWS) sh[e] = a.h[e];
        for (int e = t; e < K * D * D; e += GMM_ROWS) sL[e] = a.Lam[e];
    }
    double acc[PT];
#pragma unroll
    for (int i = 0; i < PT; ++i) acc[i] = 0.0;
    double lse_acc = 0.0;
    __syncthreads();
    const int64_t ntiles = (a.N + GMM_ROWS - 1) / GMM_ROWS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n = tile * GMM_ROWS + t;
        const bool live = n < a.N;
        double y[GMM_MAXD];
        // coalesced stage of the y tile through sZ, then per-thread registers
        {
            const int64_t base = tile * GMM_ROWS * D;
            int64_t lim = a.N * D - base;
            if (lim > (int64_t)GMM_ROWS * D) lim = (int64_t)GMM_ROWS * D;
            for (int e = t; e < GMM_ROWS * D; e += GMM_ROWS) {
                double v = e < lim ? a.Y[base + e] : 0.0;
                sZ[(e / D) * F + 1 + (e % D)] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < GMM_MAXD; ++d) y[d] = d < D ? sZ[t * F + 1 + d] : 0.0;
        sZ[t * F] = live ? 1.0 : 0.0;
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) sZ[t * F + 1 + D + i * D + j] = y[i] * y[j];
        double lse = 0.0;
        if (a.given_p) {
            // responsibilities are an input: stage them (coalesced) and skip the E-step
            __syncthreads();
            const int64_t pbase = tile * GMM_ROWS * K;
            int64_t plim = a.N * K - pbase;
            if (plim > (int64_t)GMM_ROWS * K) plim = (int64_t)GMM_ROWS * K;
            for (int e = t; e < GMM_ROWS * K; e += GMM_ROWS)
                sP[(e / K) * ldp + (e % K)] = e < plim ? a.P[pbase + e] : 0.0;
        } else {
        // log-evidence of each component (mixture.py:58-106)
        double m = -INFINITY;
        for (int k = 0; k < K; ++k) {
            double q = 0.0, lin = 0.0;
            const double *Lk = sL + k * D * D;
            for (int i = 0; i < D; ++i) {
                double s = 0.0;
                for (int j = 0; j < D; ++j) s += Lk[i * D + j] * y[j];
                q += s * y[i];
                lin += sh[k * D + i] * y[i];
            }
            double L = sc[k] + lin - 0.5 * q;
            sP[t * ldp + k] = L;
            m = fmax(m, L);
        }
        double mm = isfinite(m) ? m : 0.0;
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += exp(sP[t * ldp + k] - mm);
        lse = log(s) + mm;
        double s2 = 0.0;
        for (int k = 0; k < K; ++k) {
            double p = exp(sP[t * ldp + k] - lse);
            sP[t * ldp + k] = p;
            s2 += p;
        }
        for (int k = 0; k < K; ++k) {
            double p = live ? sP[t * ldp + k] / s2 : 0.0;
            sP[t * ldp + k] = p;
        }
        }
        if (live && !a.given_p) {
            lse_acc += lse;
            if (a.g) a.g[n] = -lse;
        }
        __syncthreads();
        // coalesced store of the responsibilities
        if (a.P && !a.given_p) {
            const int64_t base = tile * GMM_ROWS * K;
            int64_t lim = a.N * K - base;
            if (lim > (int64_t)GMM_ROWS * K) lim = (int64_t)GMM_ROWS * K;
            for (int e = t; e < lim; e += GMM_ROWS) a.P[base + e] = sP[(e / K) * ldp + (e % K)];
        }
        // statistics: thread owns outputs o = t + i*128  ->  (k, f) = (o / F, o % F)
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            int o = t + i * GMM_ROWS;
            if (o < K * F) {
                int k = o / F, f = o - k * F;
                double sacc = 0.0;
                for (int r = 0; r < GMM_ROWS; ++r) sacc += sP[r * ldp + k] * sZ[r * F + f];
                acc[i] += sacc;
            }
        }
        __syncthreads();
    }
    // per-CTA partials: layout [K*F | 1]
    double *pout = a.partial + (size_t)blockIdx.x * (K * F + 1);
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        int o = t + i * GMM_ROWS;
        if (o < K * F) pout[o] = acc[i];
    }
    lse_acc = warp_sum(lse_acc);
    if ((t & 31) == 0) red[t >> 5] = lse_acc;
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int w = 0; w < GMM_ROWS / 32; ++w) s += red[w];
        pout[K * F] = s;
    }
}

// partial (k, f) layout -> caller's [sum p (K) | sum p y (K*D) | sum p yy^T (K*D*D) | lse]
__global__ void gmm_final_kernel(const double *__restrict__ partial, int nblocks, int K, int D,
                                 double *__restrict__ stats) {
    const int F = 1 + D + D * D;
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    int total = K * F + 1;
    if (e >= total) return;
    int src;
    if (e < K) src = e * F;
    else if (e < K + K * D) { int r = e - K; src = (r / D) * F + 1 + (r % D); }
    else if (e < K * F) { int r = e - K - K * D; src = (r / (D * D)) * F + 1 + D + (r % (D * D)); }
    else src = K * F;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * total + src];
    stats[e] += s;
}

static int gmm_run(const double *Y, int64_t N, int D, int K,
                   const double *c, const double *h, const double *Lam, const double *logpi,
                   double *P, double *g, double *stats, int given_p);

extern "C" int bpk_gmm_sweep(const double *Y, int64_t N, int D, int K,
                             const double *c, const double *h, const double *Lam, const double *logpi,
                             double *P, double *g, double *stats) {
    BPK_REQUIRE_INIT();
    return gmm_run(Y, N, D, K, c, h, Lam, logpi, P, g, stats, 0);
}

extern "C" int bpk_gmm_stats(const double *Y, int64_t N, int D, int K, const double *P, double *stats) {
    BPK_REQUIRE_INIT();
    if (!P) return bpk_set_error(BPK_EINVAL, "bpk_gmm_stats: P is required");
    return gmm_run(Y, N, D, K, nullptr, nullptr, nullptr, nullptr, const_cast<double *>(P), nullptr, stats, 1);
}

static int gmm_run(const double *Y, int64_t N, int D, int K,
                   const double *c, const double *h, const double *Lam, const double *logpi,
                   double *P, double *g, double *stats, int given_p) {
    if (D < 1 || D > GMM_MAXD) return bpk_set_error(BPK_EINVAL, "bpk_gmm_sweep: D=%d outside [1,%d]", D, GMM_MAXD);
    if (K < 1 || K > GMM_MAXK) return bpk_set_error(BPK_EINVAL, "bpk_gmm_sweep: K=%d outside [1,%d]", K, GMM_MAXK);
    if (N <= 0) return BPK_OK;
    GmmArgs a;
    a.Y = Y; a.N = N; a.D = D; a.K = K; a.c = c; a.h = h; a.Lam = Lam; a.logpi = logpi; a.P = P; a.g = g;
    a.given_p = given_p;
    a.nfeat = 1 + D + D * D;
    a.per_thread = (K * a.nfeat + GMM_ROWS - 1) / GMM_ROWS;
    size_t smem = ((size_t)K * (1 + D + D * D) + (size_t)GMM_ROWS * ((K | 1) + a.nfeat)) * sizeof(double);
    if (smem > (220u << 10)) return bpk_set_error(BPK_EINVAL, "bpk_gmm_sweep: K=%d, D=%d needs %zu B of shared memory", K, D, smem);
    int64_t ntiles = (N + GMM_ROWS - 1) / GMM_ROWS;
    int grid = g_bpk.sm_count;
    if (ntiles < grid) grid = (int)ntiles;
    a.partial = bpk_scratch((size_t)grid * (K * a.nfeat + 1) * sizeof(double));
    if (!a.partial) return bpk_set_error(BPK_ECUDA, "gmm: scratch allocation failed");
#define GMM_LAUNCH(PT)                                                                                  \
    do {                                                                                                \
        BPK_CUDA(cudaFuncSetAttribute(gmm_sweep_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        BPK_LAUNCH(gmm_sweep_kernel<PT>, grid, GMM_ROWS, smem, a);                                      \
    } while (0)
    if (a.per_thread <= 8) GMM_LAUNCH(8);
    else if (a.per_thread <= 40) GMM_LAUNCH(40);
    else if (a.per_thread <= 96) GMM_LAUNCH(96);
    else return bpk_set_error(BPK_EINVAL, "bpk_gmm_sweep: K*(1+D+D^2)=%d too large for the v0 kernel", K * a.nfeat);
#undef GMM_LAUNCH
    int total = K * a.nfeat + 1;
    BPK_LAUNCH(gmm_final_kernel, (total + 127) / 128, 128, 0, a.partial, grid, K, D, stats);
    return BPK_OK;
}
