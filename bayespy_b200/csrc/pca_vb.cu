// pca_vb.cu — the device-resident VB loop of the plated Gaussian factor model (Bayesian PCA)
//
//   X = GaussianARD(mu_x, a_x, plates=(1,N), shape=(K,))        "col" factor
//   C = GaussianARD(mu_c, alpha, plates=(M,1), shape=(K,))      "row" factor, alpha ~ Gamma(a0,b0) plates (K,)
//   Y = GaussianARD(SumMultiply('k,k->', X, C), tau),  tau ~ Gamma(ta0,tb0),  Y fully observed
//
// What it replaces: the Python scheduler loop of VB.update (vmp.py:132-172) together with the
// per-node update / lower-bound calls it makes for this model — X: gaussian.py:649-706 fed by
// dot.py:581; C: the same with the roles swapped; alpha, tau: gamma.py:96-148 fed by
// gaussian.py:609-637 and :2351-2371; bound: expfamily.py:400-480 for every node, then the
// convergence test of vmp.py:717-747.  The reference runs ~10^2 NumPy calls per node per sweep;
// the first GPU version of this engine issued ~100 tiny kernels per sweep from Python and was
// host-bound (3.2 ms of dispatch around a 2.1 ms sweep kernel).  Here one sweep is
//       [pca_xsweep_kernel]  ->  [pca_vb_small_kernel: STATS, C, alpha, tau, BOUND, XPRE]
// with nothing read back until a chunk of sweeps has been enqueued.  The update ORDER is the
// user's: the host passes the per-iteration program as a list of opcodes (one per node), and
// runs of small ops between two X sweeps are executed by one single-CTA launch.
//
// Convergence is decided ON DEVICE (same test as vmp.py:738-747) and raises a stop word that
// every later kernel of the chunk checks on entry, so the posterior freezes at exactly the
// iteration the reference would have returned at.
//
// All persistent quantities live in one fp64 state vector (layout: pca_vb_offsets below,
// exported through bpk_pca_vb_layout); after the run the node objects of the Python graph
// simply view slices of it.
#include "common.cuh"
#include "pca_common.cuh"
#include "spd.cuh"

#define VB_THREADS 256
#define VB_MAXOPS 24
#define LOG2PI_D 1.8378770664093454835606594728112

enum {
    F_MUX = 0, F_AX, F_MUC, F_A0, F_B0, F_TA0, F_TB0, F_SUMSQ, F_NG, F_RESERVED,
    F_W, F_SWW, F_COVC, F_LAMC, F_LOGDETC, F_PHI0C, F_GC,
    F_AL_PHI0, F_AL_PHI1, F_AL_U0, F_AL_U1, F_AL_G,
    F_TAU_PHI0, F_TAU_PHI1, F_TAU_U0, F_TAU_U1, F_TAU_G,
    F_COVX, F_LAMX, F_LOGDETX, F_A, F_BX,
    F_STATS, F_SXXT, F_LPREV, F_STATS_LOCAL,
    F_COUNT
};

static const char *const kFieldNames[F_COUNT] = {
    "mux", "ax", "muc", "a0", "b0", "ta0", "tb0", "sumsq", "ng", "reserved",
    "w", "sww", "covc", "lamc", "logdetc", "phi0c", "gc",
    "al_phi0", "al_phi1", "al_u0", "al_u1", "al_g",
    "tau_phi0", "tau_phi1", "tau_u0", "tau_u1", "tau_g",
    "covx", "lamx", "logdetx", "A", "bx",
    "stats", "sxxt", "lprev", "stats_local",
};

__host__ __device__ inline void pca_vb_offsets(int M, int K, int64_t *off /* [F_COUNT+1] */) {
    const int64_t KK = (int64_t)K * K, MK = (int64_t)M * K, NS = MK + KK + K;
    const int64_t size[F_COUNT] = {
        K, K, K, K, K, 1, 1, 1, 1, 1,
        MK, KK, KK, KK, 1, MK, M,
        K, K, K, K, K,
        1, 1, 1, 1, 1,
        KK, KK, 1, MK, K,
        NS, KK, 1, NS,
    };
    int64_t o = 0;
    for (int f = 0; f < F_COUNT; ++f) { off[f] = o; o += size[f]; }
    off[F_COUNT] = o;
}

struct PcaVbArgs {
    int M, K, has_alpha, has_tau;
    double tol;
    double *st;
    const double *partial;   // per-CTA partials of the preceding sweep kernel (padded layout) or NULL
    int nparts;
    int local_stats;         // 1: STATS writes F_STATS_LOCAL (an all-reduce into F_STATS follows)
    double *Lhist;           // [cap][6]: Y, X, C, alpha, tau, total
    int cap;
    int *ctrl;               // [0] iterations finished, [1] stop, [2] error bits
    int nops;
    int ops[VB_MAXOPS];
};

__device__ __forceinline__ double vb_block_sum(double v, double *red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < VB_THREADS / 32; ++w) s += red[w];
    return s;
}

// S filled by the whole CTA -> warp 0 factors and inverts it into C; scal[0] = log det
__device__ __forceinline__ void vb_spd_inverse(double *S, double *B, double *C, int K, int ld, double *scal, int *ctrl) {
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int bad = spd_warp_chol_upper(S, K, ld, lane);
        double ldt = spd_warp_logdet(S, K, ld, lane);
        spd_warp_inverse(S, B, C, K, ld, lane);
        if (lane == 0) {
            scal[0] = ldt;
            if (bad) atomicOr(&ctrl[2], BPK_FLAG_NOTSPD);
        }
    }
    __syncthreads();
}

// gamma.py:124-148: a = phi1, b = -phi0
__device__ __forceinline__ void vb_gamma(double phi0, double phi1, double &u0, double &u1, double &g, int *ctrl) {
    double a = phi1, b = -phi0;
    if (!(a > 0.0) || !(b > 0.0)) atomicOr(&ctrl[2], BPK_FLAG_DOMAIN);
    double lb = log(b);
    u0 = a / b;
    u1 = bpk_digamma(a) - lb;
    g = a * lb - lgamma(a);
}

// sum_mn <(y - f)^2> = sum y^2 - 2 <W>:S_yx + tr(sum_m<ww^T> sum_n<xx^T>)   (dot.py:355,403 summed)
__device__ __forceinline__ double vb_E2(const double *st, const int64_t *o, int M, int K, double *red) {
    const double *W = st + o[F_W], *Syx = st + o[F_STATS], *SWW = st + o[F_SWW], *SXXT = st + o[F_SXXT];
    double a = 0.0, b = 0.0;
    for (int e = threadIdx.x; e < M * K; e += VB_THREADS) a += W[e] * Syx[e];
    for (int e = threadIdx.x; e < K * K; e += VB_THREADS) b += SWW[e] * SXXT[e];
    return st[o[F_SUMSQ]] + vb_block_sum(b - 2.0 * a, red);
}

__global__ void __launch_bounds__(VB_THREADS, 1) pca_vb_small_kernel(PcaVbArgs p) {
    extern __shared__ double sm[];
    const int M = p.M, K = p.K, ld = SPD_LD(K), t = threadIdx.x;
    double *S = sm, *C = S + (size_t)K * ld, *B = C + (size_t)K * ld, *red = B + (size_t)K * 32, *scal = red + 8;
    __shared__ int64_t o[F_COUNT + 1];
    if (t == 0) pca_vb_offsets(M, K, o);
    __syncthreads();
    double *st = p.st;
    volatile int *stop = p.ctrl + 1;
    const double Ng = st[o[F_NG]];

    for (int ip = 0; ip < p.nops; ++ip) {
        __syncthreads();
        if (*stop) return;
        const int op = p.ops[ip];
        if (op == BPK_VBOP_STATS) {
            // fixed-order grid reduction of the sweep kernel's per-CTA partials
            double *dst = st + (p.local_stats ? o[F_STATS_LOCAL] : o[F_STATS]);
            const int total = M * K + K * K + K;
            if (p.partial) {
                for (int e = t; e < total; e += VB_THREADS) {
                    int pe;
                    if (e < M * K) { int m = e / K, k = e - m * K; pe = m * PCA_KP + k; }
                    else if (e < M * K + K * K) { int r = e - M * K; int i = r / K, j = r - i * K; pe = PCA_MP * PCA_KP + i * PCA_KP + j; }
                    else pe = PCA_MP * PCA_KP + PCA_KP * PCA_KP + (e - M * K - K * K);
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    int b = 0;
                    for (; b + 3 < p.nparts; b += 4) {
                        s0 += p.partial[(size_t)b * PCA_NSTAT + pe];
                        s1 += p.partial[(size_t)(b + 1) * PCA_NSTAT + pe];
                        s2 += p.partial[(size_t)(b + 2) * PCA_NSTAT + pe];
                        s3 += p.partial[(size_t)(b + 3) * PCA_NSTAT + pe];
                    }
                    for (; b < p.nparts; ++b) s0 += p.partial[(size_t)b * PCA_NSTAT + pe];
                    dst[e] = (s0 + s1) + (s2 + s3);
                }
            }
        } else if (op == BPK_VBOP_SXXT) {
            // sum_n <x x^T> = N Cov_x + S_xx
            for (int e = t; e < K * K; e += VB_THREADS)
                st[o[F_SXXT] + e] = Ng * st[o[F_COVX] + e] + st[o[F_STATS] + M * K + e];
        } else if (op == BPK_VBOP_XPRE) {
            // q(X) shared part: Lam_x = diag(a_x) + tau sum_m<ww^T>; x_n = Cov_x (a_x mu_x + tau W^T y_n)
            const double tau = st[o[F_TAU_U0]];
            for (int e = t; e < K * K; e += VB_THREADS) {
                int i = e / K, j = e - i * K;
                double lam = tau * st[o[F_SWW] + e] + (i == j ? st[o[F_AX] + i] : 0.0);
                S[i * ld + j] = lam;
                st[o[F_LAMX] + e] = lam;
            }
            vb_spd_inverse(S, B, C, K, ld, scal, p.ctrl);
            for (int e = t; e < K * K; e += VB_THREADS) st[o[F_COVX] + e] = C[(e / K) * ld + (e % K)];
            if (t == 0) st[o[F_LOGDETX]] = scal[0];
            for (int k = t; k < K; k += VB_THREADS) {
                double s = 0.0;
                for (int j = 0; j < K; ++j) s += C[k * ld + j] * (st[o[F_AX] + j] * st[o[F_MUX] + j]);
                st[o[F_BX] + k] = s;
            }
            for (int e = t; e < K * M; e += VB_THREADS) {
                int k = e / M, m = e - k * M;
                double s = 0.0;
                for (int j = 0; j < K; ++j) s += C[k * ld + j] * st[o[F_W] + m * K + j];
                st[o[F_A] + e] = tau * s;
            }
        } else if (op == BPK_VBOP_ROW) {
            // q(C): Lam_c = diag<alpha> + tau sum_n<xx^T> (shared by all rows); phi0_m = <alpha> mu_c + tau S_yx[m]
            const double tau = st[o[F_TAU_U0]];
            for (int e = t; e < K * K; e += VB_THREADS) {
                int i = e / K, j = e - i * K;
                double lam = tau * st[o[F_SXXT] + e] + (i == j ? st[o[F_AL_U0] + i] : 0.0);
                S[i * ld + j] = lam;
                st[o[F_LAMC] + e] = lam;
            }
            vb_spd_inverse(S, B, C, K, ld, scal, p.ctrl);
            for (int e = t; e < K * K; e += VB_THREADS) st[o[F_COVC] + e] = C[(e / K) * ld + (e % K)];
            if (t == 0) st[o[F_LOGDETC]] = scal[0];
            for (int e = t; e < M * K; e += VB_THREADS) {
                int k = e % K;
                st[o[F_PHI0C] + e] = st[o[F_AL_U0] + k] * st[o[F_MUC] + k] + tau * st[o[F_STATS] + e];
            }
            __syncthreads();
            for (int e = t; e < M * K; e += VB_THREADS) {
                int m = e / K, i = e - m * K;
                double s = 0.0;
                for (int j = 0; j < K; ++j) s += C[i * ld + j] * st[o[F_PHI0C] + m * K + j];
                st[o[F_W] + e] = s;
            }
            __syncthreads();
            const double ldc = scal[0];
            for (int m = t; m < M; m += VB_THREADS) {
                double s = 0.0;
                for (int k = 0; k < K; ++k) s += st[o[F_W] + m * K + k] * st[o[F_PHI0C] + m * K + k];
                st[o[F_GC] + m] = -0.5 * s + 0.5 * ldc;
            }
            for (int e = t; e < K * K; e += VB_THREADS) {
                int i = e / K, j = e - i * K;
                double s = 0.0;
                for (int m = 0; m < M; ++m) s += st[o[F_W] + m * K + i] * st[o[F_W] + m * K + j];
                st[o[F_SWW] + e] = (double)M * C[i * ld + j] + s;
            }
        } else if (op == BPK_VBOP_ALPHA) {
            // gaussian.py:609-637 index 1 summed over the M rows, then gamma.py:124-148
            for (int k = t; k < K; k += VB_THREADS) {
                double sw = 0.0;
                for (int m = 0; m < M; ++m) sw += st[o[F_W] + m * K + k];
                double mu = st[o[F_MUC] + k];
                double d = st[o[F_SWW] + k * K + k] - 2.0 * mu * sw + (double)M * mu * mu;
                double phi0 = -st[o[F_B0] + k] - 0.5 * d;
                double phi1 = st[o[F_A0] + k] + 0.5 * (double)M;
                double u0, u1, g;
                vb_gamma(phi0, phi1, u0, u1, g, p.ctrl);
                st[o[F_AL_PHI0] + k] = phi0;
                st[o[F_AL_PHI1] + k] = phi1;
                st[o[F_AL_U0] + k] = u0;
                st[o[F_AL_U1] + k] = u1;
                st[o[F_AL_G] + k] = g;
            }
        } else if (op == BPK_VBOP_TAU) {
            // gaussian.py:2351-2371 summed over (M,N), then gamma.py:124-148
            double E2 = vb_E2(st, o, M, K, red);
            if (t == 0) {
                double phi0 = -st[o[F_TB0]] - 0.5 * E2;
                double phi1 = st[o[F_TA0]] + 0.5 * (double)M * Ng;
                double u0, u1, g;
                vb_gamma(phi0, phi1, u0, u1, g, p.ctrl);
                st[o[F_TAU_PHI0]] = phi0;
                st[o[F_TAU_PHI1]] = phi1;
                st[o[F_TAU_U0]] = u0;
                st[o[F_TAU_U1]] = u1;
                st[o[F_TAU_G]] = g;
            }
        } else if (op == BPK_VBOP_BOUND) {
            // expfamily.py:400-480 for Y, X, C, alpha, tau from the plate-summed statistics
            const double tau = st[o[F_TAU_U0]], logtau = st[o[F_TAU_U1]];
            double E2 = vb_E2(st, o, M, K, red);
            double LY = -0.5 * tau * E2 + 0.5 * (double)M * Ng * (logtau - LOG2PI_D);
            // X
            const double *Sxx = st + o[F_STATS] + M * K, *sx = Sxx + K * K;
            double a = 0.0;   // tr(Lam_x S_xx), (phi1_p - phi1_q):sum<xx^T>
            double b = 0.0;
            for (int e = t; e < K * K; e += VB_THREADS) {
                int i = e / K, j = e - i * K;
                double lam = st[o[F_LAMX] + e];
                a += lam * Sxx[e];
                b += (0.5 * lam - (i == j ? 0.5 * st[o[F_AX] + i] : 0.0)) * st[o[F_SXXT] + e];
            }
            double c = 0.0;   // phi0_p . s_x  and the prior cgf
            for (int k = t; k < K; k += VB_THREADS) {
                double ax = st[o[F_AX] + k], mu = st[o[F_MUX] + k];
                c += ax * mu * sx[k] + Ng * (-0.5 * ax * mu * mu + 0.5 * log(ax));
            }
            double trLS = vb_block_sum(a, red);
            double LX = vb_block_sum(b + c, red) - trLS + 0.5 * trLS - 0.5 * Ng * st[o[F_LOGDETX]];
            // C
            double d = 0.0;
            for (int e = t; e < M * K; e += VB_THREADS) {
                int k = e % K;
                d += (st[o[F_AL_U0] + k] * st[o[F_MUC] + k] - st[o[F_PHI0C] + e]) * st[o[F_W] + e];
            }
            for (int e = t; e < K * K; e += VB_THREADS) {
                int i = e / K, j = e - i * K;
                d += (0.5 * st[o[F_LAMC] + e] - (i == j ? 0.5 * st[o[F_AL_U0] + i] : 0.0)) * st[o[F_SWW] + e];
            }
            for (int m = t; m < M; m += VB_THREADS) d -= st[o[F_GC] + m];
            for (int k = t; k < K; k += VB_THREADS) {
                double mu = st[o[F_MUC] + k];
                d += (double)M * (-0.5 * st[o[F_AL_U0] + k] * mu * mu + 0.5 * st[o[F_AL_U1] + k]);
            }
            double LC = vb_block_sum(d, red);
            // alpha
            double f = 0.0;
            if (p.has_alpha) {
                for (int k = t; k < K; k += VB_THREADS) {
                    double a0 = st[o[F_A0] + k], b0 = st[o[F_B0] + k];
                    f += (-b0 - st[o[F_AL_PHI0] + k]) * st[o[F_AL_U0] + k]
                       + (a0 - st[o[F_AL_PHI1] + k]) * st[o[F_AL_U1] + k]
                       + (a0 * log(b0) - lgamma(a0)) - st[o[F_AL_G] + k];
                }
            }
            double LA = vb_block_sum(f, red);
            if (t == 0) {
                double LT = 0.0;
                if (p.has_tau) {
                    double a0 = st[o[F_TA0]], b0 = st[o[F_TB0]];
                    LT = (-b0 - st[o[F_TAU_PHI0]]) * tau + (a0 - st[o[F_TAU_PHI1]]) * logtau
                       + (a0 * log(b0) - lgamma(a0)) - st[o[F_TAU_G]];
                }
                double L = (((LY + LX) + LC) + LA) + LT;
                int it = p.ctrl[0];
                if (it < p.cap) {
                    double *row = p.Lhist + (size_t)it * 6;
                    row[0] = LY; row[1] = LX; row[2] = LC; row[3] = LA; row[4] = LT; row[5] = L;
                }
                double L0 = st[o[F_LPREV]];
                st[o[F_LPREV]] = L;
                p.ctrl[0] = it + 1;
                // vmp.py:738-747 (tol < 0 or no previous bound: test disabled)
                if (p.tol >= 0.0 && L0 == L0) {
                    double div = 0.5 * (fabs(L0) + fabs(L));
                    if ((L - L0) / div < p.tol) p.ctrl[1] = 1;
                }
                if (p.ctrl[2]) p.ctrl[1] = 1;
                __threadfence();
            }
        }
    }
}

extern "C" int bpk_pca_vb_layout(int M, int K, int64_t *offsets, int *nfields) {
    if (M < 1 || K < 1) return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_layout: bad shape");
    if (offsets) pca_vb_offsets(M, K, offsets);
    if (nfields) *nfields = F_COUNT;
    return BPK_OK;
}
extern "C" const char *bpk_pca_vb_field_name(int i) { return (i >= 0 && i < F_COUNT) ? kFieldNames[i] : nullptr; }

static int g_vb_timers[64];
static int g_vb_ntimers = 0, g_vb_timer_pos = 0;
extern "C" int bpk_pca_vb_set_timers(const int *ids, int n) {
    if (n < 0 || n > 64) return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_set_timers: at most 64 timers");
    for (int i = 0; i < n; ++i) g_vb_timers[i] = ids[i];
    g_vb_ntimers = n;
    g_vb_timer_pos = 0;
    return BPK_OK;
}
extern "C" int bpk_pca_vb_timers_used(void) { return g_vb_timer_pos; }

static int vb_launch_small(PcaVbArgs &a, size_t smem) {
    if (a.nops == 0) return BPK_OK;
    BPK_LAUNCH(pca_vb_small_kernel, 1, VB_THREADS, smem, a);
    a.nops = 0;
    a.partial = nullptr;
    a.nparts = 0;
    return BPK_OK;
}

extern "C" int bpk_pca_vb_run(const double *Y, int64_t M, int64_t N, int K, double *X, double *state,
                              const int *ops, int nops, int niter, int has_alpha, int has_tau, double tol,
                              double *Lhist, int cap, int *ctrl) {
    BPK_REQUIRE_INIT();
    if (M < 1 || K < 1 || K > BPK_MAXDIM || N < 0 || M > (1 << 20))
        return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: bad shape");
    if (nops < 1 || niter < 0) return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: empty program");
    int nranks = 1, rank = 0;
    bpk_comm_size(&nranks, &rank);
    int64_t off[F_COUNT + 1];
    pca_vb_offsets((int)M, K, off);
    const int ld = SPD_LD(K);
    size_t smem = ((size_t)2 * K * ld + (size_t)K * 32 + 16) * sizeof(double);
    if (smem > (48u << 10))
        BPK_CUDA(cudaFuncSetAttribute(pca_vb_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool fast = (M <= PCA_MP && K <= PCA_KP);
    PcaVbArgs a;
    a.M = (int)M; a.K = K; a.has_alpha = has_alpha; a.has_tau = has_tau; a.tol = tol;
    a.st = state; a.partial = nullptr; a.nparts = 0; a.local_stats = nranks > 1;
    a.Lhist = Lhist; a.cap = cap; a.ctrl = ctrl; a.nops = 0;
    const int64_t nstat = M * K + (int64_t)K * K + K;
    double *stats_dst = state + (nranks > 1 ? off[F_STATS_LOCAL] : off[F_STATS]);
    for (int it = 0; it < niter; ++it) {
        for (int i = 0; i < nops; ++i) {
            const int op = ops[i];
            if (op == BPK_VBOP_XSWEEP) {
                int rc = vb_launch_small(a, smem);
                if (rc) return rc;
                if (N == 0) {
                    BPK_CUDA(cudaMemsetAsync(stats_dst, 0, nstat * sizeof(double), g_bpk.stream));
                    continue;
                }
                int tid = -1;
                if (g_vb_timer_pos < g_vb_ntimers) tid = g_vb_timers[g_vb_timer_pos++];
                if (tid >= 0) bpk_timer_record(tid, 0);
                if (fast) {
                    double *partial = nullptr;
                    int nparts = 0;
                    rc = pca_xsweep_partials(Y, M, N, K, state + off[F_A], state + off[F_BX], X, ctrl + 1,
                                             &partial, &nparts);
                    if (rc) return rc;
                    a.partial = partial;
                    a.nparts = nparts;
                } else {
                    // generic shapes: host decides per iteration (the caller uses niter == 1)
                    BPK_CUDA(cudaMemsetAsync(stats_dst, 0, nstat * sizeof(double), g_bpk.stream));
                    rc = bpk_pca_xsweep(Y, M, N, K, state + off[F_A], state + off[F_BX], X, stats_dst);
                    if (rc) return rc;
                }
                if (tid >= 0) bpk_timer_record(tid, 1);
            } else if (op == BPK_VBOP_STATS) {
                if (fast && N > 0) a.ops[a.nops++] = op;
                if (nranks > 1) {
                    int rc = vb_launch_small(a, smem);
                    if (rc) return rc;
                    rc = bpk_allreduce_sum_f64_oop(state + off[F_STATS_LOCAL], state + off[F_STATS], (uint64_t)nstat);
                    if (rc) return rc;
                }
            } else if (op >= BPK_VBOP_SXXT && op <= BPK_VBOP_BOUND) {
                a.ops[a.nops++] = op;
            } else {
                return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: unknown opcode %d", op);
            }
            if (a.nops == VB_MAXOPS) {
                int rc = vb_launch_small(a, smem);
                if (rc) return rc;
            }
        }
    }
    return vb_launch_small(a, smem);
}
