// pca_vb_ops.cuh — state layout, opcodes and the single-CTA "small ops" of the device-resident PCA VB loop.
// Included by pca_vb.cu (standalone small kernel, driver) and pca.cu (tail of the fused sweep kernel).
#pragma once
#include "common.cuh"
#include "pca_common.cuh"
#include "spd.cuh"

#define VB_THREADS 256      // block size of the standalone small kernel
#define VB_MAXOPS 24
#define LOG2PI_D 1.8378770664093454835606594728112

enum {
    F_MUX = 0, F_AX, F_MUC, F_A0, F_B0, F_TA0, F_TB0, F_SUMSQ, F_NG, F_RESERVED,
    F_W, F_SWW, F_COVC, F_LAMC, F_LOGDETC, F_PHI0C, F_GC,
    F_AL_PHI0, F_AL_PHI1, F_AL_U0, F_AL_U1, F_AL_G,
    F_TAU_PHI0, F_TAU_PHI1, F_TAU_U0, F_TAU_U1, F_TAU_G,
    F_COVX, F_LAMX, F_LOGDETX, F_A, F_BX,
    F_STATS, F_SXXT, F_LPREV, F_STATS_LOCAL, F_PHI1X, F_PHI1C,
    F_COUNT
};

__attribute__((unused)) static const char *const kFieldNames[F_COUNT] = {
    "mux", "ax", "muc", "a0", "b0", "ta0", "tb0", "sumsq", "ng", "reserved",
    "w", "sww", "covc", "lamc", "logdetc", "phi0c", "gc",
    "al_phi0", "al_phi1", "al_u0", "al_u1", "al_g",
    "tau_phi0", "tau_phi1", "tau_u0", "tau_u1", "tau_g",
    "covx", "lamx", "logdetx", "A", "bx",
    "stats", "sxxt", "lprev", "stats_local", "phi1x", "phi1c",
};

__device__ __forceinline__ void vb_stamp(unsigned long long *dbg, int slot) {
    if (dbg && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg[slot] = t;
    }
}

__device__ __forceinline__ void vb_stamp_lane(unsigned long long *dbg, int slot) {       // lane 0 of the calling warp
    if (dbg && (threadIdx.x & 31) == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        dbg[slot] = t;
    }
}

// ---- LL packets of the peer-memory exchange (layout: common.cuh) ---------------------------------------
// slot of element e deposited by rank r with parity par, inside the window `win`
__device__ __forceinline__ unsigned long long *ll_slot(double *win, int par, int r, int e) {
    return (unsigned long long *)win + BPK_XCHG_DATA + ((size_t)(par * BPK_XCHG_MAXRANKS + r) * BPK_XCHG_CAP + e) * 2;
}
// one fp64 value as two self-tagged 8-byte packets (each store is atomic; no fence needed around it)
__device__ __forceinline__ void ll_store(unsigned long long *slot, double v, unsigned int seq) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long w0 = ((unsigned long long)seq << 32) | (b & 0xffffffffull);
    const unsigned long long w1 = ((unsigned long long)seq << 32) | (b >> 32);
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};\n" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ll_load(const unsigned long long *slot, unsigned long long &w0, unsigned long long &w1) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
}
// slot of the rank-summed element e (written by the local CTA that owns e, read by the CTA that runs the small ops)
__device__ __forceinline__ unsigned long long *ll_total_slot(double *win, int par, int e) {
    return (unsigned long long *)win + BPK_XCHG_TOTALS + ((size_t)par * BPK_XCHG_CAP + e) * 2;
}
// spin until both packets of the slot carry this exchange's tag; ~20 s watchdog -> error bit 4 (a peer died; ranks
// may start seconds apart on a busy host)
__device__ __forceinline__ double ll_wait(const unsigned long long *slot, unsigned int seq, int *errword) {
    unsigned long long w0, w1;
    long long t0 = 0;
    bool timing = false;
    for (;;) {
        ll_load(slot, w0, w1);
        if ((unsigned int)(w0 >> 32) == seq && (unsigned int)(w1 >> 32) == seq) break;
        if (!timing) { t0 = clock64(); timing = true; }
        else if (clock64() - t0 > 40000000000ll) { atomicOr(errword, 4); return 0.0; }
    }
    return __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
}

__host__ __device__ inline void pca_vb_offsets(int M, int K, int64_t *off /* [F_COUNT+1] */) {
    const int64_t KK = (int64_t)K * K, MK = (int64_t)M * K, NS = MK + KK + K;
    const int64_t size[F_COUNT] = {
        K, K, K, K, K, 1, 1, 1, 1, 1,
        MK, KK, KK, KK, 1, MK, M,
        K, K, K, K, K,
        1, 1, 1, 1, 1,
        KK, KK, 1, MK, K,
        NS, KK, 1, NS, KK, KK,
    };
    int64_t o = 0;
    for (int f = 0; f < F_COUNT; ++f) { off[f] = o; o += size[f]; }
    off[F_COUNT] = o;
}

// scratch of the small ops in doubles: augmented K x 2K Gauss-Jordan tile (odd pitch), two pivot rows / columns
// (the elimination takes two pivots per step), pivots, reduction slots; then a copy of Lam_x and a second set of
// reduction slots for the op that runs side by side with another one (BOUND next to XPRE, TAU next to ALPHA), and a
// backup of what XPRE overwrites (restored if the bound it ran next to raised the stop word; M <= 64 only)
__host__ __device__ inline size_t pca_vb_smem_doubles(int K) {
    return (size_t)K * (2 * K + 1) + 7 * (size_t)K + 64 + (size_t)K * K + 256 + ((size_t)PCA_MP * K + 3 * (size_t)K * K + K + 1);
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
    int *ctrl;               // [0] iterations finished, [1] stop, [2] error bits (1 not SPD, 2 domain, 4 exchange timeout)
    unsigned long long *dbg; // optional %globaltimer stamps (tools/vb_tail_timing.py), NULL in production
    int xranks, xrank;       // > 1: STATS all-reduces over the peer-memory windows below (no NCCL call)
    double *xwin[BPK_XCHG_MAXRANKS];
    double *xown;            // = xwin[xrank]
    int ll;                  // 1 (fused sweep kernel): every CTA has pushed its slice of the reduced statistics into the
                             // windows as LL packets (xranks may be 1: the window is then this GPU's own); STATS gathers
    int gj2;                 // 1: K x K inverses eliminate two pivots per step (half the barriers)
    int par;                 // 1: independent neighbours in the op list run side by side on two halves of the CTA
    int niter;               // fused sweep kernel: sweeps per launch; ops[0..nops) follow every sweep but the last,
    int derive_sxx;          // 1: the sweep kernel did not accumulate S_xx; STATS forms it as A S_yx + b s_x^T
    int dry_every;           // 1: CTA 0 dry-runs the tail in every sweep of the launch, 0: only in the first
    int nops_last;           // ops[0..nops_last) the last one (the ops of the next sweep's head are dropped)
    int nops;
    int ops[VB_MAXOPS];
};

__device__ __forceinline__ double vb_block_sum(double v, double *red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

// a TEAM: n threads (a multiple of 32; thread index t within it) that meet at named barrier `bar` (0 = the whole CTA)
struct VbTeam { int t, n, bar; };
__device__ __forceinline__ void vb_team_sync(const VbTeam &T) {
    if (T.bar == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;\n" ::"r"(T.bar), "r"(T.n) : "memory");
}
__device__ __forceinline__ double vb_team_sum(double v, double *red, const VbTeam &T) {
    v = warp_sum(v);
    vb_team_sync(T);
    if ((T.t & 31) == 0) red[T.t >> 5] = v;
    vb_team_sync(T);
    double s = 0.0;
    const int nw = T.n >> 5;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
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
__device__ __forceinline__ double vb_E2(const double *st, const int64_t *o, int M, int K, double *red, const VbTeam &T) {
    const double *W = st + o[F_W], *Syx = st + o[F_STATS], *SWW = st + o[F_SWW], *SXXT = st + o[F_SXXT];
    double a = 0.0, b = 0.0;
    for (int e = T.t; e < M * K; e += T.n) a += W[e] * Syx[e];
    for (int e = T.t; e < K * K; e += T.n) b += SWW[e] * SXXT[e];
    return st[o[F_SUMSQ]] + vb_team_sum(b - 2.0 * a, red, T);
}

// One run of small ops by ONE CTA (any multiple of 32 threads up to 1024).  sm: shared-memory scratch of
// sm_doubles doubles, at least pca_vb_smem_doubles(K); when the whole state vector fits behind it, the state
// is staged in shared memory for the duration of the ops (every dependent step then costs a shared-memory
// round trip instead of an L2 one) and written back at the end.  KC/MC > 0 fix the shape at compile time
// (index arithmetic becomes shifts; this code runs once per launch from a cold instruction cache, so small
// and fast matters more than general).
template <int KC, int MC>
// rmode 1 (service CTA of the persistent loop kernel): the state vector is RESIDENT in shared memory for the whole launch
// (pca_vb_state_load / pca_vb_state_store bracket the launch): nothing is staged in or out here; only what the rest of
// the grid reads — A and b of the next sweep — is published to global memory.
static __device__ __noinline__ void pca_vb_ops_t(const PcaVbArgs &p, double *sm, size_t sm_doubles, bool dry, int nops_run,
                                                  unsigned long long xseq, int rmode) {
    const int VBT = blockDim.x;
    const int M = MC ? MC : p.M, K = KC ? KC : p.K, K2 = 2 * K, ldg = K2 + 1, t = threadIdx.x;
    double *G = sm, *rowk = G + (size_t)K * ldg, *colk = rowk + 2 * K2, *piv = colk + 2 * K, *red = piv + K, *scal = red + 32;
    double *lamx_old = scal + 32, *red2 = lamx_old + (size_t)K * K, *xbak = red2 + 256;
    __shared__ int64_t o[F_COUNT + 1];
    __shared__ int dry_ctrl[4];
    if (t == 0) pca_vb_offsets(M, K, o);
    if (t < 4) dry_ctrl[t] = 0;
    __syncthreads();
    double *st = p.st;
    const int64_t nstate = o[F_COUNT];
    const bool staged = pca_vb_smem_doubles(K) + (size_t)nstate <= sm_doubles;
    // dry run: same instruction stream on a throw-away copy of the state, no global side effects —
    // used by the fused sweep kernel to pull this (once-per-launch, otherwise cold) code into the
    // instruction cache while the rest of the grid is still streaming data
    if (dry && !staged) return;
    int *ctrl = dry ? dry_ctrl : p.ctrl;
    const int xranks = dry ? 1 : p.xranks;
    const int cap = dry ? 0 : p.cap;
    unsigned long long *dbg = dry ? nullptr : p.dbg;
    if (staged) {
        st = sm + pca_vb_smem_doubles(K);
        if (rmode == 0) {
            for (int64_t e = t; e < nstate; e += VBT) st[e] = __ldcg(p.st + e);
            __syncthreads();
        } else if (dry) {
            // resident state: the warm-up run works on a scratch copy behind it
            if (pca_vb_smem_doubles(K) + 2 * (size_t)nstate > sm_doubles) return;
            double *st2 = st + nstate;
            for (int64_t e = t; e < nstate; e += VBT) st2[e] = st[e];
            st = st2;
            __syncthreads();
        }
    }
    volatile int *stop = ctrl + 1;
    const double Ng = st[o[F_NG]];
#define CINV(i, j) G[(i) * ldg + K + (j)]


    // ---- ops that may run on a TEAM (half of the CTA) next to an independent neighbour ----------------------------
    const VbTeam full{t, VBT, 0};
    auto op_xpre = [&](const VbTeam &T) {
        // q(X) shared part: Lam_x = diag(a_x) + tau sum_m<ww^T>; x_n = Cov_x (a_x mu_x + tau W^T y_n)
        const double tau = st[o[F_TAU_U0]];
        for (int e = T.t; e < K * K2; e += T.n) {
            const int i = e / K2, j = e - i * K2;
            double v;
            if (j < K) {
                v = tau * st[o[F_SWW] + i * K + j] + (i == j ? st[o[F_AX] + i] : 0.0);
                st[o[F_LAMX] + i * K + j] = v;
                st[o[F_PHI1X] + i * K + j] = -0.5 * v;          // natural parameter phi_1 of q(X) (shared by all columns)
            } else v = (j - K == i) ? 1.0 : 0.0;
            G[i * ldg + j] = v;
        }
        if (T.bar == 0) spd_cta_inverse_gj<KC>(G, rowk, colk, piv, K, scal, &ctrl[2], p.gj2);
        else spd_cta_inverse_gj<KC>(G, rowk, colk, piv, K, scal, &ctrl[2], p.gj2, T.t, T.n, T.bar);
        for (int e = T.t; e < K * K; e += T.n) st[o[F_COVX] + e] = CINV(e / K, e % K);
        if (T.t == 0) st[o[F_LOGDETX]] = scal[0];
        for (int k = T.t; k < K; k += T.n) {
            double s = 0.0;
            for (int j = 0; j < K; ++j) s += CINV(k, j) * (st[o[F_AX] + j] * st[o[F_MUX] + j]);
            st[o[F_BX] + k] = s;
        }
        for (int e = T.t; e < K * M; e += T.n) {
            const int k = e / M, m = e - k * M;
            double s = 0.0;
#pragma unroll 4
            for (int j = 0; j < K; ++j) s += CINV(k, j) * st[o[F_W] + m * K + j];
            st[o[F_A] + e] = tau * s;
        }
    };
    auto op_alpha = [&](const VbTeam &T) {
        // gaussian.py:609-637 index 1 summed over the M rows, then gamma.py:124-148
        for (int k = T.t; k < K; k += T.n) {
            double sw = 0.0;
            for (int m = 0; m < M; ++m) sw += st[o[F_W] + m * K + k];
            double mu = st[o[F_MUC] + k];
            double d = st[o[F_SWW] + k * K + k] - 2.0 * mu * sw + (double)M * mu * mu;
            double phi0 = -st[o[F_B0] + k] - 0.5 * d;
            double phi1 = st[o[F_A0] + k] + 0.5 * (double)M;
            double u0, u1, g;
            vb_gamma(phi0, phi1, u0, u1, g, ctrl);
            st[o[F_AL_PHI0] + k] = phi0;
            st[o[F_AL_PHI1] + k] = phi1;
            st[o[F_AL_U0] + k] = u0;
            st[o[F_AL_U1] + k] = u1;
            st[o[F_AL_G] + k] = g;
        }
    };
    auto op_tau = [&](const VbTeam &T, double *redp) {
        // gaussian.py:2351-2371 summed over (M,N), then gamma.py:124-148
        double E2 = vb_E2(st, o, M, K, redp, T);
        if (T.t == 0) {
            double phi0 = -st[o[F_TB0]] - 0.5 * E2;
            double phi1 = st[o[F_TA0]] + 0.5 * (double)M * Ng;
            double u0, u1, g;
            vb_gamma(phi0, phi1, u0, u1, g, ctrl);
            st[o[F_TAU_PHI0]] = phi0;
            st[o[F_TAU_PHI1]] = phi1;
            st[o[F_TAU_U0]] = u0;
            st[o[F_TAU_U1]] = u1;
            st[o[F_TAU_G]] = g;
        }
    };
    auto op_bound = [&](const VbTeam &T, double *redp, const double *lamx, const double logdetx) {
        // expfamily.py:400-480 for Y, X, C, alpha, tau from the plate-summed statistics.  Every thread accumulates its
        // share of the six sums the bound needs; ONE team reduction (two barriers) combines them.
        const double tau = st[o[F_TAU_U0]], logtau = st[o[F_TAU_U1]];
        const double *W = st + o[F_W], *Syx = st + o[F_STATS], *SWW = st + o[F_SWW], *SXXT = st + o[F_SXXT];
        const double *Sxx = st + o[F_STATS] + M * K, *sx = Sxx + K * K;
        double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};   // [0] E2 cross terms, [1] tr(Lam_x S_xx), [2] X prior terms, [3] C, [4] alpha
        for (int e = T.t; e < M * K; e += T.n) {
            const int k = e % K;
            v[0] -= 2.0 * W[e] * Syx[e];
            v[3] += (st[o[F_AL_U0] + k] * st[o[F_MUC] + k] - st[o[F_PHI0C] + e]) * W[e];
        }
        for (int e = T.t; e < K * K; e += T.n) {
            const int i = e / K, j = e - i * K;
            const double lam = lamx[e];
            v[0] += SWW[e] * SXXT[e];
            v[1] += lam * Sxx[e];
            v[2] += (0.5 * lam - (i == j ? 0.5 * st[o[F_AX] + i] : 0.0)) * SXXT[e];
            v[3] += (0.5 * st[o[F_LAMC] + e] - (i == j ? 0.5 * st[o[F_AL_U0] + i] : 0.0)) * SWW[e];
        }
        for (int k = T.t; k < K; k += T.n) {
            const double ax = st[o[F_AX] + k], mux = st[o[F_MUX] + k], muc = st[o[F_MUC] + k];
            v[2] += ax * mux * sx[k] + Ng * (-0.5 * ax * mux * mux + 0.5 * log(ax));
            v[3] += (double)M * (-0.5 * st[o[F_AL_U0] + k] * muc * muc + 0.5 * st[o[F_AL_U1] + k]);
            if (p.has_alpha) {
                const double a0 = st[o[F_A0] + k], b0 = st[o[F_B0] + k];
                v[4] += (-b0 - st[o[F_AL_PHI0] + k]) * st[o[F_AL_U0] + k]
                      + (a0 - st[o[F_AL_PHI1] + k]) * st[o[F_AL_U1] + k]
                      + (a0 * log(b0) - lgamma(a0)) - st[o[F_AL_G] + k];
            }
        }
        for (int m = T.t; m < M; m += T.n) v[3] -= st[o[F_GC] + m];
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] = warp_sum(v[q]);
        vb_team_sync(T);
        if ((T.t & 31) == 0) {
#pragma unroll
            for (int q = 0; q < 5; ++q) redp[(T.t >> 5) * 8 + q] = v[q];
        }
        vb_team_sync(T);
        if (T.t == 0) {
            double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            const int nw = T.n >> 5;
            for (int w = 0; w < nw; ++w)
#pragma unroll
                for (int q = 0; q < 5; ++q) s[q] += redp[w * 8 + q];
            const double E2 = st[o[F_SUMSQ]] + s[0];
            const double LY = -0.5 * tau * E2 + 0.5 * (double)M * Ng * (logtau - LOG2PI_D);
            const double LX = s[2] - s[1] + 0.5 * s[1] - 0.5 * Ng * logdetx;
            const double LC = s[3], LA = s[4];
            double LT = 0.0;
            if (p.has_tau) {
                double a0 = st[o[F_TA0]], b0 = st[o[F_TB0]];
                LT = (-b0 - st[o[F_TAU_PHI0]]) * tau + (a0 - st[o[F_TAU_PHI1]]) * logtau
                   + (a0 * log(b0) - lgamma(a0)) - st[o[F_TAU_G]];
            }
            double L = (((LY + LX) + LC) + LA) + LT;
            int it = ctrl[0];
            if (it < cap) {
                double *row = p.Lhist + (size_t)it * 6;
                row[0] = LY; row[1] = LX; row[2] = LC; row[3] = LA; row[4] = LT; row[5] = L;
            }
            double L0 = st[o[F_LPREV]];
            st[o[F_LPREV]] = L;
            ctrl[0] = it + 1;
            // vmp.py:738-747 (tol < 0 or no previous bound: test disabled)
            if (p.tol >= 0.0 && L0 == L0) {
                double div = 0.5 * (fabs(L0) + fabs(L));
                if ((L - L0) / div < p.tol) ctrl[1] = 1;
            }
            if (ctrl[2]) ctrl[1] = 1;
            __threadfence();
        }
    };
    // side-by-side execution needs whole warps on both sides and enough threads for the two-pivot elimination
    const bool can_pair = p.par && (VBT % 64) == 0 && VBT / 2 >= 2 * K2 + 2 * K + 1;

    for (int ip = 0; ip < nops_run; ++ip) {
        __syncthreads();
        if (*stop) break;
        const int op = p.ops[ip];
        vb_stamp(dbg, 8 + ip);
        if (can_pair && ip + 1 < nops_run) {
            const int op2 = p.ops[ip + 1];
            if ((op == BPK_VBOP_ALPHA && op2 == BPK_VBOP_TAU) || (op == BPK_VBOP_TAU && op2 == BPK_VBOP_ALPHA)) {
                // alpha.update() and tau.update() read <W>, sum<ww^T>, the statistics; neither reads what the other writes
                if (t < 32) op_alpha(VbTeam{t, 32, 1});
                else op_tau(VbTeam{t - 32, VBT - 32, 2}, red2);
                vb_stamp(dbg, 8 + ip + 1);
                ++ip;
                continue;
            }
            if (op == BPK_VBOP_BOUND && op2 == BPK_VBOP_XPRE && M <= PCA_MP) {
                // the bound of this sweep next to the shared part of the NEXT q(X): XPRE overwrites Lam_x and log det Lam_x,
                // which the bound still needs from the q(X) that is in force — it reads them from a copy.  If the bound
                // raises the stop word (converged: the reference would not have touched X again), XPRE is undone.
                const int64_t nblk = o[F_BX + 1] - o[F_COVX];          // Cov_x, Lam_x, log det, A, b: one contiguous block
                for (int e = t; e < K * K; e += VBT) { lamx_old[e] = st[o[F_LAMX] + e]; xbak[nblk + e] = st[o[F_PHI1X] + e]; }
                for (int64_t e = t; e < nblk; e += VBT) xbak[e] = st[o[F_COVX] + e];
                const double logdetx_old = st[o[F_LOGDETX]];
                __syncthreads();
                const int h = VBT / 2;
                if (t < h) op_xpre(VbTeam{t, h, 1});
                else op_bound(VbTeam{t - h, h, 2}, red2, lamx_old, logdetx_old);
                __syncthreads();
                if (*stop) {
                    for (int e = t; e < K * K; e += VBT) st[o[F_PHI1X] + e] = xbak[nblk + e];
                    for (int64_t e = t; e < nblk; e += VBT) st[o[F_COVX] + e] = xbak[e];
                }
                vb_stamp(dbg, 8 + ip + 1);
                ++ip;
                continue;
            }
        }
        if (op == BPK_VBOP_STATS) {
            // fixed-order grid reduction of the sweep kernel's per-CTA partials + the sweep's ONE exchange
            const int total = M * K + K * K + K;
            const bool llx = !dry && p.ll;                   // fused kernel: the grid has already pushed its slices
            const bool xch = !dry && !p.ll && xranks > 1;    // single-CTA launch: push and gather here
            double *dst = st + ((p.local_stats || xch) ? o[F_STATS_LOCAL] : o[F_STATS]);
            auto padded = [&](int e) -> int {
                if (e < M * K) { const int m = e / K; return m * PCA_KP + (e - m * K); }
                if (e < M * K + K * K) { const int r = e - M * K, i = r / K; return PCA_MP * PCA_KP + i * PCA_KP + (r - i * K); }
                return PCA_MP * PCA_KP + PCA_KP * PCA_KP + (e - M * K - K * K);
            };
            if (!llx && p.partial) {
                for (int e = t; e < total; e += VBT) {
                    const int pe = padded(e);
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
            if (llx || xch) {
                // The exchange, inside this kernel: every rank's deposit lands in MY window as self-tagged LL packets
                // (pushed over NVLink by the peers' CTAs, by this GPU's own CTAs for its own share) and is summed in
                // rank order — bit-identical on every rank — as soon as the tags show this exchange's number: by the
                // CTA that owns the element (fused sweep kernel; this CTA then only collects the totals) or here.
                // Two parities of slots: a rank cannot get more than one exchange ahead of a peer, because it needs
                // that peer's deposit to finish the one in between.
                __syncthreads();
                vb_stamp(dbg, 40);
                double *own = p.xown;
                const unsigned long long seq64 = llx ? xseq : *(volatile unsigned long long *)own + 1ull;
                const unsigned int seq = (unsigned int)seq64;
                const int par = (int)(seq & 1u);
                if (xch) {
                    for (int e = t; e < total; e += VBT) {
                        const int pe = padded(e);
                        const double v = dst[e];
                        for (int r = 0; r < xranks; ++r) ll_store(ll_slot(p.xwin[r], par, p.xrank, pe), v, seq);
                    }
                }
                vb_stamp(dbg, 41);
                for (int e = t; e < total; e += VBT) {
                    if (p.derive_sxx && e >= M * K && e < M * K + K * K) continue;      // formed below
                    const int pe = padded(e);
                    double s;
                    if (llx) s = ll_wait(ll_total_slot(own, par, pe), seq, &ctrl[2]);    // the owner CTA summed over ranks
                    else {
                        s = 0.0;
                        for (int r = 0; r < xranks; ++r) s += ll_wait(ll_slot(own, par, r, pe), seq, &ctrl[2]);
                    }
                    st[o[F_STATS] + e] = s;
                }
                __syncthreads();
                if (t == 0) *(volatile unsigned long long *)own = seq64;
                vb_stamp(dbg, 42);
            }
            if (p.derive_sxx) {
                // x_n = A y_n + b  =>  sum_n x_n x_n^T = A (sum_n y_n x_n^T) + b (sum_n x_n)^T, with the A, b this sweep used.
                // Two threads per output element (halves of the m range), combined by one shuffle.
                __syncthreads();
                const double *Sg = st + o[F_STATS];
                const int half = t & 1, mh = (M + 1) >> 1;
                const int m0 = half ? mh : 0, m1 = half ? M : mh;
                for (int eb = 0; eb < K * K; eb += VBT >> 1) {
                    const int e = eb + (t >> 1);
                    const bool act = e < K * K;
                    const int i = act ? e / K : 0, j = act ? e - i * K : 0;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                    int m = m0;
                    for (; m + 3 < m1; m += 4) {
                        s0 += st[o[F_A] + i * M + m] * Sg[m * K + j];
                        s1 += st[o[F_A] + i * M + m + 1] * Sg[(m + 1) * K + j];
                        s2 += st[o[F_A] + i * M + m + 2] * Sg[(m + 2) * K + j];
                        s3 += st[o[F_A] + i * M + m + 3] * Sg[(m + 3) * K + j];
                    }
                    for (; m < m1; ++m) s0 += st[o[F_A] + i * M + m] * Sg[m * K + j];
                    double s = (s0 + s1) + (s2 + s3);
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    if (act && half == 0) G[i * K + j] = s + st[o[F_BX] + i] * Sg[M * K + K * K + j];
                }
                __syncthreads();
                for (int e = t; e < K * K; e += VBT) {
                    const int i = e / K, j = e - i * K;
                    st[o[F_STATS] + M * K + e] = 0.5 * (G[i * K + j] + G[j * K + i]);
                }
            }
        } else if (op == BPK_VBOP_SXXT) {
            // sum_n <x x^T> = N Cov_x + S_xx
            for (int e = t; e < K * K; e += VBT)
                st[o[F_SXXT] + e] = Ng * st[o[F_COVX] + e] + st[o[F_STATS] + M * K + e];
        } else if (op == BPK_VBOP_XPRE) {
            op_xpre(full);
        } else if (op == BPK_VBOP_ROW) {
            // q(C): Lam_c = diag<alpha> + tau sum_n<xx^T> (shared by all rows); phi0_m = <alpha> mu_c + tau S_yx[m]
            const double tau = st[o[F_TAU_U0]];
            for (int e = t; e < K * K2; e += VBT) {
                const int i = e / K2, j = e - i * K2;
                double v;
                if (j < K) {
                    v = tau * st[o[F_SXXT] + i * K + j] + (i == j ? st[o[F_AL_U0] + i] : 0.0);
                    st[o[F_LAMC] + i * K + j] = v;
                    st[o[F_PHI1C] + i * K + j] = -0.5 * v;
                } else v = (j - K == i) ? 1.0 : 0.0;
                G[i * ldg + j] = v;
            }
            for (int e = t; e < M * K; e += VBT) {
                const int k = e % K;
                st[o[F_PHI0C] + e] = st[o[F_AL_U0] + k] * st[o[F_MUC] + k] + tau * st[o[F_STATS] + e];
            }
            spd_cta_inverse_gj<KC>(G, rowk, colk, piv, K, scal, &ctrl[2], p.gj2);
            for (int e = t; e < K * K; e += VBT) st[o[F_COVC] + e] = CINV(e / K, e % K);
            const double ldc = scal[0];
            if (t == 0) st[o[F_LOGDETC]] = ldc;
            for (int e = t; e < M * K; e += VBT) {
                const int m = e / K, i = e - m * K;
                double s = 0.0;
#pragma unroll 4
                for (int j = 0; j < K; ++j) s += CINV(i, j) * st[o[F_PHI0C] + m * K + j];
                st[o[F_W] + e] = s;
            }
            __syncthreads();
            for (int m = t; m < M; m += VBT) {
                double s = 0.0;
#pragma unroll 4
                for (int k = 0; k < K; ++k) s += st[o[F_W] + m * K + k] * st[o[F_PHI0C] + m * K + k];
                st[o[F_GC] + m] = -0.5 * s + 0.5 * ldc;
            }
            for (int e = t; e < K * K; e += VBT) {
                const int i = e / K, j = e - i * K;
                double s0 = 0.0, s1 = 0.0;
                int m = 0;
                for (; m + 1 < M; m += 2) {
                    s0 += st[o[F_W] + m * K + i] * st[o[F_W] + m * K + j];
                    s1 += st[o[F_W] + (m + 1) * K + i] * st[o[F_W] + (m + 1) * K + j];
                }
                if (m < M) s0 += st[o[F_W] + m * K + i] * st[o[F_W] + m * K + j];
                st[o[F_SWW] + e] = (double)M * CINV(i, j) + (s0 + s1);
            }
        } else if (op == BPK_VBOP_ALPHA) {
            op_alpha(full);
        } else if (op == BPK_VBOP_TAU) {
            op_tau(full, red);
        } else if (op == BPK_VBOP_BOUND) {
            op_bound(full, red2, st + o[F_LAMX], st[o[F_LOGDETX]]);
        }
    }
    if (staged && !dry) {
        __syncthreads();
        if (rmode == 0) {
            for (int64_t e = o[F_W] + t; e < nstate; e += VBT) p.st[e] = st[e];      // hyper-parameters are read-only
        } else {
            for (int64_t e = o[F_A] + t; e < o[F_BX + 1]; e += VBT) p.st[e] = st[e]; // A, b: all the other CTAs need
        }
    }
#undef CINV
}

// state vector <-> shared memory, once per launch (service CTA of the persistent loop kernel; requires that the state
// fits behind the scratch, which it does whenever the fast sweep kernel applies: M <= 64, K <= 16)
static __device__ __forceinline__ void pca_vb_state_load(const PcaVbArgs &p, double *sm) {
    int64_t o[F_COUNT + 1];
    pca_vb_offsets(p.M, p.K, o);
    double *st = sm + pca_vb_smem_doubles(p.K);
    for (int64_t e = threadIdx.x; e < o[F_COUNT]; e += blockDim.x) st[e] = __ldcg(p.st + e);
    __syncthreads();
}
static __device__ __forceinline__ void pca_vb_state_store(const PcaVbArgs &p, double *sm) {
    int64_t o[F_COUNT + 1];
    pca_vb_offsets(p.M, p.K, o);
    const double *st = sm + pca_vb_smem_doubles(p.K);
    __syncthreads();
    for (int64_t e = o[F_W] + threadIdx.x; e < o[F_COUNT]; e += blockDim.x) p.st[e] = st[e];
}

// xseq: sequence number of this sweep's exchange when the caller (fused sweep kernel) has already pushed the grid's
// slices (p.ll); ignored otherwise.
static __device__ __forceinline__ void pca_vb_ops(const PcaVbArgs &p, double *sm, size_t sm_doubles, bool dry = false,
                                                  int nops_run = -1, unsigned long long xseq = 0ull, int rmode = 0) {
    if (nops_run < 0) nops_run = p.nops;
    if (p.M == PCA_MP && p.K == PCA_KP) pca_vb_ops_t<PCA_KP, PCA_MP>(p, sm, sm_doubles, dry, nops_run, xseq, rmode);
    else pca_vb_ops_t<0, 0>(p, sm, sm_doubles, dry, nops_run, xseq, rmode);
}


// pca.cu: the one-pass sweep (requires M<=64, K<=16).  Without `tail` the per-CTA partial statistics
// are left in scratch for the caller to reduce; with `tail` (aligned inputs) the launch also reduces
// them and runs the given small ops in its last CTA (*tail_done = 1), i.e. one launch per VB sweep.
int pca_xsweep_partials(const double *Y, int64_t M, int64_t N, int K, const double *A, const double *b,
                        double *X, const int *stop, double **partial_out, int *nparts_out,
                        const PcaVbArgs *tail, int *tail_done);
bool pca_ws_available(const double *Y, int64_t N);
