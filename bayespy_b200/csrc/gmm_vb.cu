// gmm_vb.cu — the device-resident VB loop of the Gaussian mixture model (gmm.rst:71-98)
//
//   alpha ~ Dirichlet(a0)                       plates ()      dims (K,)
//   Z     ~ Categorical(alpha)                  plates (N,)
//   mu    ~ Gaussian(m0, L0)                    plates (K,)    dims (D,)
//   Lam   ~ Wishart(n0, V0)                     plates (K,)    dims (D,D)
//   Y     ~ Mixture(Z, Gaussian, mu, Lam)       plates (N,), fully observed
//
// What it replaces: the Python scheduler loop of VB.update (vmp.py:132-172) with the per-node update and
// lower-bound calls it makes for this model — mixture.py:108-225 (messages to mu / Lambda / Z from the
// plate-summed statistics), gaussian.py:397-463, wishart.py:149-205, dirichlet.py:120-170,
// categorical.py / multinomial.py:83-130, expfamily.py:400-480 (the bound of every node) and the convergence
// test of vmp.py:717-747.  Driven node by node from Python, one sweep of this model is ~90 tiny launches around
// the 5.9 ms responsibilities kernel (10.9 ms per sweep at N = 1e7, D = 8, K = 64).  Here one sweep is
//       [gmm_vb_small_kernel: ..., ZPRE] -> [gmm sweep kernel] -> [partials -> statistics] (-> all-reduce)
//       -> [gmm_vb_small_kernel: ZPOST, ..., BOUND | the ops that precede the next sweep]
// and nothing is read back until a chunk of sweeps has been enqueued.  The update ORDER is the user's: the host
// passes one opcode per node.  Convergence is decided on the device and raises a stop word that every later
// kernel of the chunk checks on entry, so the posterior freezes at the iteration the reference stops at.
//
// The per-cluster work (two D x D factorisations per cluster and sweep) is one warp per cluster on the
// warp-cooperative SPD routines of warp_spd.cuh; cross-cluster sums are summed in a fixed order.
#include "warp_spd.cuh"
#include <vector>
#include <stdlib.h>

#define GV_THREADS 512
#define GV_MAXOPS 16

enum GmmVbField {
    G_PM0, G_PM1, G_GPM, G_PL0, G_PL1, G_GPL, G_PA, G_GPA, G_NG, G_LPREV,
    G_MPHI0, G_MPHI1, G_MU0, G_MCOV, G_MU1, G_MG,
    G_LPHI0, G_LPHI1, G_LU0, G_LU1, G_LG,
    G_APHI, G_AU, G_AG,
    G_ZG, G_ZH, G_ZLOGPI, G_ZT,
    G_STATS, G_XSTATS,
    G_COUNT
};
static const char *kGmmFieldNames[G_COUNT] = {
    "pm0", "pm1", "gpm", "pl0", "pl1", "gpl", "pa", "gpa", "ng", "lprev",
    "mu_phi0", "mu_phi1", "mu_u0", "mu_cov", "mu_u1", "mu_g",
    "lam_phi0", "lam_phi1", "lam_u0", "lam_u1", "lam_g",
    "al_phi", "al_u", "al_g",
    "z_g", "z_h", "z_logpi", "z_t",
    "stats", "xstats",
};

struct GmmVbOffsets { int64_t o[G_COUNT + 1]; };

static void gmm_vb_offsets(int D, int K, int64_t *o) {
    const int64_t KD = (int64_t)K * D, KDD = KD * D, NS = K + KD + KDD + 1;
    const int64_t size[G_COUNT] = {
        KD, KDD, K, KDD, K, K, K, 1, 1, 1,
        KD, KDD, KD, KDD, KDD, K,
        KDD, K, KDD, K, K,
        K, K, 1,
        K, KD, K, 1,
        NS, NS,
    };
    int64_t at = 0;
    for (int f = 0; f < G_COUNT; ++f) {
        o[f] = at;
        at += (size[f] + 1) & ~(int64_t)1;          // every field starts on a 16-byte boundary
    }
    o[G_COUNT] = at;
}

struct GmmVbArgs {
    int D, K;
    double *st;
    GmmVbOffsets off;
    int ops[GV_MAXOPS];
    int nops;
    double tol;
    double *Lhist;
    int cap;
    int *ctrl;              // [0] iterations finished, [1] stop, [2] error bits (1 not SPD, 2 domain)
};

// internal opcodes (the public ones are BPK_GMMOP_* in bpk.h; Z expands to ZPRE + sweep + ZPOST)
#define GOP_ZPRE 6
#define GOP_ZPOST 7

#define LOG2PI_GV 1.8378770664093453

// fixed-order sum of one value per warp (deterministic whatever the warp finishing order)
__device__ __forceinline__ void gv_cta_sums(double *red, int nq, const double *mine, int w, int lane, int nw, double *out) {
    __syncthreads();
    if (lane == 0)
        for (int q = 0; q < nq; ++q) red[w * 8 + q] = mine[q];
    __syncthreads();
    for (int q = 0; q < nq; ++q) {
        double s = 0.0;
        for (int i = 0; i < nw; ++i) s += red[i * 8 + q];
        out[q] = s;
    }
}

__global__ void __launch_bounds__(GV_THREADS, 1) gmm_vb_small_kernel(GmmVbArgs p) {
    extern __shared__ double sm[];
    __shared__ double red[(GV_THREADS / 32) * 8];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, nw = GV_THREADS >> 5;
    const int D = p.D, K = p.K, DD = D * D, ld = LD(D);
    const int64_t *o = p.off.o;
    double *st = p.st;
    int *ctrl = p.ctrl;
    volatile int *stop = ctrl + 1;
    // per warp: factor tile S, right-hand sides B, inverse C, two vectors
    double *S = sm + (size_t)w * (2 * (size_t)D * ld + (size_t)D * 32 + 2 * D);
    double *B = S + (size_t)D * ld;
    double *C = B + (size_t)D * 32;
    double *v0 = C + (size_t)D * ld, *v1 = v0 + D;
    const double *R = st + o[G_STATS], *S1 = R + K, *S2 = S1 + (size_t)K * D;

    for (int ip = 0; ip < p.nops; ++ip) {
        __syncthreads();
        if (*stop) break;
        const int op = p.ops[ip];
        if (op == BPK_GMMOP_MU) {
            // q(mu_k): phi = prior + [<Lam_k> s1_k, -1/2 R_k <Lam_k>]  (mixture.py:108-160, gaussian.py:341-375)
            for (int k = w; k < K; k += nw) {
                const double Rk = R[k];
                const double *Lk = st + o[G_LU0] + (size_t)k * DD;
                __syncwarp();
                for (int i = lane; i < D; i += 32) {
                    double s = st[o[G_PM0] + k * D + i];
                    for (int j = 0; j < D; ++j) s += Lk[i * D + j] * S1[k * D + j];
                    st[o[G_MPHI0] + k * D + i] = s;
                    v0[i] = s;
                }
                for (int e = lane; e < DD; e += 32) {
                    double p1 = st[o[G_PM1] + (size_t)k * DD + e] - 0.5 * Rk * Lk[e];
                    st[o[G_MPHI1] + (size_t)k * DD + e] = p1;
                    S[(e / D) * ld + (e % D)] = -2.0 * p1;
                }
                __syncwarp();
                int bad = warp_chol_upper(S, D, ld, lane);
                if (bad && lane == 0) atomicOr(&ctrl[2], BPK_FLAG_NOTSPD);
                const double ldt = warp_logdet(S, D, ld, lane);
                warp_inverse_from_factor(S, D, ld, B, lane, st + o[G_MCOV] + (size_t)k * DD, C);
                double dot = 0.0;
                for (int i = lane; i < D; i += 32) {
                    double s = 0.0;
                    for (int j = 0; j < D; ++j) s += C[i * ld + j] * v0[j];
                    st[o[G_MU0] + k * D + i] = s;
                    v1[i] = s;
                    dot += s * v0[i];
                }
                dot = warp_sum(dot);
                if (lane == 0) st[o[G_MG] + k] = -0.5 * dot + 0.5 * ldt;
                __syncwarp();
                for (int e = lane; e < DD; e += 32)
                    st[o[G_MU1] + (size_t)k * DD + e] = C[(e / D) * ld + (e % D)] + v1[e / D] * v1[e % D];
            }
        } else if (op == BPK_GMMOP_LAMBDA) {
            // q(Lam_k): phi = prior + [-1/2 (S2 - s1 mu^T - mu s1^T + R <mu mu^T>), R/2]  (gaussian.py:2496-2522)
            for (int k = w; k < K; k += nw) {
                const double Rk = R[k];
                const double *mu = st + o[G_MU0] + k * D, *mumu = st + o[G_MU1] + (size_t)k * DD;
                const double *s1 = S1 + k * D, *s2 = S2 + (size_t)k * DD;
                __syncwarp();
                for (int e = lane; e < DD; e += 32) {
                    const int i = e / D, j = e % D;
                    double tt = ((s2[e] - s1[i] * mu[j]) - mu[i] * s1[j]) + mumu[e] * Rk;
                    double p0 = st[o[G_PL0] + (size_t)k * DD + e] - 0.5 * tt;
                    st[o[G_LPHI0] + (size_t)k * DD + e] = p0;
                    S[i * ld + j] = -p0;
                }
                const double nu2 = st[o[G_PL1] + k] + 0.5 * Rk;
                __syncwarp();
                int bad = warp_chol_upper(S, D, ld, lane);
                if (bad && lane == 0) atomicOr(&ctrl[2], BPK_FLAG_NOTSPD);
                const double ldt = warp_logdet(S, D, ld, lane);
                warp_inverse_from_factor(S, D, ld, B, lane, nullptr, C);
                for (int e = lane; e < DD; e += 32)
                    st[o[G_LU0] + (size_t)k * DD + e] = nu2 * C[(e / D) * ld + (e % D)];
                if (lane == 0) {
                    st[o[G_LPHI1] + k] = nu2;
                    st[o[G_LU1] + k] = -ldt + bpk_mvdigamma(nu2, D);
                    st[o[G_LG] + k] = nu2 * ldt - bpk_mvlgamma(nu2, D);
                }
            }
        } else if (op == BPK_GMMOP_ALPHA) {
            // q(alpha): phi = a0 + R  (multinomial.py:83-90, dirichlet.py:120-170)
            if (w == 0) {
                double s = 0.0, lg = 0.0;
                int bad = 0;
                for (int k = lane; k < K; k += 32) {
                    double a = st[o[G_PA] + k] + R[k];
                    st[o[G_APHI] + k] = a;
                    if (!(a > 0.0)) bad = 1;
                    s += a;
                    lg += lgamma(a);
                }
                s = warp_sum(s);
                lg = warp_sum(lg);
                if (bad) atomicOr(&ctrl[2], BPK_FLAG_DOMAIN);
                const double ps = bpk_digamma(s);
                for (int k = lane; k < K; k += 32) st[o[G_AU] + k] = bpk_digamma(st[o[G_APHI] + k]) - ps;
                if (lane == 0) st[o[G_AG]] = lgamma(s) - lg;
            }
        } else if (op == GOP_ZPRE) {
            // parameters of the responsibilities kernel: g_k = -1/2 tr(<mu mu^T><Lam>) + 1/2 <log|Lam|>,
            // h_k = <Lam_k><mu_k>  (gaussian.py:377-394, :449-463), <log pi>  (categorical.py)
            for (int k = w; k < K; k += nw) {
                const double *Lk = st + o[G_LU0] + (size_t)k * DD, *mumu = st + o[G_MU1] + (size_t)k * DD;
                const double *mu = st + o[G_MU0] + k * D;
                double tr = 0.0;
                for (int e = lane; e < DD; e += 32) tr += Lk[e] * mumu[e];
                tr = warp_sum(tr);
                for (int i = lane; i < D; i += 32) {
                    double s = 0.0;
                    for (int j = 0; j < D; ++j) s += Lk[i * D + j] * mu[j];
                    st[o[G_ZH] + k * D + i] = s;
                }
                if (lane == 0) {
                    st[o[G_ZG] + k] = -0.5 * tr + 0.5 * st[o[G_LU1] + k];
                    st[o[G_ZLOGPI] + k] = st[o[G_AU] + k];
                }
            }
            const int64_t ns = o[G_XSTATS + 1] - o[G_XSTATS];
            for (int64_t e = t; e < ns; e += GV_THREADS) st[o[G_XSTATS] + e] = 0.0;
        } else if (op == GOP_ZPOST) {
            // the statistics of the sweep (summed over ranks) become current; T_q = sum_nk p_nk (message from Y)
            const int64_t ns = K + (int64_t)K * D + (int64_t)K * DD + 1;
            for (int64_t e = t; e < ns; e += GV_THREADS) st[o[G_STATS] + e] = st[o[G_XSTATS] + e];
            __syncthreads();
            double mine[1] = {0.0};
            for (int k = w; k < K; k += nw) {
                const double *Lk = st + o[G_LU0] + (size_t)k * DD;
                double a = 0.0;
                for (int e = lane; e < DD; e += 32) a += -0.5 * Lk[e] * S2[(size_t)k * DD + e];
                for (int i = lane; i < D; i += 32) a += st[o[G_ZH] + k * D + i] * S1[k * D + i];
                a = warp_sum(a);
                mine[0] += a + st[o[G_ZG] + k] * R[k];
            }
            double tot[1];
            gv_cta_sums(red, 1, mine, w, lane, nw, tot);
            if (t == 0) st[o[G_ZT]] = tot[0];
        } else if (op == BPK_GMMOP_BOUND) {
            // expfamily.py:400-480 for every node, from the plate-summed statistics
            double mine[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            for (int k = w; k < K; k += nw) {
                const double *Lk = st + o[G_LU0] + (size_t)k * DD, *mumu = st + o[G_MU1] + (size_t)k * DD;
                const double *mu = st + o[G_MU0] + k * D;
                double tr = 0.0, ls2 = 0.0, dmu1 = 0.0, dl0 = 0.0;
                for (int e = lane; e < DD; e += 32) {
                    tr += Lk[e] * mumu[e];
                    ls2 += Lk[e] * S2[(size_t)k * DD + e];
                    dmu1 += (st[o[G_PM1] + (size_t)k * DD + e] - st[o[G_MPHI1] + (size_t)k * DD + e]) * mumu[e];
                    dl0 += (st[o[G_PL0] + (size_t)k * DD + e] - st[o[G_LPHI0] + (size_t)k * DD + e]) * Lk[e];
                }
                double hs = 0.0, dmu0 = 0.0;
                for (int i = lane; i < D; i += 32) {
                    double h = 0.0;
                    for (int j = 0; j < D; ++j) h += Lk[i * D + j] * mu[j];
                    hs += h * S1[k * D + i];
                    dmu0 += (st[o[G_PM0] + k * D + i] - st[o[G_MPHI0] + k * D + i]) * mu[i];
                }
                tr = warp_sum(tr); ls2 = warp_sum(ls2); dmu1 = warp_sum(dmu1); dl0 = warp_sum(dl0);
                hs = warp_sum(hs); dmu0 = warp_sum(dmu0);
                const double lu1 = st[o[G_LU1] + k];
                const double gk = -0.5 * tr + 0.5 * lu1;
                mine[0] += gk * R[k] + hs - 0.5 * ls2;
                mine[1] += (st[o[G_AU] + k] - st[o[G_ZLOGPI] + k]) * R[k];
                mine[2] += (st[o[G_GPM] + k] - st[o[G_MG] + k]) + dmu0 + dmu1;
                mine[3] += (st[o[G_GPL] + k] - st[o[G_LG] + k]) + dl0 + (st[o[G_PL1] + k] - st[o[G_LPHI1] + k]) * lu1;
                mine[4] += (st[o[G_PA] + k] - st[o[G_APHI] + k]) * st[o[G_AU] + k];
            }
            double s[5];
            gv_cta_sums(red, 5, mine, w, lane, nw, s);
            if (t == 0) {
                const double lse = st[o[G_STATS] + K + (size_t)K * D + (size_t)K * DD];
                const double LY = s[0] - 0.5 * (double)D * LOG2PI_GV * st[o[G_NG]];
                const double LZ = (lse - st[o[G_ZT]]) + s[1];
                const double LM = s[2], LL = s[3];
                const double LA = (st[o[G_GPA]] - st[o[G_AG]]) + s[4];
                const double L = (((LY + LZ) + LM) + LL) + LA;
                int it = ctrl[0];
                if (it < p.cap) {
                    double *row = p.Lhist + (size_t)it * 6;
                    row[0] = LY; row[1] = LZ; row[2] = LM; row[3] = LL; row[4] = LA; row[5] = L;
                }
                const double L0 = st[o[G_LPREV]];
                st[o[G_LPREV]] = L;
                ctrl[0] = it + 1;
                // vmp.py:738-747 (tol < 0 or no previous bound: test disabled)
                if (p.tol >= 0.0 && L0 == L0) {
                    const double div = 0.5 * (fabs(L0) + fabs(L));
                    if ((L - L0) / div < p.tol) ctrl[1] = 1;
                }
                if (ctrl[2]) ctrl[1] = 1;
                __threadfence();
            }
        }
    }
}

extern "C" int bpk_gmm_vb_layout(int D, int K, int64_t *offsets, int *nfields) {
    if (D < 1 || K < 1) return bpk_set_error(BPK_EINVAL, "bpk_gmm_vb_layout: bad shape");
    if (offsets) gmm_vb_offsets(D, K, offsets);
    if (nfields) *nfields = G_COUNT;
    return BPK_OK;
}
extern "C" const char *bpk_gmm_vb_field_name(int i) { return (i >= 0 && i < G_COUNT) ? kGmmFieldNames[i] : nullptr; }

static int g_gv_timers[64];
static int g_gv_ntimers = 0, g_gv_timer_pos = 0;
extern "C" int bpk_gmm_vb_set_timers(const int *ids, int n) {
    if (n < 0 || n > 64) return bpk_set_error(BPK_EINVAL, "bpk_gmm_vb_set_timers: at most 64 timers");
    for (int i = 0; i < n; ++i) g_gv_timers[i] = ids[i];
    g_gv_ntimers = n;
    g_gv_timer_pos = 0;
    return BPK_OK;
}

static int gv_launch_small(GmmVbArgs &a, size_t smem) {
    if (a.nops == 0) return BPK_OK;
    BPK_LAUNCH(gmm_vb_small_kernel, 1, GV_THREADS, smem, a);
    a.nops = 0;
    return BPK_OK;
}

extern "C" int bpk_gmm_vb_run(const double *Y, int64_t N, int D, int K, double *P, double *gz, double *state,
                              const int *ops, int nops, int niter, double tol, double *Lhist, int cap, int *ctrl) {
    BPK_REQUIRE_INIT();
    if (D < 1 || D > BPK_MAXDIM || K < 1 || N < 0) return bpk_set_error(BPK_EINVAL, "bpk_gmm_vb_run: bad shape");
    if (nops < 1 || niter < 0) return bpk_set_error(BPK_EINVAL, "bpk_gmm_vb_run: empty program");
    if (!P || !state || !ctrl) return bpk_set_error(BPK_EINVAL, "bpk_gmm_vb_run: null buffer");
    for (int i = 0; i < nops; ++i)
        if (ops[i] < BPK_GMMOP_Z || ops[i] > BPK_GMMOP_BOUND)
            return bpk_set_error(BPK_EINVAL, "bpk_gmm_vb_run: unknown opcode %d at %d", ops[i], i);
    int nranks = 1, rank = 0;
    bpk_comm_size(&nranks, &rank);
    GmmVbArgs a;
    a.D = D; a.K = K; a.st = state; a.nops = 0; a.tol = tol; a.Lhist = Lhist; a.cap = cap; a.ctrl = ctrl;
    gmm_vb_offsets(D, K, a.off.o);
    const int64_t *o = a.off.o;
    const size_t smem = (size_t)(GV_THREADS / 32) * (2 * (size_t)D * LD(D) + (size_t)D * 32 + 2 * D) * sizeof(double);
    if (smem > (48u << 10))
        BPK_CUDA(cudaFuncSetAttribute(gmm_vb_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ns = K + (int64_t)K * D + (int64_t)K * D * D + 1;
    for (int it = 0; it < niter; ++it)
        for (int i = 0; i < nops; ++i) {
            if (ops[i] != BPK_GMMOP_Z) {
                a.ops[a.nops++] = ops[i];
                if (a.nops == GV_MAXOPS) { int rc = gv_launch_small(a, smem); if (rc) return rc; }
                continue;
            }
            a.ops[a.nops++] = GOP_ZPRE;
            int rc = gv_launch_small(a, smem);
            if (rc) return rc;
            int tid = -1;
            if (g_gv_timer_pos < g_gv_ntimers) tid = g_gv_timers[g_gv_timer_pos++];
            if (tid >= 0) bpk_timer_record(tid, 0);
            if (N > 0) {
                rc = bpk_gmm_sweep_resident(Y, N, D, K, state + o[G_ZG], state + o[G_ZH], state + o[G_LU0], state + o[G_ZLOGPI],
                                            P, gz, state + o[G_XSTATS], ctrl + 1);
                if (rc) return rc;
            }
            if (tid >= 0) bpk_timer_record(tid, 1);
            if (nranks > 1) {
                // harmless after a stop: the exchange buffer is only consumed by ZPOST, which then does not run
                rc = bpk_allreduce_sum_f64(state + o[G_XSTATS], ns);
                if (rc) return rc;
            }
            a.ops[a.nops++] = GOP_ZPOST;
        }
    return gv_launch_small(a, smem);
}
