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
// (Kept for D > 8, K > 64 and for bpk_gmm_stats; the tensor-pipe kernel v1 below serves the rest.)
#include "common.cuh"
#include <stdlib.h>

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
    const int *stop;    // resident loop: a raised word makes the launch a no-op (convergence was reached)
};

template <int PT>
__global__ void __launch_bounds__(GMM_ROWS, 1) gmm_sweep_kernel(GmmArgs a) {
    if (a.stop && *a.stop) return;
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

// =====================================================================================
// v1: both contractions on the fp64 tensor pipe (D <= 8, K <= 64).
// Per data row the monomial features z = [1, y_i (8), y_i y_j for i<=j (36), 0,0,0] (48, symmetric half
// only: 38 % fewer flops than the full y y^T) turn the E-step into the GEMM  L (rows x K) = Z Theta^T
// with Theta_k = [c_k + logpi_k, h_k, -1/2 Lam_ii, -(Lam_ij + Lam_ji)/2 ...] and the statistics into
// S (K x 48) = P^T Z.  CTA = 8 warps, tile = 128 rows:
//   E phase  warp w owns rows 16w..16w+15: builds its Z rows in shared memory, 192 DMMA.8x8x4 for L,
//            softmax in the accumulator layout (a row's 64 components live in 4 lanes: 2 shuffles),
//            writes P to HBM (16-byte stores, 64 B contiguous per lane group) and to the P tile;
//   M phase  warp w owns a 16 x 24 block of S and contracts it over all 128 rows of the tile (192 DMMA).
// 24 DMMA per row = 12.3 kflop; 576 B per row of HBM traffic: fp64-pipe bound.
// Pitches 52 / 68 doubles (= 4 mod 16) make every fragment load conflict-free.
// =====================================================================================
#define G1_ROWS 128
#define G1_WARPS 8
#define G1_KP 64
#define G1_DP 8
#define G1_FP 48
#define G1_LDZ 52
#define G1_LDP 68
#define G1_NF 45

__device__ __forceinline__ void g1_dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// feature f -> (i, j): f = 0 -> 1; 1..8 -> y_{f-1}; 9..44 -> y_i y_j (i <= j, row-major upper triangle)
__device__ __forceinline__ void g1_pair(int f, int &i, int &j) {
    int r = f - 9, ii = 0;
    while (r >= G1_DP - ii) { r -= G1_DP - ii; ++ii; }
    i = ii; j = ii + r;
}

__global__ void __launch_bounds__(G1_WARPS * 32, 1) gmm_sweep_dmma_kernel(GmmArgs a) {
    if (a.stop && *a.stop) return;
    extern __shared__ __align__(16) double sm[];
    double *sT = sm;                                   // [64][52]  Theta
    double *sZ = sT + G1_KP * G1_LDZ;                  // [128][52] features of the tile
    double *sP = sZ + G1_ROWS * G1_LDZ;                // [128][68] responsibilities of the tile
    __shared__ double red[G1_WARPS];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, gr = lane >> 2, tg = lane & 3;
    const int D = a.D, K = a.K;

    for (int e = t; e < G1_KP * G1_FP; e += blockDim.x) {
        const int k = e / G1_FP, f = e - k * G1_FP;
        double v = 0.0;
        if (k >= K) v = (f == 0) ? -1e300 : 0.0;        // padded components never win the softmax
        else if (f == 0) v = a.c[k] + a.logpi[k];
        else if (f <= G1_DP) v = (f - 1 < D) ? a.h[k * D + f - 1] : 0.0;
        else if (f < G1_NF) {
            int i, j;
            g1_pair(f, i, j);
            if (i < D && j < D)
                v = (i == j) ? -0.5 * a.Lam[(k * D + i) * D + i]
                             : -0.5 * (a.Lam[(k * D + i) * D + j] + a.Lam[(k * D + j) * D + i]);
        }
        sT[k * G1_LDZ + f] = v;
    }
    // statistics block of this warp: components 16*(w&3) .. +15, features 24*(w>>2) .. +23
    const int mb0 = 2 * (w & 3), fb0 = 3 * (w >> 2);
    double acc[2][3][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double lse_acc = 0.0;
    __syncthreads();

    const int64_t ntiles = (a.N + G1_ROWS - 1) / G1_ROWS;
    double ynext[G1_DP];
    {
        const int64_t n = (int64_t)blockIdx.x * G1_ROWS + w * 16 + (lane >> 1);
#pragma unroll
        for (int d = 0; d < G1_DP; ++d) ynext[d] = (n < a.N && d < D) ? a.Y[n * D + d] : 0.0;
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * G1_ROWS;
        // ---- E phase: this warp's 16 rows; two lanes per row build its 48 monomial features from registers ----
        const int r0 = w * 16;
        {
            const int r = r0 + (lane >> 1);
            const int64_t n = row0 + r;
            const bool live = n < a.N;
            double y[G1_DP];
#pragma unroll
            for (int d = 0; d < G1_DP; ++d) y[d] = (live && d < D) ? ynext[d] : 0.0;
            double *zr = sZ + r * G1_LDZ;
            if ((lane & 1) == 0) {
                zr[0] = live ? 1.0 : 0.0;
#pragma unroll
                for (int d = 0; d < G1_DP; ++d) zr[1 + d] = y[d];
            }
            {
                // the 36 products y_i y_j (i <= j) in feature order; features < 24 belong to the even lane
                int f = 9;
#pragma unroll
                for (int i = 0; i < G1_DP; ++i)
#pragma unroll
                    for (int j = i; j < G1_DP; ++j, ++f)
                        if ((f < 24) == ((lane & 1) == 0)) zr[f] = y[i] * y[j];
                if (lane & 1) zr[45] = zr[46] = zr[47] = 0.0;
            }
            // prefetch this lane's row of the next tile (hides the HBM latency behind both contraction phases)
            const int64_t nn = n + (int64_t)gridDim.x * G1_ROWS;
#pragma unroll
            for (int d = 0; d < G1_DP; ++d) ynext[d] = (nn < a.N && d < D) ? a.Y[nn * D + d] : 0.0;
        }
        __syncwarp();
        double L[2][8][2];
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) L[g][nb][0] = L[g][nb][1] = 0.0;
#pragma unroll 2
        for (int s = 0; s < G1_FP / 4; ++s) {
            const double z0 = sZ[(r0 + gr) * G1_LDZ + 4 * s + tg];
            const double z1 = sZ[(r0 + 8 + gr) * G1_LDZ + 4 * s + tg];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const double th = sT[(nb * 8 + gr) * G1_LDZ + 4 * s + tg];
                g1_dmma(L[0][nb][0], L[0][nb][1], z0, th);
                g1_dmma(L[1][nb][0], L[1][nb][1], z1, th);
            }
        }
        // softmax over the 64 components of row (g*8 + gr): 16 values here, the rest in lanes tg^1, tg^2
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int r = r0 + g * 8 + gr;
            const int64_t n = row0 + r;
            const bool live = n < a.N;
            double m = -INFINITY;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) m = fmax(m, fmax(L[g][nb][0], L[g][nb][1]));
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
            if (!isfinite(m)) m = 0.0;
            double ssum = 0.0;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                L[g][nb][0] = exp(L[g][nb][0] - m);
                L[g][nb][1] = exp(L[g][nb][1] - m);
                ssum += L[g][nb][0] + L[g][nb][1];
            }
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
            const double lse = log(ssum) + m;
            const double inv = 1.0 / ssum;
            double s2 = 0.0;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                L[g][nb][0] *= inv; L[g][nb][1] *= inv;
                s2 += L[g][nb][0] + L[g][nb][1];
            }
            s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
            const double inv2 = live ? 1.0 / s2 : 0.0;       // misc.py:1398-1401 second renormalisation
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const double p0 = L[g][nb][0] * inv2, p1 = L[g][nb][1] * inv2;
                const int k = nb * 8 + 2 * tg;
                sP[r * G1_LDP + k] = p0;
                sP[r * G1_LDP + k + 1] = p1;
                if (live && a.P) {
                    double *dst = a.P + n * K + k;
                    if (K == G1_KP) *reinterpret_cast<double2 *>(dst) = make_double2(p0, p1);
                    else { if (k < K) dst[0] = p0; if (k + 1 < K) dst[1] = p1; }
                }
            }
            if (live && tg == 0) {
                lse_acc += lse;
                if (a.g) a.g[n] = -lse;
            }
        }
        __syncthreads();
        // ---- M phase: S[16 x 24 block] += P^T Z over the 128 rows of the tile ----
#pragma unroll 4
        for (int q = 0; q < G1_ROWS / 4; ++q) {
            const int r = 4 * q + tg;
            const double pa0 = sP[r * G1_LDP + mb0 * 8 + gr], pa1 = sP[r * G1_LDP + (mb0 + 1) * 8 + gr];
#pragma unroll
            for (int fb = 0; fb < 3; ++fb) {
                const double zb = sZ[r * G1_LDZ + (fb0 + fb) * 8 + gr];
                g1_dmma(acc[0][fb][0], acc[0][fb][1], pa0, zb);
                g1_dmma(acc[1][fb][0], acc[1][fb][1], pa1, zb);
            }
        }
        __syncthreads();
    }
    // per-CTA partial in the v0 (k, f) layout with F = 1 + D + D*D, so that gmm_final_kernel serves both
    const int F = a.nfeat;
    double *pout = a.partial + (size_t)blockIdx.x * (K * F + 1);
    __syncthreads();
    double *sS = sZ;                                    // [64][48] staging of the CTA's S
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int fb = 0; fb < 3; ++fb)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                sS[((mb0 + i) * 8 + gr) * G1_FP + (fb0 + fb) * 8 + 2 * tg + j] = acc[i][fb][j];
    __syncthreads();
    for (int e = t; e < K * F; e += blockDim.x) {
        const int k = e / F, f = e - k * F;
        double v;
        if (f <= D) v = sS[k * G1_FP + f];              // 1, y_i  (f-1 < D <= 8: same slot)
        else {
            int i = (f - 1 - D) / D, j = (f - 1 - D) % D;
            if (i > j) { int tmp = i; i = j; j = tmp; }
            const int sf = 9 + i * G1_DP - (i * (i - 1)) / 2 + (j - i);
            v = sS[k * G1_FP + sf];
        }
        pout[e] = v;
    }
    lse_acc = warp_sum(lse_acc);
    if (lane == 0) red[w] = lse_acc;
    __syncthreads();
    if (t == 0) {
        double sum = 0.0;
        for (int ww = 0; ww < G1_WARPS; ++ww) sum += red[ww];
        pout[K * F] = sum;
    }
}

// =====================================================================================
// v2: the same two contractions, but the CTA-wide barriers between the E and M phases are gone.  The 8 warps
// form 4 PAIRS; a pair owns a tile of 32 rows: each warp runs the E phase of 16 of them (192 DMMA), the pair
// meets at its own named barrier (64 threads), then each warp contracts ALL 32 rows into its half of the
// statistics (all 64 components x 24 features: 48 accumulators, 192 DMMA).  The four pairs drift apart, so one
// pair's exp / softmax / stores overlap another pair's DMMA streams instead of everybody idling at a
// __syncthreads (v1: 56 % of the fp64 pipe).  255 registers per thread (8 warps per SM fit the register file).
// =====================================================================================
#define G2_PAIRS 4
#define G2_ROWS 32
__device__ __forceinline__ void g2_pair_sync(int p) {
    asm volatile("bar.sync %0, 64;\n" ::"r"(p + 1) : "memory");
}

__global__ void __launch_bounds__(G2_PAIRS * 64, 1) gmm_sweep_dmma2_kernel(GmmArgs a) {
    if (a.stop && *a.stop) return;
    extern __shared__ __align__(16) double sm[];
    double *sT = sm;                                                  // [64][52]  Theta
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, gr = lane >> 2, tg = lane & 3;
    const int pr = w >> 1, hf = w & 1;
    double *sZ = sT + G1_KP * G1_LDZ + (size_t)pr * (G2_ROWS * G1_LDZ + G2_ROWS * G1_LDP);   // [32][52] this pair's features
    double *sP = sZ + G2_ROWS * G1_LDZ;                                                       // [32][68] responsibilities
    __shared__ double red[2 * G2_PAIRS];
    const int D = a.D, K = a.K;

    for (int e = t; e < G1_KP * G1_FP; e += blockDim.x) {
        const int k = e / G1_FP, f = e - k * G1_FP;
        double v = 0.0;
        if (k >= K) v = (f == 0) ? -1e300 : 0.0;        // padded components never win the softmax
        else if (f == 0) v = a.c[k] + a.logpi[k];
        else if (f <= G1_DP) v = (f - 1 < D) ? a.h[k * D + f - 1] : 0.0;
        else if (f < G1_NF) {
            int i, j;
            g1_pair(f, i, j);
            if (i < D && j < D)
                v = (i == j) ? -0.5 * a.Lam[(k * D + i) * D + i]
                             : -0.5 * (a.Lam[(k * D + i) * D + j] + a.Lam[(k * D + j) * D + i]);
        }
        sT[k * G1_LDZ + f] = v;
    }
    // this warp's half of the statistics: all 8 component blocks x feature blocks 3*hf .. 3*hf+2
    const int fb0 = 3 * hf;
    double acc[8][3][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double lse_acc = 0.0;
    __syncthreads();

    const int64_t ntiles = (a.N + G2_ROWS - 1) / G2_ROWS;
    const int64_t tstride = (int64_t)gridDim.x * G2_PAIRS;
    const int r0 = hf * 16;                              // this warp's rows inside the pair tile
    double ynext[G1_DP];
    {
        const int64_t n = ((int64_t)blockIdx.x * G2_PAIRS + pr) * G2_ROWS + r0 + (lane >> 1);
#pragma unroll
        for (int d = 0; d < G1_DP; ++d) ynext[d] = (n < a.N && d < D) ? a.Y[n * D + d] : 0.0;
    }
    for (int64_t tile = (int64_t)blockIdx.x * G2_PAIRS + pr; tile < ntiles; tile += tstride) {
        const int64_t row0 = tile * G2_ROWS;
        // ---- E phase: this warp's 16 rows; two lanes per row build its 48 monomial features from registers ----
        {
            const int r = r0 + (lane >> 1);
            const int64_t n = row0 + r;
            const bool live = n < a.N;
            double y[G1_DP];
#pragma unroll
            for (int d = 0; d < G1_DP; ++d) y[d] = (live && d < D) ? ynext[d] : 0.0;
            double *zr = sZ + r * G1_LDZ;
            if ((lane & 1) == 0) {
                zr[0] = live ? 1.0 : 0.0;
#pragma unroll
                for (int d = 0; d < G1_DP; ++d) zr[1 + d] = y[d];
            }
            {
                int f = 9;
#pragma unroll
                for (int i = 0; i < G1_DP; ++i)
#pragma unroll
                    for (int j = i; j < G1_DP; ++j, ++f)
                        if ((f < 24) == ((lane & 1) == 0)) zr[f] = y[i] * y[j];
                if (lane & 1) zr[45] = zr[46] = zr[47] = 0.0;
            }
            const int64_t nn = n + tstride * G2_ROWS;       // this lane's row of the pair's next tile
#pragma unroll
            for (int d = 0; d < G1_DP; ++d) ynext[d] = (nn < a.N && d < D) ? a.Y[nn * D + d] : 0.0;
        }
        __syncwarp();
        double L[2][8][2];
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) L[g][nb][0] = L[g][nb][1] = 0.0;
#pragma unroll 2
        for (int s = 0; s < G1_FP / 4; ++s) {
            const double z0 = sZ[(r0 + gr) * G1_LDZ + 4 * s + tg];
            const double z1 = sZ[(r0 + 8 + gr) * G1_LDZ + 4 * s + tg];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const double th = sT[(nb * 8 + gr) * G1_LDZ + 4 * s + tg];
                g1_dmma(L[0][nb][0], L[0][nb][1], z0, th);
                g1_dmma(L[1][nb][0], L[1][nb][1], z1, th);
            }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int r = r0 + g * 8 + gr;
            const int64_t n = row0 + r;
            const bool live = n < a.N;
            double m = -INFINITY;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) m = fmax(m, fmax(L[g][nb][0], L[g][nb][1]));
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
            if (!isfinite(m)) m = 0.0;
            double ssum = 0.0;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                L[g][nb][0] = exp(L[g][nb][0] - m);
                L[g][nb][1] = exp(L[g][nb][1] - m);
                ssum += L[g][nb][0] + L[g][nb][1];
            }
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
            const double lse = log(ssum) + m;
            const double inv = 1.0 / ssum;
            double s2 = 0.0;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                L[g][nb][0] *= inv; L[g][nb][1] *= inv;
                s2 += L[g][nb][0] + L[g][nb][1];
            }
            s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
            const double inv2 = live ? 1.0 / s2 : 0.0;       // misc.py:1398-1401 second renormalisation
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const double p0 = L[g][nb][0] * inv2, p1 = L[g][nb][1] * inv2;
                const int k = nb * 8 + 2 * tg;
                sP[r * G1_LDP + k] = p0;
                sP[r * G1_LDP + k + 1] = p1;
                if (live && a.P) {
                    double *dst = a.P + n * K + k;
                    if (K == G1_KP) *reinterpret_cast<double2 *>(dst) = make_double2(p0, p1);
                    else { if (k < K) dst[0] = p0; if (k + 1 < K) dst[1] = p1; }
                }
            }
            if (live && tg == 0) {
                lse_acc += lse;
                if (a.g) a.g[n] = -lse;
            }
        }
        g2_pair_sync(pr);
        // ---- M phase: S[all 64 components x 24 features] += P^T Z over the pair's 32 rows ----
#pragma unroll 2
        for (int q = 0; q < G2_ROWS / 4; ++q) {
            const int r = 4 * q + tg;
            double zb[3];
#pragma unroll
            for (int fb = 0; fb < 3; ++fb) zb[fb] = sZ[r * G1_LDZ + (fb0 + fb) * 8 + gr];
#pragma unroll
            for (int mb = 0; mb < 8; ++mb) {
                const double pa = sP[r * G1_LDP + mb * 8 + gr];
#pragma unroll
                for (int fb = 0; fb < 3; ++fb) g1_dmma(acc[mb][fb][0], acc[mb][fb][1], pa, zb[fb]);
            }
        }
        g2_pair_sync(pr);          // the pair's buffers are free for its next tile
    }
    // per-CTA partial in the v0 (k, f) layout with F = 1 + D + D*D, so that gmm_final_kernel serves both
    const int F = a.nfeat;
    double *pout = a.partial + (size_t)blockIdx.x * (K * F + 1);
    __syncthreads();
    double *sS = sT + G1_KP * G1_LDZ;                   // [4 pairs][64][48] staging over the pair buffers (4 x 30.7 KB >= 96 KB)
#pragma unroll
    for (int mb = 0; mb < 8; ++mb)
#pragma unroll
        for (int fb = 0; fb < 3; ++fb)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                sS[(size_t)pr * G1_KP * G1_FP + (mb * 8 + gr) * G1_FP + (fb0 + fb) * 8 + 2 * tg + j] = acc[mb][fb][j];
    __syncthreads();
    for (int e = t; e < K * F; e += blockDim.x) {
        const int k = e / F, f = e - k * F;
        int sf;
        if (f <= D) sf = f;                               // 1, y_i  (f-1 < D <= 8: same slot)
        else {
            int i = (f - 1 - D) / D, j = (f - 1 - D) % D;
            if (i > j) { int tmp = i; i = j; j = tmp; }
            sf = 9 + i * G1_DP - (i * (i - 1)) / 2 + (j - i);
        }
        double v = 0.0;
#pragma unroll
        for (int pp = 0; pp < G2_PAIRS; ++pp) v += sS[(size_t)pp * G1_KP * G1_FP + k * G1_FP + sf];
        pout[e] = v;
    }
    lse_acc = warp_sum(lse_acc);
    if (lane == 0) red[w] = lse_acc;
    __syncthreads();
    if (t == 0) {
        double sum = 0.0;
        for (int ww = 0; ww < 2 * G2_PAIRS; ++ww) sum += red[ww];
        pout[K * F] = sum;
    }
}

// =====================================================================================
// v3: sixteen warps of 128 registers instead of eight of 255 (the v2 profile shows the fp64 tensor pipe 58 % busy with
// "wait" stalls on top: two warps per scheduler cannot cover each other's exp / softmax sections).  The warps form 4
// QUADS; a quad owns a tile of 32 rows: each warp runs the E phase of 8 of them (96 DMMA), the quad meets at its own
// named barrier (128 threads), then each warp contracts all 32 rows into a quarter of the statistics (32 components x
// 24 features: 24 accumulators, 96 DMMA).
// =====================================================================================
#define G3_QUADS 4
__device__ __forceinline__ void g3_quad_sync(int q) {
    asm volatile("bar.sync %0, 128;\n" ::"r"(q + 1) : "memory");
}

__global__ void __launch_bounds__(G3_QUADS * 128, 1) gmm_sweep_dmma3_kernel(GmmArgs a) {
    if (a.stop && *a.stop) return;
    extern __shared__ __align__(16) double sm[];
    double *sT = sm;                                                  // [64][52]  Theta
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, gr = lane >> 2, tg = lane & 3;
    const int qd = w >> 2, hf = w & 3;
    double *sZ = sT + G1_KP * G1_LDZ + (size_t)qd * (G2_ROWS * G1_LDZ + G2_ROWS * G1_LDP);   // [32][52] this quad's features
    double *sP = sZ + G2_ROWS * G1_LDZ;                                                       // [32][68] responsibilities
    __shared__ double red[4 * G3_QUADS];
    const int D = a.D, K = a.K;

    for (int e = t; e < G1_KP * G1_FP; e += blockDim.x) {
        const int k = e / G1_FP, f = e - k * G1_FP;
        double v = 0.0;
        if (k >= K) v = (f == 0) ? -1e300 : 0.0;        // padded components never win the softmax
        else if (f == 0) v = a.c[k] + a.logpi[k];
        else if (f <= G1_DP) v = (f - 1 < D) ? a.h[k * D + f - 1] : 0.0;
        else if (f < G1_NF) {
            int i, j;
            g1_pair(f, i, j);
            if (i < D && j < D)
                v = (i == j) ? -0.5 * a.Lam[(k * D + i) * D + i]
                             : -0.5 * (a.Lam[(k * D + i) * D + j] + a.Lam[(k * D + j) * D + i]);
        }
        sT[k * G1_LDZ + f] = v;
    }
    // this warp's quarter of the statistics: component blocks 4*(hf&1) .. +3, feature blocks 3*(hf>>1) .. +2
    const int mb0 = 4 * (hf & 1), fb0 = 3 * (hf >> 1);
    double acc[4][3][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double lse_acc = 0.0;
    __syncthreads();

    const int64_t ntiles = (a.N + G2_ROWS - 1) / G2_ROWS;
    const int64_t tstride = (int64_t)gridDim.x * G3_QUADS;
    const int r0 = hf * 8;                               // this warp's rows inside the quad tile
    const int part = lane & 3;                           // four lanes per row build 12 features each
    double ynext[G1_DP];
    {
        const int64_t n = ((int64_t)blockIdx.x * G3_QUADS + qd) * G2_ROWS + r0 + (lane >> 2);
#pragma unroll
        for (int d = 0; d < G1_DP; ++d) ynext[d] = (n < a.N && d < D) ? a.Y[n * D + d] : 0.0;
    }
    for (int64_t tile = (int64_t)blockIdx.x * G3_QUADS + qd; tile < ntiles; tile += tstride) {
        const int64_t row0 = tile * G2_ROWS;
        // ---- E phase: this warp's 8 rows; four lanes per row build its 48 monomial features from registers ----
        {
            const int r = r0 + (lane >> 2);
            const int64_t n = row0 + r;
            const bool live = n < a.N;
            double y[G1_DP];
#pragma unroll
            for (int d = 0; d < G1_DP; ++d) y[d] = (live && d < D) ? ynext[d] : 0.0;
            double *zr = sZ + r * G1_LDZ;
            if (part == 0) {
                zr[0] = live ? 1.0 : 0.0;
#pragma unroll
                for (int d = 0; d < G1_DP; ++d) zr[1 + d] = y[d];
            }
            {
                // the 36 products y_i y_j (i <= j) in feature order; lane `part` owns features 12*part .. 12*part+11
                int f = 9;
#pragma unroll
                for (int i = 0; i < G1_DP; ++i)
#pragma unroll
                    for (int j = i; j < G1_DP; ++j, ++f)
                        if (f / 12 == part) zr[f] = y[i] * y[j];
                if (part == 3) zr[45] = zr[46] = zr[47] = 0.0;
            }
            const int64_t nn = n + tstride * G2_ROWS;       // this lane's row of the quad's next tile
#pragma unroll
            for (int d = 0; d < G1_DP; ++d) ynext[d] = (nn < a.N && d < D) ? a.Y[nn * D + d] : 0.0;
        }
        __syncwarp();
        double L[8][2];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) L[nb][0] = L[nb][1] = 0.0;
#pragma unroll 3
        for (int s = 0; s < G1_FP / 4; ++s) {
            const double z0 = sZ[(r0 + gr) * G1_LDZ + 4 * s + tg];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const double th = sT[(nb * 8 + gr) * G1_LDZ + 4 * s + tg];
                g1_dmma(L[nb][0], L[nb][1], z0, th);
            }
        }
        {
            const int r = r0 + gr;
            const int64_t n = row0 + r;
            const bool live = n < a.N;
            double m = -INFINITY;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) m = fmax(m, fmax(L[nb][0], L[nb][1]));
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
            if (!isfinite(m)) m = 0.0;
            double ssum = 0.0;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                L[nb][0] = exp(L[nb][0] - m);
                L[nb][1] = exp(L[nb][1] - m);
                ssum += L[nb][0] + L[nb][1];
            }
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
            const double lse = log(ssum) + m;
            const double inv = 1.0 / ssum;
            double s2 = 0.0;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                L[nb][0] *= inv; L[nb][1] *= inv;
                s2 += L[nb][0] + L[nb][1];
            }
            s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
            const double inv2 = live ? 1.0 / s2 : 0.0;       // misc.py:1398-1401 second renormalisation
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const double p0 = L[nb][0] * inv2, p1 = L[nb][1] * inv2;
                const int k = nb * 8 + 2 * tg;
                sP[r * G1_LDP + k] = p0;
                sP[r * G1_LDP + k + 1] = p1;
                if (live && a.P) {
                    double *dst = a.P + n * K + k;
                    if (K == G1_KP) *reinterpret_cast<double2 *>(dst) = make_double2(p0, p1);
                    else { if (k < K) dst[0] = p0; if (k + 1 < K) dst[1] = p1; }
                }
            }
            if (live && tg == 0) {
                lse_acc += lse;
                if (a.g) a.g[n] = -lse;
            }
        }
        g3_quad_sync(qd);
        // ---- M phase: S[32 components x 24 features] += P^T Z over the quad's 32 rows ----
#pragma unroll 2
        for (int q = 0; q < G2_ROWS / 4; ++q) {
            const int r = 4 * q + tg;
            double zb[3];
#pragma unroll
            for (int fb = 0; fb < 3; ++fb) zb[fb] = sZ[r * G1_LDZ + (fb0 + fb) * 8 + gr];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                const double pa = sP[r * G1_LDP + (mb0 + mb) * 8 + gr];
#pragma unroll
                for (int fb = 0; fb < 3; ++fb) g1_dmma(acc[mb][fb][0], acc[mb][fb][1], pa, zb[fb]);
            }
        }
        g3_quad_sync(qd);          // the quad's buffers are free for its next tile
    }
    const int F = a.nfeat;
    double *pout = a.partial + (size_t)blockIdx.x * (K * F + 1);
    __syncthreads();
    double *sS = sT + G1_KP * G1_LDZ;                   // [4 quads][64][48] staging over the quad buffers
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int fb = 0; fb < 3; ++fb)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                sS[(size_t)qd * G1_KP * G1_FP + ((mb0 + mb) * 8 + gr) * G1_FP + (fb0 + fb) * 8 + 2 * tg + j] = acc[mb][fb][j];
    __syncthreads();
    for (int e = t; e < K * F; e += blockDim.x) {
        const int k = e / F, f = e - k * F;
        int sf;
        if (f <= D) sf = f;
        else {
            int i = (f - 1 - D) / D, j = (f - 1 - D) % D;
            if (i > j) { int tmp = i; i = j; j = tmp; }
            sf = 9 + i * G1_DP - (i * (i - 1)) / 2 + (j - i);
        }
        double v = 0.0;
#pragma unroll
        for (int pp = 0; pp < G3_QUADS; ++pp) v += sS[(size_t)pp * G1_KP * G1_FP + k * G1_FP + sf];
        pout[e] = v;
    }
    lse_acc = warp_sum(lse_acc);
    if (lane == 0) red[w] = lse_acc;
    __syncthreads();
    if (t == 0) {
        double sum = 0.0;
        for (int ww = 0; ww < 4 * G3_QUADS; ++ww) sum += red[ww];
        pout[K * F] = sum;
    }
}

// partial (k, f) layout -> caller's [sum p (K) | sum p y (K*D) | sum p yy^T (K*D*D) | lse]
__global__ void gmm_final_kernel(const double *__restrict__ partial, int nblocks, int K, int D,
                                 double *__restrict__ stats, const int *stop) {
    if (stop && *stop) return;
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
                   double *P, double *g, double *stats, int given_p, const int *stop);

extern "C" int bpk_gmm_sweep(const double *Y, int64_t N, int D, int K,
                             const double *c, const double *h, const double *Lam, const double *logpi,
                             double *P, double *g, double *stats) {
    BPK_REQUIRE_INIT();
    return gmm_run(Y, N, D, K, c, h, Lam, logpi, P, g, stats, 0, nullptr);
}

int bpk_gmm_sweep_resident(const double *Y, int64_t N, int D, int K, const double *c, const double *h, const double *Lam,
                           const double *logpi, double *P, double *g, double *stats, const int *stop) {
    return gmm_run(Y, N, D, K, c, h, Lam, logpi, P, g, stats, 0, stop);
}

extern "C" int bpk_gmm_stats(const double *Y, int64_t N, int D, int K, const double *P, double *stats) {
    BPK_REQUIRE_INIT();
    if (!P) return bpk_set_error(BPK_EINVAL, "bpk_gmm_stats: P is required");
    return gmm_run(Y, N, D, K, nullptr, nullptr, nullptr, nullptr, const_cast<double *>(P), nullptr, stats, 1, nullptr);
}

static int gmm_run(const double *Y, int64_t N, int D, int K,
                   const double *c, const double *h, const double *Lam, const double *logpi,
                   double *P, double *g, double *stats, int given_p, const int *stop) {
    if (D < 1 || D > GMM_MAXD) return bpk_set_error(BPK_EINVAL, "bpk_gmm_sweep: D=%d outside [1,%d]", D, GMM_MAXD);
    if (K < 1 || K > GMM_MAXK) return bpk_set_error(BPK_EINVAL, "bpk_gmm_sweep: K=%d outside [1,%d]", K, GMM_MAXK);
    if (N <= 0) return BPK_OK;
    GmmArgs a;
    a.Y = Y; a.N = N; a.D = D; a.K = K; a.c = c; a.h = h; a.Lam = Lam; a.logpi = logpi; a.P = P; a.g = g;
    a.given_p = given_p;
    a.stop = stop;
    a.nfeat = 1 + D + D * D;
    a.per_thread = (K * a.nfeat + GMM_ROWS - 1) / GMM_ROWS;
    size_t smem = ((size_t)K * (1 + D + D * D) + (size_t)GMM_ROWS * ((K | 1) + a.nfeat)) * sizeof(double);
    if (smem > (220u << 10)) return bpk_set_error(BPK_EINVAL, "bpk_gmm_sweep: K=%d, D=%d needs %zu B of shared memory", K, D, smem);
    int64_t ntiles = (N + GMM_ROWS - 1) / GMM_ROWS;
    int grid = g_bpk.sm_count;
    if (ntiles < grid) grid = (int)ntiles;
    a.partial = bpk_scratch((size_t)grid * (K * a.nfeat + 1) * sizeof(double));
    if (!a.partial) return bpk_set_error(BPK_ECUDA, "gmm: scratch allocation failed");
    if (!given_p && D <= G1_DP && K <= G1_KP && !getenv("BPK_GMM_V0")) {
        size_t smem1 = ((size_t)G1_KP * G1_LDZ + (size_t)G1_ROWS * G1_LDZ + (size_t)G1_ROWS * G1_LDP +
                        (size_t)G1_ROWS * G1_DP) * sizeof(double);
        int64_t nt1 = (N + G1_ROWS - 1) / G1_ROWS;
        int grid1 = g_bpk.sm_count;
        if (nt1 < grid1) grid1 = (int)nt1;
        a.partial = bpk_scratch((size_t)grid1 * (K * a.nfeat + 1) * sizeof(double));
        if (!a.partial) return bpk_set_error(BPK_ECUDA, "gmm: scratch allocation failed");
        if (getenv("BPK_GMM_V1")) {
            BPK_CUDA(cudaFuncSetAttribute(gmm_sweep_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            BPK_LAUNCH(gmm_sweep_dmma_kernel, grid1, G1_WARPS * 32, smem1, a);
        } else if (getenv("BPK_GMM_V3")) {
            // v3: warp quads, 16 warps of 128 registers
            size_t smem3 = ((size_t)G1_KP * G1_LDZ + (size_t)G3_QUADS * (G2_ROWS * G1_LDZ + G2_ROWS * G1_LDP)) * sizeof(double);
            BPK_CUDA(cudaFuncSetAttribute(gmm_sweep_dmma3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
            BPK_LAUNCH(gmm_sweep_dmma3_kernel, grid1, G3_QUADS * 128, smem3, a);
        } else {
            // v2: warp pairs on 32-row tiles (no CTA-wide barrier in the main loop)
            size_t smem2 = ((size_t)G1_KP * G1_LDZ + (size_t)G2_PAIRS * (G2_ROWS * G1_LDZ + G2_ROWS * G1_LDP)) * sizeof(double);
            BPK_CUDA(cudaFuncSetAttribute(gmm_sweep_dmma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            BPK_LAUNCH(gmm_sweep_dmma2_kernel, grid1, G2_PAIRS * 64, smem2, a);
        }
        int total1 = K * a.nfeat + 1;
        BPK_LAUNCH(gmm_final_kernel, (total1 + 127) / 128, 128, 0, a.partial, grid1, K, D, stats, stop);
        return BPK_OK;
    }
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
    BPK_LAUNCH(gmm_final_kernel, (total + 127) / 128, 128, 0, a.partial, grid, K, D, stats, stop);
    return BPK_OK;
}
