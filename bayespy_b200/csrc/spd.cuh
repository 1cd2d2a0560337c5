// spd.cuh — warp-cooperative dense SPD routines on a K x K tile held in shared memory.
// Used where ONE warp owns one small matrix (per-column precision of the masked factor
// model, the shared K x K posteriors of the resident PCA sweep).  Same arithmetic as
// linalg.py:31-223 (cho_factor(lower=False) -> U with A = U^T U; inverse by two
// triangular solves against the identity; log-det = 2 sum log diag U).
#pragma once
#include "common.cuh"

#define SPD_LD(D) ((D) | 1)      // odd pitch: conflict-free column walks

// In-place upper Cholesky of the (row-major, pitch ld) tile S; the strict lower triangle is
// left untouched.  Returns 1 in every lane if a pivot was not positive/finite.
__device__ __forceinline__ int spd_warp_chol_upper(double *S, int D, int ld, int lane) {
    int bad = 0;
    for (int k = 0; k < D; ++k) {
        double akk = S[k * ld + k];
        if (!(akk > 0.0) || !isfinite(akk)) bad = 1;
        double d = sqrt(akk), inv = 1.0 / d;
        __syncwarp();
        for (int j = k + lane; j < D; j += 32) S[k * ld + j] = (j == k) ? d : S[k * ld + j] * inv;
        __syncwarp();
        int m = D - k - 1;
        for (int idx = lane; idx < m * m; idx += 32) {
            int i = k + 1 + idx / m, j = k + 1 + idx % m;
            if (j >= i) S[i * ld + j] -= S[k * ld + i] * S[k * ld + j];
        }
        __syncwarp();
    }
    return bad;
}

// 2 sum_i log U_ii, same value in every lane
__device__ __forceinline__ double spd_warp_logdet(const double *S, int D, int ld, int lane) {
    double s = 0.0;
    for (int i = lane; i < D; i += 32) s += log(S[i * ld + i]);
    return 2.0 * warp_sum(s);
}

// C = (U^T U)^-1 for the factor U in S.  B is a K x 32 scratch tile (lane-per-column solves
// against the identity, 32 columns per pass).
__device__ __forceinline__ void spd_warp_inverse(const double *S, double *B, double *C, int K, int ld, int lane) {
    for (int c0 = 0; c0 < K; c0 += 32) {
        int nc = K - c0 < 32 ? K - c0 : 32;
        __syncwarp();
        for (int e = lane; e < K * nc; e += 32) {
            int i = e / nc, c = e % nc;
            B[i * 32 + c] = (i == c0 + c) ? 1.0 : 0.0;
        }
        __syncwarp();
        if (lane < nc) {
            for (int i = 0; i < K; ++i) {
                double s = B[i * 32 + lane];
                for (int j = 0; j < i; ++j) s -= S[j * ld + i] * B[j * 32 + lane];
                B[i * 32 + lane] = s / S[i * ld + i];
            }
            for (int i = K - 1; i >= 0; --i) {
                double s = B[i * 32 + lane];
                for (int j = i + 1; j < K; ++j) s -= S[i * ld + j] * B[j * 32 + lane];
                B[i * 32 + lane] = s / S[i * ld + i];
            }
        }
        __syncwarp();
        for (int e = lane; e < K * nc; e += 32) {
            int i = e / nc, c = e % nc;
            C[i * ld + c0 + c] = B[i * 32 + c];
        }
    }
    __syncwarp();
}

// In-place inverse of the SPD matrix held in the left half of the augmented tile G (K x 2K, pitch ldg,
// right half = identity) by Gauss-Jordan elimination without pivoting, ALL threads of the CTA working on
// every step (a warp-level Cholesky + triangular solves of the same tile took ~13 us, this takes ~3).
// On exit G[:, K:2K] = A^-1, scal[0] = log det A (sum of log pivots; linalg.py:209-223 gives the same
// number as 2 sum log diag U).  Non-positive pivot -> BPK_FLAG_NOTSPD.
// two = 1 (K even, >= 6K+1 threads): TWO pivots per step — the 2 x 2 pivot block of an SPD Schur complement is
// SPD, so it is inverted in closed form and rows k, k+1 are eliminated together: K/2 steps of two barriers
// instead of K (the steps are barrier-latency bound, not flop bound).  rowk: 2 x 2K, colk: 2 x K scratch.
// tt >= 0: the routine is run by a TEAM of tn threads (thread index tt within it) that meets at named barrier tbar
// instead of the whole CTA (pca_vb_ops runs two small ops side by side on the two halves of its CTA).
#define SPD_GJ_SYNC()                                                                        \
    do {                                                                                     \
        if (tt < 0) __syncthreads();                                                         \
        else asm volatile("bar.sync %0, %1;\n" ::"r"(tbar), "r"(nt) : "memory");             \
    } while (0)
template <int KC>
__device__ __forceinline__ void spd_cta_inverse_gj(double *G, double *rowk, double *colk, double *piv, int Krt, double *scal,
                                                   int *flagword, int two = 0, int tt = -1, int tn = 0, int tbar = 0) {
    const int K = KC ? KC : Krt, K2 = 2 * K, ldg = K2 + 1;
    const int t = tt < 0 ? (int)threadIdx.x : tt, nt = tt < 0 ? (int)blockDim.x : tn;
    if (two && (K & 1) == 0 && nt >= 2 * K2 + 2 * K + 1) {
        for (int k = 0; k < K; k += 2) {
            SPD_GJ_SYNC();
            const double p00 = G[k * ldg + k], p01 = G[k * ldg + k + 1];
            const double p10 = G[(k + 1) * ldg + k], p11 = G[(k + 1) * ldg + k + 1];
            const double det = p00 * p11 - p01 * p10;
            const double rd = 1.0 / det;
            if (t < 2 * K2) {
                const int r = t >= K2, j = t - r * K2;
                const double g0 = G[k * ldg + j], g1 = G[(k + 1) * ldg + j];
                rowk[t] = r ? (p00 * g1 - p10 * g0) * rd : (p11 * g0 - p01 * g1) * rd;     // P^-1 [row k; row k+1]
            } else if (t < 2 * K2 + 2 * K) {
                const int c = (t - 2 * K2) >= K, i = t - 2 * K2 - c * K;
                colk[c * K + i] = G[i * ldg + k + c];
            } else if (t == 2 * K2 + 2 * K) {
                piv[k] = p00;                 // both must be positive for an SPD matrix; their product is the block's det
                piv[k + 1] = det / p00;
            }
            SPD_GJ_SYNC();
            for (int e = t; e < K * K2; e += nt) {
                const int i = e / K2, j = e - i * K2;
                const double r0 = rowk[j], r1 = rowk[K2 + j];
                double v;
                if (i == k) v = r0;
                else if (i == k + 1) v = r1;
                else v = G[i * ldg + j] - (colk[i] * r0 + colk[K + i] * r1);
                G[i * ldg + j] = v;
            }
        }
    } else {
    for (int k = 0; k < K; ++k) {
        SPD_GJ_SYNC();
        const double p = G[k * ldg + k];
        const double r = 1.0 / p;
        if (t < K2) rowk[t] = G[k * ldg + t] * r;
        else if (t < K2 + K) colk[t - K2] = G[(t - K2) * ldg + k];
        if (t == K2 + K) piv[k] = p;
        if (nt < K2 + K + 1) {          // narrow blocks: let the first threads take the leftovers
            for (int e = t + nt; e < K2 + K + 1; e += nt) {
                if (e < K2) rowk[e] = G[k * ldg + e] * r;
                else if (e < K2 + K) colk[e - K2] = G[(e - K2) * ldg + k];
                else piv[k] = p;
            }
        }
        SPD_GJ_SYNC();
        for (int e = t; e < K * K2; e += nt) {
            const int i = e / K2, j = e - i * K2;
            const double rj = rowk[j];
            G[i * ldg + j] = (i == k) ? rj : G[i * ldg + j] - colk[i] * rj;
        }
    }
    }
    SPD_GJ_SYNC();
    if (t < 32) {
        double s = 0.0;
        int bad = 0;
        for (int k = t; k < K; k += 32) {
            const double p = piv[k];
            if (!(p > 0.0) || !isfinite(p)) bad = 1;
            s += log(p);
        }
        s = warp_sum(s);
        bad = __any_sync(0xffffffffu, bad);
        if (t == 0) {
            scal[0] = s;
            if (bad) atomicOr(flagword, BPK_FLAG_NOTSPD);
        }
    }
    SPD_GJ_SYNC();
}


#undef SPD_GJ_SYNC

// (A single-warp, register-resident variant of this elimination for K = 16 — lane = column, shuffles instead of
// barriers — was measured SLOWER inside the fused sweep's tail: 20.6 us vs 10.4 us per ROW op; it is not kept.)
