// warp_spd.cuh — warp-cooperative small SPD building blocks on a shared-memory tile
// (bayespy/utils/linalg.py:31-223: chol, chol_solve, chol_inv, chol_logdet), shared by
// linalg.cu's batched kernels and the resident mixture loop (gmm_vb.cu).
#pragma once
#include "common.cuh"

#define LD(D) ((D) | 1)

// ---- warp-level building blocks on a shared-memory tile ---------------------
// in-place upper Cholesky A = U^T U on S[D][ld] (upper triangle used/written).
__device__ __forceinline__ int warp_chol_upper(double *S, int D, int ld, int lane) {
    int bad = 0;
    for (int k = 0; k < D; ++k) {
        double akk = S[k * ld + k];
        if (!(akk > 0.0) || !isfinite(akk)) bad = 1;
        double d = sqrt(akk);
        double inv = 1.0 / d;
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

// Solve (U^T U) x = b for `ncol` columns held in B[D][32]; lane c owns column c.
__device__ __forceinline__ void warp_chol_solve_cols(const double *S, int D, int ld, double *B, int lane, int ncol) {
    if (lane < ncol) {
        for (int i = 0; i < D; ++i) {          // U^T y = b
            double s = B[i * 32 + lane];
            for (int j = 0; j < i; ++j) s -= S[j * ld + i] * B[j * 32 + lane];
            B[i * 32 + lane] = s / S[i * ld + i];
        }
        for (int i = D - 1; i >= 0; --i) {     // U x = y
            double s = B[i * 32 + lane];
            for (int j = i + 1; j < D; ++j) s -= S[i * ld + j] * B[j * 32 + lane];
            B[i * 32 + lane] = s / S[i * ld + i];
        }
    }
    __syncwarp();
}

__device__ __forceinline__ double warp_logdet(const double *S, int D, int ld, int lane) {
    double s = 0.0;
    for (int i = lane; i < D; i += 32) s += log(S[i * ld + i]);
    return 2.0 * warp_sum(s);
}

// Inverse of the SPD matrix whose factor is in S; result into global `out`
// (and optionally left in Sout[D][ld] for a fused consumer).
__device__ __forceinline__ void warp_inverse_from_factor(const double *S, int D, int ld, double *B, int lane,
                                                         double *out, double *Sout) {
    for (int c0 = 0; c0 < D; c0 += 32) {
        int nc = D - c0 < 32 ? D - c0 : 32;
        __syncwarp();
        for (int e = lane; e < D * nc; e += 32) {
            int i = e / nc, c = e % nc;
            B[i * 32 + c] = (i == c0 + c) ? 1.0 : 0.0;
        }
        __syncwarp();
        warp_chol_solve_cols(S, D, ld, B, lane, nc);
        for (int e = lane; e < D * nc; e += 32) {
            int i = e / nc, c = e % nc;
            double v = B[i * 32 + c];
            if (out) out[i * D + c0 + c] = v;
            if (Sout) Sout[i * ld + c0 + c] = v;
        }
    }
    __syncwarp();
}

