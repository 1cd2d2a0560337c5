// linalg.cu — batched small SPD linear algebra (bayespy/utils/linalg.py:31-223)
// and the fused Gaussian / Wishart moment kernels built on it.
//
// Mapping: one warp owns one D x D matrix staged in shared memory (leading
// dimension odd -> conflict-free fp64 columns).  Factorisation is
// warp-cooperative; solves/inverses are lane-per-right-hand-side so that every
// lane reads the same U element (smem broadcast) and its own rhs column.
// The reference does this with a Python loop calling LAPACK per matrix.
#include "warp_spd.cuh"

static inline size_t tile_doubles(int D) { return (size_t)D * LD(D) + (size_t)D * 32; }
static inline int warps_per_block(int D) {
    size_t per = tile_doubles(D) * sizeof(double);
    int w = (int)((160u << 10) / per);
    if (w > 8) w = 8;
    if (w < 1) w = 1;
    return w;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > (48u << 10))
        BPK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return BPK_OK;
}

// ---- chol ---------------------------------------------------------------------
__global__ void chol_kernel(const double *__restrict__ A, double *__restrict__ U, int64_t batch, int D, int *flag) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ld = LD(D);
    double *S = smem + (size_t)w * ((size_t)D * ld + (size_t)D * 32);
    for (int64_t b = (int64_t)blockIdx.x * nw + w; b < batch; b += (int64_t)gridDim.x * nw) {
        const double *a = A + b * D * D;
        for (int e = lane; e < D * D; e += 32) S[(e / D) * ld + (e % D)] = a[e];
        __syncwarp();
        int bad = warp_chol_upper(S, D, ld, lane);
        if (bad && lane == 0) atomicOr(flag, BPK_FLAG_NOTSPD);
        double *u = U + b * D * D;
        for (int e = lane; e < D * D; e += 32) {
            int i = e / D, j = e % D;
            u[e] = (j >= i) ? S[i * ld + j] : 0.0;
        }
        __syncwarp();
    }
}

extern "C" int bpk_chol(const double *A, double *U, int64_t batch, int D, int check) {
    BPK_REQUIRE_INIT();
    if (D < 1 || D > BPK_MAXDIM) return bpk_set_error(BPK_EINVAL, "bpk_chol: D=%d outside [1,%d]", D, BPK_MAXDIM);
    if (batch <= 0) return BPK_OK;
    int nw = warps_per_block(D);
    size_t smem = nw * tile_doubles(D) * sizeof(double);
    int rc = set_smem(chol_kernel, smem);
    if (rc) return rc;
    int64_t blocks = (batch + nw - 1) / nw;
    int64_t cap = (int64_t)g_bpk.sm_count * 8;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(chol_kernel, (unsigned)blocks, nw * 32, smem, A, U, batch, D, g_bpk.d_flag);
    if (check) return bpk_check_flag(BPK_ENOTSPD);
    return BPK_OK;
}

// ---- chol_solve ---------------------------------------------------------------
__global__ void chol_solve_kernel(const double *__restrict__ U, int64_t batchU,
                                  const double *__restrict__ Bm, int64_t batchB,
                                  double *__restrict__ X, int64_t batch, int D, int nrhs) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ld = LD(D);
    double *S = smem + (size_t)w * ((size_t)D * ld + (size_t)D * 32);
    double *B = S + (size_t)D * ld;
    for (int64_t b = (int64_t)blockIdx.x * nw + w; b < batch; b += (int64_t)gridDim.x * nw) {
        const double *u = U + (batchU == 1 ? 0 : b) * D * D;
        for (int e = lane; e < D * D; e += 32) S[(e / D) * ld + (e % D)] = u[e];
        const double *bm = Bm + (batchB == 1 ? 0 : b) * (int64_t)D * nrhs;
        double *x = X + b * (int64_t)D * nrhs;
        for (int c0 = 0; c0 < nrhs; c0 += 32) {
            int nc = nrhs - c0 < 32 ? nrhs - c0 : 32;
            __syncwarp();
            for (int e = lane; e < D * nc; e += 32) {
                int i = e / nc, c = e % nc;
                B[i * 32 + c] = bm[(int64_t)i * nrhs + c0 + c];
            }
            __syncwarp();
            warp_chol_solve_cols(S, D, ld, B, lane, nc);
            for (int e = lane; e < D * nc; e += 32) {
                int i = e / nc, c = e % nc;
                x[(int64_t)i * nrhs + c0 + c] = B[i * 32 + c];
            }
        }
        __syncwarp();
    }
}

extern "C" int bpk_chol_solve(const double *U, int64_t batchU, const double *B, int64_t batchB,
                              double *X, int64_t batch, int D, int nrhs) {
    BPK_REQUIRE_INIT();
    if (D < 1 || D > BPK_MAXDIM) return bpk_set_error(BPK_EINVAL, "bpk_chol_solve: D=%d outside [1,%d]", D, BPK_MAXDIM);
    if (nrhs < 1) return bpk_set_error(BPK_EINVAL, "bpk_chol_solve: nrhs=%d", nrhs);
    if ((batchU != 1 && batchU != batch) || (batchB != 1 && batchB != batch))
        return bpk_set_error(BPK_EINVAL, "bpk_chol_solve: operand batch must be 1 or %lld", (long long)batch);
    if (batch <= 0) return BPK_OK;
    int nw = warps_per_block(D);
    size_t smem = nw * tile_doubles(D) * sizeof(double);
    int rc = set_smem(chol_solve_kernel, smem);
    if (rc) return rc;
    int64_t blocks = (batch + nw - 1) / nw;
    int64_t cap = (int64_t)g_bpk.sm_count * 8;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(chol_solve_kernel, (unsigned)blocks, nw * 32, smem, U, batchU, B, batchB, X, batch, D, nrhs);
    return BPK_OK;
}

// ---- chol_inv -------------------------------------------------------------------
__global__ void chol_inv_kernel(const double *__restrict__ U, double *__restrict__ Ainv, int64_t batch, int D) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ld = LD(D);
    double *S = smem + (size_t)w * ((size_t)D * ld + (size_t)D * 32);
    double *B = S + (size_t)D * ld;
    for (int64_t b = (int64_t)blockIdx.x * nw + w; b < batch; b += (int64_t)gridDim.x * nw) {
        const double *u = U + b * D * D;
        __syncwarp();
        for (int e = lane; e < D * D; e += 32) S[(e / D) * ld + (e % D)] = u[e];
        __syncwarp();
        warp_inverse_from_factor(S, D, ld, B, lane, Ainv + b * D * D, nullptr);
    }
}

extern "C" int bpk_chol_inv(const double *U, double *Ainv, int64_t batch, int D) {
    BPK_REQUIRE_INIT();
    if (D < 1 || D > BPK_MAXDIM) return bpk_set_error(BPK_EINVAL, "bpk_chol_inv: D=%d outside [1,%d]", D, BPK_MAXDIM);
    if (batch <= 0) return BPK_OK;
    int nw = warps_per_block(D);
    size_t smem = nw * tile_doubles(D) * sizeof(double);
    int rc = set_smem(chol_inv_kernel, smem);
    if (rc) return rc;
    int64_t blocks = (batch + nw - 1) / nw;
    int64_t cap = (int64_t)g_bpk.sm_count * 8;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(chol_inv_kernel, (unsigned)blocks, nw * 32, smem, U, Ainv, batch, D);
    return BPK_OK;
}

// ---- chol_logdet ------------------------------------------------------------------
__global__ void chol_logdet_kernel(const double *__restrict__ U, double *__restrict__ out, int64_t batch, int D) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; b < batch; b += (int64_t)gridDim.x * blockDim.x) {
        const double *u = U + b * D * D;
        double s = 0.0;
        for (int i = 0; i < D; ++i) s += log(u[i * D + i]);
        out[b] = 2.0 * s;
    }
}

extern "C" int bpk_chol_logdet(const double *U, double *out, int64_t batch, int D) {
    BPK_REQUIRE_INIT();
    if (D < 1) return bpk_set_error(BPK_EINVAL, "bpk_chol_logdet: D=%d", D);
    if (batch <= 0) return BPK_OK;
    int64_t blocks = (batch + 127) / 128;
    int64_t cap = (int64_t)g_bpk.sm_count * 16;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(chol_logdet_kernel, (unsigned)blocks, 128, 0, U, out, batch, D);
    return BPK_OK;
}

// ---- fused Gaussian moments (gaussian.py:397-446, :672-706) -----------------------
// Stage 1: per precision matrix: Cov = (-2 phi1)^-1, logdet; and when the mean
// natural parameter has the same batch, u0 and g too.
__global__ void gauss_factor_kernel(const double *__restrict__ phi0, int64_t n0,
                                    const double *__restrict__ phi1, int64_t n1, int K,
                                    double *__restrict__ u0, double *__restrict__ cov,
                                    double *__restrict__ g, double *__restrict__ logdet,
                                    int fuse_mean, int *flag) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ld = LD(K);
    // per warp: factor tile S, rhs tile B, inverse tile C
    double *S = smem + (size_t)w * (2 * (size_t)K * ld + (size_t)K * 32);
    double *B = S + (size_t)K * ld;
    double *C = B + (size_t)K * 32;
    for (int64_t b = (int64_t)blockIdx.x * nw + w; b < n1; b += (int64_t)gridDim.x * nw) {
        const double *p1 = phi1 + b * K * K;
        __syncwarp();
        for (int e = lane; e < K * K; e += 32) S[(e / K) * ld + (e % K)] = -2.0 * p1[e];
        __syncwarp();
        int bad = warp_chol_upper(S, K, ld, lane);
        if (bad && lane == 0) atomicOr(flag, BPK_FLAG_NOTSPD);
        double ldt = warp_logdet(S, K, ld, lane);
        if (logdet && lane == 0) logdet[b] = ldt;
        warp_inverse_from_factor(S, K, ld, B, lane, cov ? cov + b * K * K : nullptr, C);
        if (fuse_mean) {
            const double *p0 = phi0 + (n0 == 1 ? 0 : b) * K;
            double dot = 0.0;
            for (int i = lane; i < K; i += 32) {
                double s = 0.0;
                for (int j = 0; j < K; ++j) s += C[i * ld + j] * p0[j];
                if (u0) u0[b * K + i] = s;
                dot += s * p0[i];
            }
            dot = warp_sum(dot);
            if (g && lane == 0) g[b] = -0.5 * dot + 0.5 * ldt;
        }
    }
}

// Stage 2: shared covariance applied to N mean parameters:
// u0[n] = Cov phi0[n],  g[n] = -1/2 u0[n].phi0[n] + 1/2 logdet.
// One thread per (n); Cov staged in smem (broadcast reads), rows staged in
// smem with coalesced loads.
__global__ void __launch_bounds__(128) gauss_apply_kernel(const double *__restrict__ phi0, const double *__restrict__ cov,
                                   const double *__restrict__ logdet, int64_t N, int K,
                                   double *__restrict__ u0, double *__restrict__ g) {
    extern __shared__ double smem[];
    const int ldr = K | 1;
    double *C = smem;                       // [K][K]
    double *R = smem + K * K;               // [128][ldr] input rows
    for (int e = threadIdx.x; e < K * K; e += blockDim.x) C[e] = cov[e];
    const double hl = 0.5 * logdet[0];
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < N; base += (int64_t)gridDim.x * blockDim.x) {
        int64_t rows = N - base < blockDim.x ? N - base : blockDim.x;
        __syncthreads();
        for (int64_t e = threadIdx.x; e < rows * K; e += blockDim.x)
            R[(e / K) * ldr + (e % K)] = phi0[base * K + e];
        __syncthreads();
        int r = threadIdx.x;
        double dot = 0.0;
        if (r < rows) {
            for (int i = 0; i < K; ++i) {
                double s = 0.0;
                for (int j = 0; j < K; ++j) s += C[i * K + j] * R[r * ldr + j];
                dot += s * R[r * ldr + i];
                // stage the result in place of nothing: write straight out (strided, L2-merged)
                if (u0) u0[(base + r) * K + i] = s;
            }
            if (g) g[base + r] = -0.5 * dot + hl;
        }
    }
}

extern "C" int bpk_gaussian_moments(const double *phi0, int64_t n0, const double *phi1, int64_t n1,
                                    int64_t N, int K, double *u0, double *cov, double *g,
                                    double *logdet, int check) {
    BPK_REQUIRE_INIT();
    if (K < 1 || K > BPK_MAXDIM) return bpk_set_error(BPK_EINVAL, "bpk_gaussian_moments: K=%d outside [1,%d]", K, BPK_MAXDIM);
    if ((n0 != 1 && n0 != N) || (n1 != 1 && n1 != N))
        return bpk_set_error(BPK_EINVAL, "bpk_gaussian_moments: n0,n1 must be 1 or N");
    if (N <= 0) return BPK_OK;
    const int ld = LD(K);
    size_t per = (2 * (size_t)K * ld + (size_t)K * 32) * sizeof(double);
    int nw = (int)((160u << 10) / per);
    if (nw > 8) nw = 8;
    if (nw < 1) nw = 1;
    size_t smem = nw * per;
    int rc = set_smem(gauss_factor_kernel, smem);
    if (rc) return rc;
    const bool shared_cov = (n1 == 1 && N > 1);
    double *ld_dev = logdet;
    if (shared_cov && !ld_dev) {
        ld_dev = bpk_scratch(sizeof(double));
        if (!ld_dev) return bpk_set_error(BPK_ECUDA, "scratch allocation failed");
    }
    // shared covariance needs a dense copy for stage 2
    double *cov_dev = cov;
    if (shared_cov && !cov_dev) {
        double *s = bpk_scratch((1 + (size_t)K * K) * sizeof(double));
        if (!s) return bpk_set_error(BPK_ECUDA, "scratch allocation failed");
        if (!logdet) ld_dev = s;
        cov_dev = s + 1;
    }
    int64_t blocks = (n1 + nw - 1) / nw;
    int64_t cap = (int64_t)g_bpk.sm_count * 8;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(gauss_factor_kernel, (unsigned)blocks, nw * 32, smem, phi0, n0, phi1, n1, K,
               u0, cov_dev, g, ld_dev, shared_cov ? 0 : 1, g_bpk.d_flag);
    if (shared_cov && (u0 || g)) {
        size_t sm2 = ((size_t)K * K + 128 * (size_t)(K | 1)) * sizeof(double);
        rc = set_smem(gauss_apply_kernel, sm2);
        if (rc) return rc;
        int64_t b2 = (N + 127) / 128;
        int64_t cap2 = (int64_t)g_bpk.sm_count * 8;
        if (b2 > cap2) b2 = cap2;
        BPK_LAUNCH(gauss_apply_kernel, (unsigned)b2, 128, sm2, phi0, cov_dev, ld_dev, N, K, u0, g);
    }
    if (check) return bpk_check_flag(BPK_ENOTSPD);
    return BPK_OK;
}

__global__ void outer_add_kernel(const double *__restrict__ u0, const double *__restrict__ cov, int64_t ncov,
                                 int64_t N, int K, double *__restrict__ u1) {
    const int64_t KK = (int64_t)K * K;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; e < N * KK; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = e / KK;
        int r = (int)(e - n * KK);
        int i = r / K, j = r - i * K;
        double c = cov ? cov[(ncov == 1 ? 0 : n) * KK + r] : 0.0;
        u1[e] = c + u0[n * K + i] * u0[n * K + j];
    }
}

extern "C" int bpk_outer_add(const double *u0, const double *cov, int64_t ncov,
                             int64_t N, int K, double *u1) {
    BPK_REQUIRE_INIT();
    if (N <= 0 || K <= 0) return BPK_OK;
    int64_t total = N * (int64_t)K * K;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)g_bpk.sm_count * 32;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(outer_add_kernel, (unsigned)blocks, 256, 0, u0, cov, ncov, N, K, u1);
    return BPK_OK;
}

// ---- Wishart moments (wishart.py:165-188) -----------------------------------------
__global__ void wishart_kernel(const double *__restrict__ phi0, const double *__restrict__ phi1, int64_t n1,
                               int64_t n, int D, double *__restrict__ u0, double *__restrict__ u1,
                               double *__restrict__ g, int *flag) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ld = LD(D);
    double *S = smem + (size_t)w * (2 * (size_t)D * ld + (size_t)D * 32);
    double *B = S + (size_t)D * ld;
    double *C = B + (size_t)D * 32;
    for (int64_t b = (int64_t)blockIdx.x * nw + w; b < n; b += (int64_t)gridDim.x * nw) {
        const double *p0 = phi0 + b * D * D;
        const double nu2 = phi1[n1 == 1 ? 0 : b];
        __syncwarp();
        for (int e = lane; e < D * D; e += 32) S[(e / D) * ld + (e % D)] = -p0[e];
        __syncwarp();
        int bad = warp_chol_upper(S, D, ld, lane);
        if (bad && lane == 0) atomicOr(flag, BPK_FLAG_NOTSPD);
        double ldt = warp_logdet(S, D, ld, lane);
        warp_inverse_from_factor(S, D, ld, B, lane, nullptr, C);
        if (u0)
            for (int e = lane; e < D * D; e += 32) u0[b * D * D + e] = nu2 * C[(e / D) * ld + (e % D)];
        if (lane == 0) {
            if (u1) u1[b] = -ldt + bpk_mvdigamma(nu2, D);
            if (g) g[b] = nu2 * ldt - bpk_mvlgamma(nu2, D);
        }
    }
}

extern "C" int bpk_wishart_moments(const double *phi0, const double *phi1, int64_t n1,
                                   int64_t n, int D, double *u0, double *u1, double *g, int check) {
    BPK_REQUIRE_INIT();
    if (D < 1 || D > BPK_MAXDIM) return bpk_set_error(BPK_EINVAL, "bpk_wishart_moments: D=%d outside [1,%d]", D, BPK_MAXDIM);
    if (n1 != 1 && n1 != n) return bpk_set_error(BPK_EINVAL, "bpk_wishart_moments: n1 must be 1 or n");
    if (n <= 0) return BPK_OK;
    const int ld = LD(D);
    size_t per = (2 * (size_t)D * ld + (size_t)D * 32) * sizeof(double);
    int nw = (int)((160u << 10) / per);
    if (nw > 8) nw = 8;
    if (nw < 1) nw = 1;
    size_t smem = nw * per;
    int rc = set_smem(wishart_kernel, smem);
    if (rc) return rc;
    int64_t blocks = (n + nw - 1) / nw;
    int64_t cap = (int64_t)g_bpk.sm_count * 8;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(wishart_kernel, (unsigned)blocks, nw * 32, smem, phi0, phi1, n1, n, D, u0, u1, g, g_bpk.d_flag);
    if (check) return bpk_check_flag(BPK_ENOTSPD);
    return BPK_OK;
}
