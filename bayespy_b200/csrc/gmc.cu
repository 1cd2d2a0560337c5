// gmc.cu — block-tridiagonal SPD solve of a Gaussian Markov chain (information-form RTS smoother)
//
// Reference: utils/linalg.py:468-575 block_banded_solve, the whole of
// TemplateGaussianMarkovChainDistribution.compute_moments_and_cgf (gaussian_markov_chain.py:89-123): a
// Python loop over the T time steps with ~8 SciPy calls per step (340 us/step measured).
// The posterior precision of the chain is block tridiagonal,
//      [ A_0  B_0                 ]
//      [ B_0' A_1  B_1            ]      A_n: D x D diagonal blocks,  B_n: super-diagonal blocks
//      [      B_1' A_2  ...       ]
// and VMP needs from its inverse only the diagonal blocks V_n = Cov(x_n), the super-diagonal blocks
// C_n = Cov(x_n, x_{n+1}), the solution x = P^-1 y and log det P.
//
// v1 mapping: one CTA per chain (the plate axis of the node = grid dimension), the recursion over n runs
// inside the kernel, and every D x D block operation of a step uses all threads of the CTA:
//   forward   S_n = A_n - B_{n-1}' S_{n-1}^-1 B_{n-1}  (symmetrised, as the reference does),
//             S_n^-1 by all-thread Gauss-Jordan (spd.cuh; log det = sum of log pivots),
//             C_n = S_n^-1 B_n,  xt_{n+1} = y_{n+1} - B_n' S_n^-1 xt_n
//   backward  x_n = S_n^-1 (xt_n - B_n x_{n+1}),  V_n = S_n^-1 + C_n V_{n+1} C_n',  C_n <- -C_n V_{n+1}
// S_n^-1 is parked in V and C_n in C between the two passes, so nothing else is allocated.
// Sequential in T (T=1e5, D=32: ~1 s); the parallel-in-T block cyclic reduction is the next step.
#include "common.cuh"
#include "spd.cuh"

#define GMC_THREADS 512

__device__ __forceinline__ void gmc_load(double *dst, int ld, const double *src, int D) {
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) dst[(e / D) * ld + (e % D)] = src[e];
}

__global__ void __launch_bounds__(GMC_THREADS, 1)
gmc_block_banded_kernel(const double *A, const double *B, const double *y, int64_t T, int D,
                        double *V, double *C, double *x, double *logdet, int *flag) {
    extern __shared__ double sm[];
    const int ld = D + 1, ldg = 2 * D + 1, t = threadIdx.x, nt = blockDim.x;
    double *G = sm;                         // [D][2D+1] Gauss-Jordan tile
    double *rowk = G + (size_t)D * ldg;     // [2D]
    double *colk = rowk + 2 * D;            // [D]
    double *piv = colk + D;                 // [D]
    double *scal = piv + D;                 // [8]
    double *Sn = scal + 8;                  // [D][ld] current pivot block / V_{n+1} in the backward pass
    double *Bn = Sn + (size_t)D * ld;       // [D][ld]
    double *Cn = Bn + (size_t)D * ld;       // [D][ld]
    double *Tm = Cn + (size_t)D * ld;       // [D][ld]
    double *xt = Tm + (size_t)D * ld;       // [D]
    double *wv = xt + D;                    // [D]
    const int64_t chain = blockIdx.x;
    A += chain * T * D * D; V += chain * T * D * D; y += chain * T * D; x += chain * T * D;
    B += chain * (T - 1) * D * D; C += chain * (T - 1) * D * D;
#define VINV(i, j) G[(i) * ldg + D + (j)]
    double ldsum = 0.0;
    gmc_load(Sn, ld, A, D);
    for (int i = t; i < D; i += nt) xt[i] = y[i];
    __syncthreads();
    // ---- forward ----
    for (int64_t n = 0; n < T; ++n) {
        for (int e = t; e < D * 2 * D; e += nt) {
            const int i = e / (2 * D), j = e - i * 2 * D;
            G[i * ldg + j] = j < D ? Sn[i * ld + j] : (j - D == i ? 1.0 : 0.0);
        }
        if (n < T - 1) gmc_load(Bn, ld, B + n * D * D, D);
        for (int i = t; i < D; i += nt) x[n * D + i] = xt[i];          // park xt_n for the backward pass
        spd_cta_inverse_gj<0>(G, rowk, colk, piv, D, scal, flag);
        ldsum += scal[0];
        for (int e = t; e < D * D; e += nt) V[n * D * D + e] = VINV(e / D, e % D);
        if (n == T - 1) break;
        for (int e = t; e < D * D; e += nt) {
            const int i = e / D, j = e - i * D;
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += VINV(i, k) * Bn[k * ld + j];
            Cn[i * ld + j] = s;
            C[n * D * D + e] = s;
        }
        for (int i = t; i < D; i += nt) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += VINV(i, k) * xt[k];
            wv[i] = s;
        }
        __syncthreads();
        for (int e = t; e < D * D; e += nt) {
            const int i = e / D, j = e - i * D;
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += Bn[k * ld + i] * Cn[k * ld + j];
            Tm[i * ld + j] = s;
        }
        for (int i = t; i < D; i += nt) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += Bn[k * ld + i] * wv[k];
            const double v = y[(n + 1) * D + i] - s;
            xt[i] = v;
        }
        __syncthreads();
        for (int e = t; e < D * D; e += nt) {
            const int i = e / D, j = e - i * D;
            Sn[i * ld + j] = A[(n + 1) * D * D + e] - 0.5 * (Tm[i * ld + j] + Tm[j * ld + i]);
        }
        __syncthreads();
    }
    if (t == 0) logdet[chain] = ldsum;
    // ---- backward: G still holds S_{T-1}^-1 = V_{T-1} ----
    __syncthreads();
    for (int e = t; e < D * D; e += nt) Sn[(e / D) * ld + (e % D)] = VINV(e / D, e % D);      // V_{n+1}
    for (int i = t; i < D; i += nt) {
        double s = 0.0;
        for (int k = 0; k < D; ++k) s += VINV(i, k) * xt[k];
        wv[i] = s;                                                                             // x_{T-1}
        x[(T - 1) * D + i] = s;
    }
    __syncthreads();
    double *Vi = G;     // reuse: [D][ld] S_n^-1 of the current step
    for (int64_t n = T - 2; n >= 0; --n) {
        gmc_load(Vi, ld, V + n * D * D, D);
        gmc_load(Cn, ld, C + n * D * D, D);
        gmc_load(Bn, ld, B + n * D * D, D);
        for (int i = t; i < D; i += nt) xt[i] = x[n * D + i];
        __syncthreads();
        // r = xt - B_n x_{n+1};  Tm = C_n V_{n+1}
        for (int i = t; i < D; i += nt) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += Bn[i * ld + k] * wv[k];
            xt[i] -= s;
        }
        for (int e = t; e < D * D; e += nt) {
            const int i = e / D, j = e - i * D;
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += Cn[i * ld + k] * Sn[k * ld + j];
            Tm[i * ld + j] = s;
            C[n * D * D + e] = -s;
        }
        __syncthreads();
        for (int i = t; i < D; i += nt) {
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += Vi[i * ld + k] * xt[k];
            x[n * D + i] = s;
            rowk[i] = s;                                   // becomes x_{n+1} of the next step
        }
        for (int e = t; e < D * D; e += nt) {
            const int i = e / D, j = e - i * D;
            double s = 0.0;
            for (int k = 0; k < D; ++k) s += Tm[i * ld + k] * Cn[j * ld + k];
            Bn[i * ld + j] = Vi[i * ld + j] + s;           // V_n before symmetrisation (Bn is free now)
        }
        __syncthreads();
        for (int e = t; e < D * D; e += nt) {
            const int i = e / D, j = e - i * D;
            const double v = 0.5 * (Bn[i * ld + j] + Bn[j * ld + i]);
            Sn[i * ld + j] = v;
            V[n * D * D + e] = v;
        }
        for (int i = t; i < D; i += nt) wv[i] = rowk[i];
        __syncthreads();
    }
#undef VINV
}

extern "C" int bpk_block_banded_solve(const double *A, const double *B, const double *y,
                                      int64_t batch, int64_t T, int D,
                                      double *V, double *C, double *x, double *logdet, int check) {
    BPK_REQUIRE_INIT();
    if (D < 1 || D > BPK_MAXDIM || T < 1 || batch < 0)
        return bpk_set_error(BPK_EINVAL, "bpk_block_banded_solve: bad shape (D=%d, T=%lld)", D, (long long)T);
    if (batch == 0) return BPK_OK;
    size_t smem = ((size_t)D * (2 * D + 1) + 4 * (size_t)D + 8 + 4 * (size_t)D * (D + 1) + 2 * (size_t)D) * sizeof(double);
    if (smem > (48u << 10))
        BPK_CUDA(cudaFuncSetAttribute(gmc_block_banded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = D * D >= GMC_THREADS ? GMC_THREADS : ((D * 2 * D + 31) / 32) * 32;
    if (threads < 64) threads = 64;
    if (threads > GMC_THREADS) threads = GMC_THREADS;
    BPK_LAUNCH(gmc_block_banded_kernel, (unsigned)batch, threads, smem, A, B, y, T, D, V, C, x, logdet, g_bpk.d_flag);
    if (check) return bpk_check_flag(BPK_ENOTSPD);
    return BPK_OK;
}
