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
#include <stdlib.h>
#include <vector>

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

// =====================================================================================
// v2: parallel in time — block cyclic reduction with selected inversion.
// Level l (stride s = 2^l) eliminates every other active node j with the Schur complement on its two
// kept neighbours i = j - s, k = j + s:
//   forward   Ainv_j = A_j^-1,  G1_j = Ainv_j P_ij^T,  G2_j = Ainv_j P_jk,  v_j = Ainv_j y_j          (gmc_bcr_elim)
//             A_i <- A_i - P_li^T G2_l - P_ir G1_r,  y_i <- y_i - P_li^T v_l - P_ir v_r,
//             P'_{i,i+2s} = -P_ir G2_r                                                               (gmc_bcr_keep)
//   backward  x_j = v_j - G1_j x_i - G2_j x_k,
//             S_ji = -(G1_j S_ii + G2_j S_ik^T),  S_jk = -(G1_j S_ik + G2_j S_kk),
//             S_jj = Ainv_j - S_ji G1_j^T - S_jk G2_j^T     (Takahashi recurrences)                   (gmc_bcr_back)
// which yields exactly the diagonal and super-diagonal blocks of the inverse at the next finer level.
// log2(T) levels, every level = one CTA per node, all D x D block operations by all threads of the CTA;
// scatter-free (every kept node GATHERS from its eliminated neighbours), hence deterministic.  Any T.
// Workspace: G1, G2 (T blocks each) and the level couplings (< T blocks).
// =====================================================================================
struct BcrArgs {
    double *V;            // [T][D][D] working diagonal blocks -> Ainv of eliminated nodes -> S_nn
    double *x;            // [T][D]    working right-hand side -> v_j -> solution
    const double *B;      // [T-1][D][D] level-0 couplings (input)
    double *C;            // [T-1][D][D] level-0 super-diagonal blocks of the inverse (output)
    double *lev;          // level couplings / inverse couplings, level l >= 1 at lev + off[l] blocks
    double *G1, *G2;      // [T][D][D]
    double *ldnode;       // [T] log det of every pivot block
    int64_t T;
    int D;
    int64_t s;            // stride of this level
    int64_t off_this, off_next;   // block offsets of this / the next coarser level inside lev (level 0: B / C)
    int *flag;
};

// coupling between kept nodes (n, n + s) at the level with stride s
__device__ __forceinline__ const double *bcr_P(const BcrArgs &a, int64_t n) {
    return a.s == 1 ? a.B + n * a.D * a.D : a.lev + (a.off_this + n / a.s) * a.D * a.D;
}

__device__ __forceinline__ void gmc_ld(double *dst, int ld, const double *src, int D) {
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) dst[(e / D) * ld + (e % D)] = src[e];
}

#include "gmc_bcr3.cuh"

// eliminated node: invert the pivot, form G1, G2, v
__global__ void __launch_bounds__(256) gmc_bcr_elim(BcrArgs a, int base_only) {
    extern __shared__ double sm[];
    const int D = a.D, ld = D + 1, ldg = 2 * D + 1, t = threadIdx.x, nt = blockDim.x;
    double *G = sm, *rowk = G + (size_t)D * ldg, *colk = rowk + 2 * D, *piv = colk + D, *scal = piv + D;
    double *Pl = scal + 8, *Pr = Pl + (size_t)D * ld, *yv = Pr + (size_t)D * ld;
    const int64_t j = base_only ? 0 : a.s * (2 * (int64_t)blockIdx.x + 1);
    double *Vj = a.V + j * D * D;
    for (int e = t; e < D * 2 * D; e += nt) {
        const int i = e / (2 * D), c = e - i * 2 * D;
        G[i * ldg + c] = c < D ? 0.5 * (Vj[i * D + c] + Vj[c * D + i]) : (c - D == i ? 1.0 : 0.0);
    }
    const bool hasr = !base_only && (j + a.s < a.T);
    if (!base_only) {
        gmc_ld(Pl, ld, bcr_P(a, j - a.s), D);
        if (hasr) gmc_ld(Pr, ld, bcr_P(a, j), D);
    }
    for (int i = t; i < D; i += nt) yv[i] = a.x[j * D + i];
    spd_cta_inverse_gj<0>(G, rowk, colk, piv, D, scal, a.flag);
#define AINV(i, c) G[(i) * ldg + D + (c)]
    if (t == 0) a.ldnode[j] = scal[0];
    for (int e = t; e < D * D; e += nt) Vj[e] = AINV(e / D, e % D);
    for (int i = t; i < D; i += nt) {
        double s = 0.0;
        for (int k = 0; k < D; ++k) s += AINV(i, k) * yv[k];
        a.x[j * D + i] = s;
    }
    if (!base_only) {
        double *G1 = a.G1 + j * D * D, *G2 = a.G2 + j * D * D;
        for (int e = t; e < D * D; e += nt) {
            const int i = e / D, c = e - i * D;
            double s1 = 0.0, s2 = 0.0;
            for (int k = 0; k < D; ++k) {
                s1 += AINV(i, k) * Pl[c * ld + k];                 // Ainv P_ij^T
                if (hasr) s2 += AINV(i, k) * Pr[k * ld + c];       // Ainv P_jk
            }
            G1[e] = s1;
            G2[e] = s2;
        }
    }
#undef AINV
}

// kept node: gather the Schur-complement updates of its eliminated neighbours, emit the next level's coupling
__global__ void __launch_bounds__(256) gmc_bcr_keep(BcrArgs a) {
    extern __shared__ double sm[];
    const int D = a.D, ld = D + 1, t = threadIdx.x, nt = blockDim.x;
    double *Pli = sm, *Pir = Pli + (size_t)D * ld, *Ga = Pir + (size_t)D * ld, *Gb = Ga + (size_t)D * ld;
    double *Gc = Gb + (size_t)D * ld, *An = Gc + (size_t)D * ld, *vl = An + (size_t)D * ld, *vr = vl + D;
    const int64_t i = 2 * a.s * (int64_t)blockIdx.x, l = i - a.s, r = i + a.s;
    const bool hasl = l >= 0, hasr = r < a.T, hasn = r + a.s < a.T;
    if (hasl) { gmc_ld(Pli, ld, bcr_P(a, l), D); gmc_ld(Ga, ld, a.G2 + l * D * D, D); }
    if (hasr) { gmc_ld(Pir, ld, bcr_P(a, i), D); gmc_ld(Gb, ld, a.G1 + r * D * D, D); }
    if (hasn) gmc_ld(Gc, ld, a.G2 + r * D * D, D);
    for (int q = t; q < D; q += nt) {
        vl[q] = hasl ? a.x[l * D + q] : 0.0;
        vr[q] = hasr ? a.x[r * D + q] : 0.0;
    }
    __syncthreads();
    double *Vi = a.V + i * D * D;
    for (int e = t; e < D * D; e += nt) {
        const int p = e / D, c = e - p * D;
        double s = 0.0;
        if (hasl) for (int k = 0; k < D; ++k) s += Pli[k * ld + p] * Ga[k * ld + c];      // P_li^T G2_l
        if (hasr) for (int k = 0; k < D; ++k) s += Pir[p * ld + k] * Gb[k * ld + c];      // P_ir G1_r
        An[p * ld + c] = Vi[e] - s;
        if (hasn) {
            double u = 0.0;
            for (int k = 0; k < D; ++k) u += Pir[p * ld + k] * Gc[k * ld + c];            // P_ir G2_r
            a.lev[(a.off_next + i / (2 * a.s)) * D * D + e] = -u;
        }
    }
    for (int p = t; p < D; p += nt) {
        double s = 0.0;
        if (hasl) for (int k = 0; k < D; ++k) s += Pli[k * ld + p] * vl[k];
        if (hasr) for (int k = 0; k < D; ++k) s += Pir[p * ld + k] * vr[k];
        a.x[i * D + p] -= s;
    }
    __syncthreads();
    for (int e = t; e < D * D; e += nt) {
        const int p = e / D, c = e - p * D;
        Vi[e] = 0.5 * (An[p * ld + c] + An[c * ld + p]);
    }
}

// back-substitution for an eliminated node: solution and the blocks of the inverse that touch it
__global__ void __launch_bounds__(256) gmc_bcr_back(BcrArgs a) {
    extern __shared__ double sm[];
    const int D = a.D, ld = D + 1, t = threadIdx.x, nt = blockDim.x;
    double *G1 = sm, *G2 = G1 + (size_t)D * ld, *Sii = G2 + (size_t)D * ld, *Skk = Sii + (size_t)D * ld;
    double *Sik = Skk + (size_t)D * ld, *Sji = Sik + (size_t)D * ld, *Sjk = Sji + (size_t)D * ld, *Aj = Sjk + (size_t)D * ld;
    double *xi = Aj + (size_t)D * ld, *xk = xi + D;
    const int64_t j = a.s * (2 * (int64_t)blockIdx.x + 1), i = j - a.s, k = j + a.s;
    const bool hask = k < a.T;
    gmc_ld(G1, ld, a.G1 + j * D * D, D);
    gmc_ld(Sii, ld, a.V + i * D * D, D);
    gmc_ld(Aj, ld, a.V + j * D * D, D);
    if (hask) {
        gmc_ld(G2, ld, a.G2 + j * D * D, D);
        gmc_ld(Skk, ld, a.V + k * D * D, D);
        gmc_ld(Sik, ld, a.lev + (a.off_next + i / (2 * a.s)) * D * D, D);      // S_{i,k} from the coarser level
    }
    for (int q = t; q < D; q += nt) { xi[q] = a.x[i * D + q]; xk[q] = hask ? a.x[k * D + q] : 0.0; }
    __syncthreads();
    for (int e = t; e < D * D; e += nt) {
        const int p = e / D, c = e - p * D;
        double s1 = 0.0, s2 = 0.0;
        for (int q = 0; q < D; ++q) {
            s1 += G1[p * ld + q] * Sii[q * ld + c];
            if (hask) {
                s1 += G2[p * ld + q] * Sik[c * ld + q];                        // G2 S_ik^T
                s2 += G1[p * ld + q] * Sik[q * ld + c] + G2[p * ld + q] * Skk[q * ld + c];
            }
        }
        Sji[p * ld + c] = -s1;
        Sjk[p * ld + c] = -s2;
    }
    for (int p = t; p < D; p += nt) {
        double s = 0.0;
        for (int q = 0; q < D; ++q) s += G1[p * ld + q] * xi[q] + (hask ? G2[p * ld + q] * xk[q] : 0.0);
        a.x[j * D + p] -= s;
    }
    __syncthreads();
    double *Cf_i = a.s == 1 ? a.C + i * D * D : a.lev + (a.off_this + i / a.s) * D * D;     // S_{i,j}
    double *Cf_j = a.s == 1 ? a.C + j * D * D : a.lev + (a.off_this + j / a.s) * D * D;     // S_{j,k}
    for (int e = t; e < D * D; e += nt) {
        const int p = e / D, c = e - p * D;
        Cf_i[e] = Sji[c * ld + p];
        if (hask) Cf_j[e] = Sjk[p * ld + c];
        double s = 0.0;
        for (int q = 0; q < D; ++q) s += Sji[p * ld + q] * G1[c * ld + q] + (hask ? Sjk[p * ld + q] * G2[c * ld + q] : 0.0);
        Skk[p * ld + c] = Aj[p * ld + c] - s;  // S_jj before symmetrisation (Skk tile is free: all reads are done)
    }
    __syncthreads();
    for (int e = t; e < D * D; e += nt) {
        const int p = e / D, c = e - p * D;
        a.V[j * D * D + e] = 0.5 * (Skk[p * ld + c] + Skk[c * ld + p]);
    }
}

__global__ void gmc_sum_kernel(const double *v, int64_t n, double *out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
        *out = tot;
    }
}

static int gmc_bcr_solve(const double *A, const double *B, const double *y, int64_t T, int D,
                         double *V, double *C, double *x, double *logdet) {
    const size_t blk = (size_t)D * D;
    // level offsets (in blocks) inside the coupling workspace: level l >= 1 has ceil(T / 2^l) - 1 couplings
    std::vector<int64_t> off;
    int64_t tot = 0;
    off.push_back(0);                       // level 0 lives in B / C
    for (int64_t s = 2; s < 2 * T; s *= 2) {
        off.push_back(tot);
        int64_t nact = (T + s - 1) / s;
        tot += nact > 1 ? nact - 1 : 1;
    }
    // Workspace from the library's persistent stream-ordered scratch (2.4 GB at T = 1e5, D = 32): a cudaMallocAsync /
    // cudaFreeAsync pair per call made the pool re-map physical memory whenever other allocations had split the block
    // in between — stalls of 90-480 ms inside X.update() (round 2, session 11).
    size_t wsd = ((size_t)tot + 2 * (size_t)T) * blk + (size_t)T;
    double *ws = bpk_scratch(wsd * sizeof(double));
    if (!ws) return bpk_set_error(BPK_ECUDA, "bpk_block_banded_solve: workspace of %zu bytes", wsd * sizeof(double));
    BcrArgs a;
    a.V = V; a.x = x; a.B = B; a.C = C; a.T = T; a.D = D; a.flag = g_bpk.d_flag;
    a.lev = ws; a.G1 = ws + (size_t)tot * blk; a.G2 = a.G1 + (size_t)T * blk; a.ldnode = a.G2 + (size_t)T * blk;
    BPK_CUDA(cudaMemcpyAsync(V, A, (size_t)T * blk * sizeof(double), cudaMemcpyDeviceToDevice, g_bpk.stream));
    BPK_CUDA(cudaMemcpyAsync(x, y, (size_t)T * D * sizeof(double), cudaMemcpyDeviceToDevice, g_bpk.stream));
    const int ld = D + 1;
    size_t sm_elim = ((size_t)D * (2 * D + 1) + 4 * (size_t)D + 8 + 2 * (size_t)D * ld + D) * sizeof(double);
    size_t sm_keep = (6 * (size_t)D * ld + 2 * (size_t)D) * sizeof(double);
    size_t sm_back = (8 * (size_t)D * ld + 2 * (size_t)D) * sizeof(double);
    BPK_CUDA(cudaFuncSetAttribute(gmc_bcr_elim, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_elim));
    BPK_CUDA(cudaFuncSetAttribute(gmc_bcr_keep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_keep));
    BPK_CUDA(cudaFuncSetAttribute(gmc_bcr_back, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_back));
    const int threads = 256;
    std::vector<int64_t> strides;
    int lvl = 0;
    // version 3 (one warp per node, rows in registers: gmc_bcr3.cuh) for D <= 32; BPK_GMC_BCR_V2=1 keeps the CTA-per-node kernels
    const bool v3 = D <= 32 && !getenv("BPK_GMC_BCR_V2");
    const int DP = D <= 8 ? 8 : (D <= 16 ? 16 : 32);
    const size_t w3 = (size_t)(2 * DP * (DP + 4) + 2 * DP) * sizeof(double), w3b = (size_t)(3 * DP * (DP + 4) + 2 * DP) * sizeof(double);
    const size_t sm3_ek = BW_WARPS * w3, sm3_b = BW_WARPS * w3b;
#define BCR3_DISPATCH(KERNEL, NODES, SMEM, ...)                                                                        \
    do {                                                                                                             \
        const unsigned g3 = (unsigned)(((NODES) + BW_WARPS - 1) / BW_WARPS);                                          \
        if (DP == 8) {                                                                                               \
            BPK_CUDA(cudaFuncSetAttribute(KERNEL<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));     \
            BPK_LAUNCH(KERNEL<8>, g3, BW_WARPS * 32, SMEM, __VA_ARGS__);                                              \
        } else if (DP == 16) {                                                                                       \
            BPK_CUDA(cudaFuncSetAttribute(KERNEL<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));    \
            BPK_LAUNCH(KERNEL<16>, g3, BW_WARPS * 32, SMEM, __VA_ARGS__);                                             \
        } else {                                                                                                     \
            BPK_CUDA(cudaFuncSetAttribute(KERNEL<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM)));    \
            BPK_LAUNCH(KERNEL<32>, g3, BW_WARPS * 32, SMEM, __VA_ARGS__);                                             \
        }                                                                                                            \
    } while (0)
    for (int64_t s = 1; (T + s - 1) / s > 1; s *= 2, ++lvl) {
        const int64_t nact = (T + s - 1) / s, nel = nact / 2, nkeep = (nact + 1) / 2;
        a.s = s; a.off_this = off[lvl]; a.off_next = off[lvl + 1];
        if (v3) {
            BCR3_DISPATCH(gmc_bcr3_elim, nel, sm3_ek, a, 0, nel);
            BCR3_DISPATCH(gmc_bcr3_keep, nkeep, sm3_ek, a, nkeep);
        } else {
            BPK_LAUNCH(gmc_bcr_elim, (unsigned)nel, threads, sm_elim, a, 0);
            BPK_LAUNCH(gmc_bcr_keep, (unsigned)nkeep, threads, sm_keep, a);
        }
        strides.push_back(s);
    }
    a.s = 1; a.off_this = 0; a.off_next = 0;
    if (v3) BCR3_DISPATCH(gmc_bcr3_elim, (int64_t)1, sm3_ek, a, 1, (int64_t)1);      // the last node standing
    else BPK_LAUNCH(gmc_bcr_elim, 1, threads, sm_elim, a, 1);
    for (int q = (int)strides.size() - 1; q >= 0; --q) {
        const int64_t s = strides[q], nact = (T + s - 1) / s, nel = nact / 2;
        a.s = s; a.off_this = off[q]; a.off_next = off[q + 1];
        if (v3) BCR3_DISPATCH(gmc_bcr3_back, nel, sm3_b, a, nel);
        else BPK_LAUNCH(gmc_bcr_back, (unsigned)nel, threads, sm_back, a);
    }
#undef BCR3_DISPATCH
    BPK_LAUNCH(gmc_sum_kernel, 1, 1024, 0, a.ldnode, T, logdet);
    return BPK_OK;
}

extern "C" int bpk_block_banded_solve(const double *A, const double *B, const double *y,
                                      int64_t batch, int64_t T, int D,
                                      double *V, double *C, double *x, double *logdet, int check) {
    BPK_REQUIRE_INIT();
    if (D < 1 || D > BPK_MAXDIM || T < 1 || batch < 0)
        return bpk_set_error(BPK_EINVAL, "bpk_block_banded_solve: bad shape (D=%d, T=%lld)", D, (long long)T);
    if (batch == 0) return BPK_OK;
    {   // long chains: parallel in time (BPK_GMC_BCR_MIN overrides the switch-over length, 0 = never)
        const char *e = getenv("BPK_GMC_BCR_MIN");
        const int64_t tmin = e ? atoll(e) : 8;
        if (tmin > 0 && T >= tmin) {
            for (int64_t b = 0; b < batch; ++b) {
                int rc = gmc_bcr_solve(A + b * T * D * D, B + b * (T - 1) * D * D, y + b * T * D, T, D,
                                       V + b * T * D * D, C + b * (T - 1) * D * D, x + b * T * D, logdet + b);
                if (rc) return rc;
            }
            if (check) return bpk_check_flag(BPK_ENOTSPD);
            return BPK_OK;
        }
    }
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
