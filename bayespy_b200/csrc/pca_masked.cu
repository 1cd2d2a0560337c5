// pca_masked.cu — the factor-model sweep with MISSING VALUES (per-column precision), fused: nothing of size
// (N, K, K) is ever stored.
//
//   y[m,n] ~ N(w_m . x_n, 1/tau) observed only where mask[m,n];   q(x_n) = N(x_n | Cov_n phi_n, Cov_n)
//   Lam_n = diag(a_x) + tau sum_m mask[m,n] <w_m w_m^T>,   phi_n = a_x mu_x + tau sum_m mask[m,n] y[m,n] <w_m>
//
// Reference path being replaced, per VB sweep (SURVEY.md 3.3, 8d variant (ii)): dot.py:581 builds the (N,K) and
// (N,K,K) messages, gaussian.py:672-706 calls linalg.chol / chol_solve / chol_inv / chol_logdet, each a Python loop
// of SciPy calls over the N columns (linalg.py:50-59, 111-146, 185-195), stores <x x^T> as (N,K,K), and the parents'
// messages re-contract it with the mask (dot.py:581, node.py:650).  150-180 us per column on a CPU core.
//
// Here, per chunk of columns (a fixed-size scratch of 232 rows x chunk, reused by every chunk — it never scales
// with N and lives in the 126 MB L2), three kernels:
//   build    Lam_n (packed upper triangle, 136 rows) and phi_n (16 rows) for every column of the chunk as ONE GEMM on
//            the fp64 tensor pipe with the bit-packed mask as an operand:  [136+16 rows] x [64 m] x [columns]
//   inverse  one THREAD per column: block Cholesky on 8 x 8 register tiles (spd16.cuh) -> x_n, Cov_n + x_n x_n^T,
//            phi_n . x_n, log det Lam_n, in place in the scratch rows
//   stats    the plate sums every other node needs, again as one masked GEMM:
//            S_xx[m] = sum_n mask[m,n] <x_n x_n^T>,  S_yx[m] = sum_n mask[m,n] y[m,n] x_n,  sum_n <x x^T>, sum_n x_n,
//            sum_n phi.x, sum_n log det;  and the coalesced store of X.
// 38 + 38 DMMA.8x8x4 per column on the tensor pipe (the symmetric half only: 136 of 256 entries) + ~3 kflop per
// column of register-resident DFMA; algorithmic HBM traffic M*8 (y) + M (byte mask) + K*8 (x out) = 704 B/col.
#include "common.cuh"
#include "spd16.cuh"
#include <stdlib.h>

#define PM_MP 64
#define PM_KP 16
#define PM_NT 19                         // row tiles of the scratch matrix: 17 of Lam / <xx^T>, 2 of phi / x
#define PM_TILE 128                      // columns per CTA tile
#define PM_WARPS PM_NT
#define PM_THREADS (PM_WARPS * 32)
#define PM_NPART (PM_MP * (SPD16_NPACK + PM_KP) + SPD16_NPACK + PM_KP + 2)     // per-CTA partial statistics

__device__ __forceinline__ void pm_dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// (i, j) of packed row p (i <= j)
__device__ __forceinline__ void pm_unpack(int p, int &i, int &j) {
    int ii = 0, r = p;
    while (r >= PM_KP - ii) { r -= PM_KP - ii; ++ii; }
    i = ii; j = ii + r;
}

// 64-bit observation words of the tile's columns: bit m of bits[c] = mask[m][n0 + c]
__device__ __forceinline__ void pm_pack_mask(const uint8_t *__restrict__ mask, int64_t M, int64_t N, int64_t n0,
                                             unsigned long long *bits) {
    for (int c = threadIdx.x; c < PM_TILE; c += blockDim.x) bits[c] = 0ull;
    __syncthreads();
    // 4 threads per column, 16 rows each; byte loads are coalesced along the column axis
    for (int e = threadIdx.x; e < PM_TILE * 4; e += blockDim.x) {
        const int c = e % PM_TILE, q = e / PM_TILE;
        const int64_t n = n0 + c;
        unsigned long long wbits = 0ull;
        if (n < N) {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int64_t m = q * 16 + r;
                if (m < M && mask[m * N + n]) wbits |= 1ull << (q * 16 + r);
            }
        }
        if (wbits) atomicOr(&bits[c], wbits);
    }
    __syncthreads();
}

#define PM_LDY 132                        // pitch of the staged Y tile (= 4 mod 16: conflict-free fragment loads)

// Y[0:64, n0:n0+128] -> shared memory (zero beyond M / N); every load is independent, so the whole tile is in flight
__device__ __forceinline__ void pm_stage_y(const double *__restrict__ Y, int64_t M, int64_t N, int64_t n0, double *sY) {
    for (int e = threadIdx.x; e < PM_MP * PM_TILE; e += blockDim.x) {
        const int m = e / PM_TILE, c = e - m * PM_TILE;
        const int64_t n = n0 + c;
        sY[m * PM_LDY + c] = (m < M && n < N) ? __ldg(Y + (int64_t)m * N + n) : 0.0;
    }
}

struct PmArgs {
    const double *Y;
    const uint8_t *mask;
    int64_t M, N;
    int K;
    const double *W, *WW, *alpha, *amu;
    double tau;
    double *S;                 // scratch [SPD16_ROWS][chp]
    int64_t chp;               // scratch pitch (columns of a chunk, multiple of PM_TILE)
    int64_t c0, nc;            // this chunk: columns [c0, c0 + nc)
    double *X, *g;
    double *partial;           // [grid][PM_NPART] running per-CTA statistics
    int *flag;
};

// ---- build: rows 0..151 of the scratch for every column of the chunk ------------------------------------------
__global__ void __launch_bounds__(PM_THREADS, 1) pmask_build_kernel(PmArgs a) {
    __shared__ unsigned long long bits[PM_TILE];
    extern __shared__ __align__(16) double pm_smem[];
    double *sY = pm_smem;                              // [64][PM_LDY]: the tile's observations (operand of the phi rows)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gr = lane >> 2, tg = lane & 3;
    const int M = (int)a.M, K = a.K;
    // A fragments of this warp's row tile (constant over the sweep): T[p][m] = tau <ww^T>[m][ij(p)]  or  tau <w>[m][k]
    double af[16];
    const int p = w * 8 + gr;                          // scratch row of this lane's accumulator row
    double addv = 0.0;                                 // prior term of that row
    {
        int i = 0, j = 0;
        const bool is_phi = w >= 17;
        if (!is_phi) pm_unpack(p, i, j);
        const int k = p - SPD16_NPACK;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
            const int m = ks * 4 + tg;
            double v = 0.0;
            if (m < M) {
                if (is_phi) v = (k < K) ? a.tau * a.W[(int64_t)m * K + k] : 0.0;
                else if (i < K && j < K) v = a.tau * a.WW[((int64_t)m * K + i) * K + j];
            }
            af[ks] = v;
        }
        if (is_phi) addv = (k < K && a.amu) ? a.amu[k] : 0.0;
        else if (i == j) addv = (i < K) ? a.alpha[i] : 1.0;        // padded dimensions: identity precision
    }
    const int64_t ntiles = (a.nc + PM_TILE - 1) / PM_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t cl0 = tile * PM_TILE;            // first column of the tile inside the chunk
        const int64_t n0 = a.c0 + cl0;
        __syncthreads();
        pm_stage_y(a.Y, a.M, a.N, n0, sY);
        pm_pack_mask(a.mask, a.M, a.N, n0, bits);
        // 16 column tiles of 8, four at a time (4 independent DMMA chains)
#pragma unroll 1
        for (int nq = 0; nq < 4; ++nq) {
            double acc[4][2];
            unsigned long long wd[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[q][0] = acc[q][1] = 0.0;
                wd[q] = bits[(nq * 4 + q) * 8 + gr];
            }
            if (w < 17) {
#pragma unroll
                for (int ks = 0; ks < 16; ++ks)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double b = ((wd[q] >> (ks * 4 + tg)) & 1ull) ? 1.0 : 0.0;
                        pm_dmma(acc[q][0], acc[q][1], af[ks], b);
                    }
            } else {
#pragma unroll
                for (int ks = 0; ks < 16; ++ks)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int m = ks * 4 + tg;
                        const bool on = (wd[q] >> m) & 1ull;
                        const double b = on ? sY[m * PM_LDY + (nq * 4 + q) * 8 + gr] : 0.0;
                        pm_dmma(acc[q][0], acc[q][1], af[ks], b);
                    }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t cl = cl0 + (nq * 4 + q) * 8 + 2 * tg;
                double2 v = make_double2(acc[q][0] + addv, acc[q][1] + addv);
                *reinterpret_cast<double2 *>(a.S + (int64_t)p * a.chp + cl) = v;
            }
        }
    }
}

// ---- inverse: one thread per column -----------------------------------------------------------------------------
struct PmColAcc {
    double *base;              // &S[0][col]
    int64_t pitch;
    __device__ __forceinline__ double ld(int row) const { return base[(int64_t)row * pitch]; }
    __device__ __forceinline__ void st(int row, double v) { base[(int64_t)row * pitch] = v; }
};

__device__ __forceinline__ void pmask_inverse_body(const PmArgs &a) {
    const int64_t cl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cl >= a.nc) return;
    PmColAcc acc{a.S + cl, a.chp};
    double q, ld;
    const bool ok = spd16_solve_inverse(acc, q, ld);
    if (!ok) atomicOr(a.flag, BPK_FLAG_NOTSPD);
    // per-column scalars for the bound, parked in two scratch rows that are free again
    acc.st(SPD16_G, q);
    acc.st(SPD16_G + 1, ld);
}
// two register budgets: 255 registers / 8 warps per SM, or 128 registers / 16 warps per SM (more spills to L1, twice
// the threads to hide the latency of the dependent chains); BPK_PMASK_INV_REGS=128 selects the latter
__global__ void __launch_bounds__(128, 2) pmask_inverse_kernel(PmArgs a) { pmask_inverse_body(a); }
__global__ void __launch_bounds__(128, 4) pmask_inverse128_kernel(PmArgs a) { pmask_inverse_body(a); }

// ---- stats: masked plate sums + store of X -------------------------------------------------------------------------
__global__ void __launch_bounds__(PM_THREADS, 1) pmask_stats_kernel(PmArgs a, int first_chunk) {
    __shared__ unsigned long long bits[PM_TILE];
    __shared__ double red[PM_WARPS][2];
    extern __shared__ __align__(16) double pm_smem[];
    double *sY = pm_smem;                              // [64][PM_LDY]: the tile's observations (operand of the S_yx rows)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gr = lane >> 2, tg = lane & 3;
    const int K = a.K;
    const bool is_x = w >= 17;                         // row tiles 17, 18 of the scratch hold x (S_yx), the others <xx^T>
    double acc[8][2];                                  // [m tile][..]: rows m = mt*8 + gr, columns p = w*8 + 2tg + {0,1}
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
    double colsum = 0.0;                               // sum over columns of scratch row w*8 + gr (this lane's tg share)
    double sq = 0.0, sld = 0.0;
    const int64_t ntiles = (a.nc + PM_TILE - 1) / PM_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t cl0 = tile * PM_TILE;
        const int64_t n0 = a.c0 + cl0;
        const double *Srow = a.S + (int64_t)(w * 8 + gr) * a.chp + cl0;       // B operand: scratch row p, columns of the tile
        // the B fragments of the first 8 k-steps are requested before anything else (they come from L2)
        double bnext[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) bnext[u] = (n0 + u * 4 + tg < a.N) ? Srow[u * 4 + tg] : 0.0;
        __syncthreads();
        pm_stage_y(a.Y, a.M, a.N, n0, sY);
        pm_pack_mask(a.mask, a.M, a.N, n0, bits);
#pragma unroll 1
        for (int kg = 0; kg < PM_TILE / 32; ++kg) {
            double bcur[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) bcur[u] = bnext[u];
            if (kg + 1 < PM_TILE / 32) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = (kg + 1) * 32 + u * 4 + tg;
                    bnext[u] = (n0 + c < a.N) ? Srow[c] : 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = kg * 32 + u * 4 + tg;
                const double b = bcur[u];
                colsum += b;
                const unsigned long long wd = bits[c];
                if (!is_x) {
#pragma unroll
                    for (int mt = 0; mt < 8; ++mt) {
                        const double av = ((wd >> (mt * 8 + gr)) & 1ull) ? 1.0 : 0.0;
                        pm_dmma(acc[mt][0], acc[mt][1], av, b);
                    }
                } else {
#pragma unroll
                    for (int mt = 0; mt < 8; ++mt) {
                        const int m = mt * 8 + gr;
                        const double av = ((wd >> m) & 1ull) ? sY[m * PM_LDY + c] : 0.0;
                        pm_dmma(acc[mt][0], acc[mt][1], av, b);
                    }
                }
            }
        }
        // X[n][k] (coalesced), g[n], and the per-column scalars
        for (int e = threadIdx.x; e < PM_TILE * PM_KP; e += blockDim.x) {
            const int c = e >> 4, k = e & 15;
            const int64_t n = n0 + c;
            if (n < a.N && k < K) a.X[n * K + k] = a.S[(int64_t)(SPD16_PHI + k) * a.chp + cl0 + c];
        }
        for (int c = threadIdx.x; c < PM_TILE; c += blockDim.x) {
            const int64_t n = n0 + c;
            if (n < a.N) {
                const double q = a.S[(int64_t)SPD16_G * a.chp + cl0 + c], ld = a.S[(int64_t)(SPD16_G + 1) * a.chp + cl0 + c];
                sq += q;
                sld += ld;
                if (a.g) a.g[n] = -0.5 * q + 0.5 * ld;
            }
        }
    }
    // fold into this CTA's running partial (fixed order: deterministic)
    double *part = a.partial + (size_t)blockIdx.x * PM_NPART;
    const int ncol = SPD16_NPACK + PM_KP;               // 152 scratch rows = columns of the [64][152] statistics block
#pragma unroll
    for (int mt = 0; mt < 8; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            double *dst = part + (size_t)(mt * 8 + gr) * ncol + w * 8 + 2 * tg + j;
            *dst = first_chunk ? acc[mt][j] : *dst + acc[mt][j];
        }
    colsum += __shfl_xor_sync(0xffffffffu, colsum, 1);
    colsum += __shfl_xor_sync(0xffffffffu, colsum, 2);
    if (tg == 0) {
        double *dst = part + (size_t)PM_MP * ncol + w * 8 + gr;
        *dst = first_chunk ? colsum : *dst + colsum;
    }
    sq = warp_sum(sq);
    sld = warp_sum(sld);
    if (lane == 0) { red[w][0] = sq; red[w][1] = sld; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int ww = 0; ww < PM_WARPS; ++ww) { s0 += red[ww][0]; s1 += red[ww][1]; }
        double *dst = part + (size_t)PM_MP * ncol + ncol;
        dst[0] = first_chunk ? s0 : dst[0] + s0;
        dst[1] = first_chunk ? s1 : dst[1] + s1;
    }
}

// =====================================================================================
// v3 of the two GEMM kernels.  v2 gave a warp ONE row tile of the output, so every DMMA needed its own mask-bit
// extraction (shift, and, select): issue-bound at a third of the tensor pipe.  Here a warp owns one tile of the
// OTHER dimension and all 19 row tiles: the mask fragment of a k-step is built once and feeds 19 DMMAs, the second
// operand of each comes from shared memory (conflict-free pitch), one LDS per DMMA.
//   build3: warp w <-> column tile w (8 columns); T = tau [<ww^T> | <w>] (152 x 64) staged once per launch
//   stats3: warp w <-> (row tile m = w & 7, half of the k-steps w >> 3); the scratch tile <xx^T> | x (152 rows) of 64
//           columns at a time is staged in shared memory
// =====================================================================================
#define PM3_WARPS 16
#define PM3_THREADS (PM3_WARPS * 32)
#define PM3_LDT 68                       // pitch of T and of the staged scratch sub-tile (= 4 mod 16)
#define PM3_SUB 64                       // columns per staged scratch sub-tile (stats3)

__global__ void __launch_bounds__(PM3_THREADS, 1) pmask_build3_kernel(PmArgs a) {
    __shared__ unsigned long long bits[PM_TILE];
    __shared__ double sAdd[PM_NT * 8];
    extern __shared__ __align__(16) double pm_smem[];
    double *sY = pm_smem;                              // [64][PM_LDY]
    double *sT = sY + PM_MP * PM_LDY;                  // [152][PM3_LDT]: row p, column m
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gr = lane >> 2, tg = lane & 3;
    const int M = (int)a.M, K = a.K;
    for (int e = threadIdx.x; e < PM_NT * 8 * PM_MP; e += blockDim.x) {
        const int p = e / PM_MP, m = e - p * PM_MP;
        double v = 0.0;
        if (m < M) {
            if (p >= SPD16_NPACK) { const int k = p - SPD16_NPACK; v = (k < K) ? a.tau * a.W[(int64_t)m * K + k] : 0.0; }
            else { int i, j; pm_unpack(p, i, j); if (i < K && j < K) v = a.tau * a.WW[((int64_t)m * K + i) * K + j]; }
        }
        sT[p * PM3_LDT + m] = v;
    }
    for (int p = threadIdx.x; p < PM_NT * 8; p += blockDim.x) {
        double v = 0.0;
        if (p >= SPD16_NPACK) { const int k = p - SPD16_NPACK; v = (k < K && a.amu) ? a.amu[k] : 0.0; }
        else { int i, j; pm_unpack(p, i, j); if (i == j) v = (i < K) ? a.alpha[i] : 1.0; }
        sAdd[p] = v;
    }
    const int64_t ntiles = (a.nc + PM_TILE - 1) / PM_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t cl0 = tile * PM_TILE, n0 = a.c0 + cl0;
        __syncthreads();
        pm_stage_y(a.Y, a.M, a.N, n0, sY);
        pm_pack_mask(a.mask, a.M, a.N, n0, bits);
        double acc[PM_NT][2];
#pragma unroll
        for (int pt = 0; pt < PM_NT; ++pt) acc[pt][0] = acc[pt][1] = 0.0;
        const int c = w * 8 + gr;                       // this lane's column of the A fragment
        const unsigned long long wd = bits[c];
#pragma unroll 2
        for (int ks = 0; ks < 16; ++ks) {
            const int m = ks * 4 + tg;
            const bool on = (wd >> m) & 1ull;
            const double am = on ? 1.0 : 0.0;
            const double ay = on ? sY[m * PM_LDY + c] : 0.0;
            const double *bt = sT + gr * PM3_LDT + m;    // B[k = tg -> m][n = gr -> p]
#pragma unroll
            for (int pt = 0; pt < 17; ++pt) pm_dmma(acc[pt][0], acc[pt][1], am, bt[pt * 8 * PM3_LDT]);
            pm_dmma(acc[17][0], acc[17][1], ay, bt[17 * 8 * PM3_LDT]);
            pm_dmma(acc[18][0], acc[18][1], ay, bt[18 * 8 * PM3_LDT]);
        }
        // D[n = gr][p = pt*8 + 2tg + j] -> scratch row p, column w*8 + gr
#pragma unroll
        for (int pt = 0; pt < PM_NT; ++pt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int p = pt * 8 + 2 * tg + j;
                a.S[(int64_t)p * a.chp + cl0 + c] = acc[pt][j] + sAdd[p];
            }
    }
}

__global__ void __launch_bounds__(PM3_THREADS, 1) pmask_stats3_kernel(PmArgs a, int first_chunk) {
    __shared__ unsigned long long bits[PM_TILE];
    __shared__ double red[PM3_WARPS][2];
    extern __shared__ __align__(16) double pm_smem[];
    double *sY = pm_smem;                              // [64][PM_LDY]
    double *sS = sY + PM_MP * PM_LDY;                  // [152][PM3_LDT]: scratch rows of a 64-column sub-tile
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gr = lane >> 2, tg = lane & 3;
    const int K = a.K;
    const int mt = w & 7, hf = w >> 3;
    double acc[PM_NT][2];                              // rows m = mt*8 + gr, columns p = pt*8 + 2tg + {0,1}; this warp's half of the k-steps
#pragma unroll
    for (int pt = 0; pt < PM_NT; ++pt) acc[pt][0] = acc[pt][1] = 0.0;
    double rowsum[10];                                 // scratch rows w, w+16, ...: sum over the columns (lane 0 holds it)
#pragma unroll
    for (int i = 0; i < 10; ++i) rowsum[i] = 0.0;
    double sq = 0.0, sld = 0.0;
    const int64_t ntiles = (a.nc + PM_TILE - 1) / PM_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t cl0 = tile * PM_TILE, n0 = a.c0 + cl0;
        __syncthreads();
        pm_stage_y(a.Y, a.M, a.N, n0, sY);
        pm_pack_mask(a.mask, a.M, a.N, n0, bits);
        for (int sub = 0; sub < PM_TILE / PM3_SUB; ++sub) {
            const int cs = sub * PM3_SUB;
            __syncthreads();
            for (int e = threadIdx.x; e < PM_NT * 8 * PM3_SUB; e += blockDim.x) {
                const int p = e / PM3_SUB, cc = e - p * PM3_SUB;
                sS[p * PM3_LDT + cc] = (n0 + cs + cc < a.N) ? a.S[(int64_t)p * a.chp + cl0 + cs + cc] : 0.0;
            }
            __syncthreads();
            // column sums of the staged rows (sum_n <xx^T>, sum_n x)
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const int p = w + 16 * i;
                if (p < PM_NT * 8) {
                    double v = sS[p * PM3_LDT + lane] + sS[p * PM3_LDT + 32 + lane];
                    rowsum[i] += warp_sum(v);
                }
            }
            // this warp's 8 k-steps of the sub-tile
#pragma unroll 2
            for (int kk = 0; kk < 8; ++kk) {
                const int ks = hf * 8 + kk;
                const int cc = ks * 4 + tg;              // column inside the sub-tile (A: col = tg, B: k = tg)
                const unsigned long long wd = bits[cs + cc];
                const int m = mt * 8 + gr;
                const bool on = (wd >> m) & 1ull;
                const double am = on ? 1.0 : 0.0;
                const double ay = on ? sY[m * PM_LDY + cs + cc] : 0.0;
                const double *bs = sS + gr * PM3_LDT + cc; // B[k = tg -> column][n = gr -> p]
#pragma unroll
                for (int pt = 0; pt < 17; ++pt) pm_dmma(acc[pt][0], acc[pt][1], am, bs[pt * 8 * PM3_LDT]);
                pm_dmma(acc[17][0], acc[17][1], ay, bs[17 * 8 * PM3_LDT]);
                pm_dmma(acc[18][0], acc[18][1], ay, bs[18 * 8 * PM3_LDT]);
            }
        }
        // X[n][k] (coalesced), g[n], and the per-column scalars
        for (int e = threadIdx.x; e < PM_TILE * PM_KP; e += blockDim.x) {
            const int c = e >> 4, k = e & 15;
            const int64_t n = n0 + c;
            if (n < a.N && k < K) a.X[n * K + k] = a.S[(int64_t)(SPD16_PHI + k) * a.chp + cl0 + c];
        }
        for (int c = threadIdx.x; c < PM_TILE; c += blockDim.x) {
            const int64_t n = n0 + c;
            if (n < a.N) {
                const double q = a.S[(int64_t)SPD16_G * a.chp + cl0 + c], ld = a.S[(int64_t)(SPD16_G + 1) * a.chp + cl0 + c];
                sq += q;
                sld += ld;
                if (a.g) a.g[n] = -0.5 * q + 0.5 * ld;
            }
        }
    }
    // combine the two halves of the k-steps through shared memory, then fold into this CTA's running partial
    double *part = a.partial + (size_t)blockIdx.x * PM_NPART;
    const int ncol = SPD16_NPACK + PM_KP;
    __syncthreads();
    double *sH = pm_smem;                               // [64][152] staging (reuses the tile buffers)
    if (hf == 1) {
#pragma unroll
        for (int pt = 0; pt < PM_NT; ++pt)
#pragma unroll
            for (int j = 0; j < 2; ++j) sH[(mt * 8 + gr) * ncol + pt * 8 + 2 * tg + j] = acc[pt][j];
    }
    __syncthreads();
    if (hf == 0) {
#pragma unroll
        for (int pt = 0; pt < PM_NT; ++pt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int idx = (mt * 8 + gr) * ncol + pt * 8 + 2 * tg + j;
                const double v = acc[pt][j] + sH[idx];
                part[idx] = first_chunk ? v : part[idx] + v;
            }
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const int p = w + 16 * i;
            if (p < PM_NT * 8) {
                double *dst = part + (size_t)PM_MP * ncol + p;
                *dst = first_chunk ? rowsum[i] : *dst + rowsum[i];
            }
        }
    }
    sq = warp_sum(sq);
    sld = warp_sum(sld);
    if (lane == 0) { red[w][0] = sq; red[w][1] = sld; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int ww = 0; ww < PM3_WARPS; ++ww) { s0 += red[ww][0]; s1 += red[ww][1]; }
        double *dst = part + (size_t)PM_MP * ncol + ncol;
        dst[0] = first_chunk ? s0 : dst[0] + s0;
        dst[1] = first_chunk ? s1 : dst[1] + s1;
    }
}

// partials -> caller's layout  [ S_yx (M*K) | S_xx (M*K*K, symmetric, full) | sum<xx^T> (K*K) | sum x (K) | sum phi.x | sum logdet ]
__global__ void pmask_final_kernel(const double *__restrict__ partial, int nblocks, int M, int K, double *__restrict__ stats) {
    const int ncol = SPD16_NPACK + PM_KP;
    const int total = M * K + M * K * K + K * K + K + 2;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    int src;
    if (e < M * K) { const int m = e / K, k = e - m * K; src = m * ncol + SPD16_NPACK + k; }
    else if (e < M * K + M * K * K) {
        const int r = e - M * K, m = r / (K * K), ij = r - m * K * K;
        int i = ij / K, j = ij - i * K;
        if (i > j) { const int t = i; i = j; j = t; }
        src = m * ncol + spd16_pu(i, j);
    } else if (e < M * K + M * K * K + K * K) {
        const int ij = e - M * K - M * K * K;
        int i = ij / K, j = ij - i * K;
        if (i > j) { const int t = i; i = j; j = t; }
        src = PM_MP * ncol + spd16_pu(i, j);
    } else if (e < total - 2) src = PM_MP * ncol + SPD16_NPACK + (e - (M * K + M * K * K + K * K));
    else src = PM_MP * ncol + ncol + (e - (total - 2));
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * PM_NPART + src];
    stats[e] += s;
}

static double *g_pm_scratch = nullptr;
static size_t g_pm_scratch_bytes = 0;

extern "C" int bpk_pca_xsweep_masked_fused(const double *Y, const uint8_t *mask, int64_t M, int64_t N, int K,
                                           const double *W, const double *WW, double tau,
                                           const double *alpha, const double *amu,
                                           double *X, double *g, double *stats, int check) {
    BPK_REQUIRE_INIT();
    if (M < 1 || M > PM_MP || K < 1 || K > PM_KP || N < 0)
        return bpk_set_error(BPK_EINVAL, "bpk_pca_xsweep_masked_fused: needs M <= %d and K <= %d (got M=%lld, K=%d)",
                             PM_MP, PM_KP, (long long)M, K);
    if (!Y || !mask || !W || !WW || !alpha || !X || !stats)
        return bpk_set_error(BPK_EINVAL, "bpk_pca_xsweep_masked_fused: null argument");
    if (N == 0) return BPK_OK;
    const int grid = g_bpk.sm_count;
    int tiles_per_cta = 8;       // 148 x 8 x 128 columns per chunk: 50.3 ms at N = 1e7 against 52.9 ms with 4 (profiles/r02_pmask_v2_timing.txt)
    if (const char *e = getenv("BPK_PMASK_CHUNK_TILES")) { tiles_per_cta = atoi(e); if (tiles_per_cta < 1) tiles_per_cta = 1; }
    int64_t chunk = (int64_t)grid * tiles_per_cta * PM_TILE;
    if (chunk > ((N + PM_TILE - 1) / PM_TILE) * PM_TILE) chunk = ((N + PM_TILE - 1) / PM_TILE) * PM_TILE;
    const size_t need = (size_t)SPD16_ROWS * chunk * sizeof(double) + (size_t)grid * PM_NPART * sizeof(double);
    if (need > g_pm_scratch_bytes) {
        if (g_pm_scratch) BPK_CUDA(cudaFreeAsync(g_pm_scratch, g_bpk.stream));
        g_pm_scratch = nullptr;
        g_pm_scratch_bytes = 0;
        BPK_CUDA(cudaMallocAsync((void **)&g_pm_scratch, need, g_bpk.stream));
        g_pm_scratch_bytes = need;
    }
    PmArgs a;
    a.Y = Y; a.mask = mask; a.M = M; a.N = N; a.K = K; a.W = W; a.WW = WW; a.alpha = alpha; a.amu = amu; a.tau = tau;
    a.S = g_pm_scratch; a.chp = chunk; a.X = X; a.g = g;
    a.partial = g_pm_scratch + (size_t)SPD16_ROWS * chunk;
    a.flag = g_bpk.d_flag;
    const size_t ysm = (size_t)PM_MP * PM_LDY * sizeof(double);
    BPK_CUDA(cudaFuncSetAttribute(pmask_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ysm));
    BPK_CUDA(cudaFuncSetAttribute(pmask_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ysm));
    const size_t sm3 = (size_t)(PM_MP * PM_LDY + PM_NT * 8 * PM3_LDT) * sizeof(double);
    // v3 (one warp = all 19 row tiles of a column tile, operands staged in shared memory) measured 64.2 ms against
    // v2's 52.9 ms at N = 1e7 (round 2, session 8; identical results): kept for experiments behind BPK_PMASK_V3
    const bool v2 = getenv("BPK_PMASK_V3") == nullptr;
    if (!v2) {
        BPK_CUDA(cudaFuncSetAttribute(pmask_build3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
        BPK_CUDA(cudaFuncSetAttribute(pmask_stats3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
    }
    const char *ir = getenv("BPK_PMASK_INV_REGS");
    const bool inv255 = !(ir && atoi(ir) == 128);
    int first = 1;
    for (int64_t c0 = 0; c0 < N; c0 += chunk) {
        a.c0 = c0;
        a.nc = (N - c0 < chunk) ? N - c0 : chunk;
        const int64_t ntiles = (a.nc + PM_TILE - 1) / PM_TILE;
        const int gtiles = (int)(ntiles < grid ? ntiles : grid);
        if (v2) BPK_LAUNCH(pmask_build_kernel, gtiles, PM_THREADS, ysm, a);
        else BPK_LAUNCH(pmask_build3_kernel, gtiles, PM3_THREADS, sm3, a);
        // the inverse runs on whole tiles: padded columns of the last tile hold the prior precision (SPD), results unused
        PmArgs b = a;
        b.nc = ntiles * PM_TILE;
        if (inv255) BPK_LAUNCH(pmask_inverse_kernel, (unsigned)((b.nc + 127) / 128), 128, 0, b);
        else BPK_LAUNCH(pmask_inverse128_kernel, (unsigned)((b.nc + 127) / 128), 128, 0, b);
        if (v2) BPK_LAUNCH(pmask_stats_kernel, grid, PM_THREADS, ysm, a, first);
        else BPK_LAUNCH(pmask_stats3_kernel, grid, PM3_THREADS, sm3, a, first);
        first = 0;
    }
    const int total = (int)(M * K + M * K * K + (int64_t)K * K + K + 2);
    BPK_LAUNCH(pmask_final_kernel, (total + 127) / 128, 128, 0, a.partial, grid, (int)M, K, stats);
    if (check) return bpk_check_flag(BPK_ENOTSPD);
    return BPK_OK;
}
