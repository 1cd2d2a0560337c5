// gmc_bcr3.cuh — block cyclic reduction, version 3: one WARP per node, matrix rows in registers.
//
// Same algorithm, workspace and outputs as the CTA-per-node kernels of gmc.cu (gmc_bcr_elim / _keep / _back:
// linalg.py:468-575 restated as a parallel-in-time elimination with selected inversion).  Those spend their time in
// block-wide barriers (a 32-step Gauss-Jordan with two __syncthreads per pivot, scalar shared-memory dot products):
// 22 ms per smoother pass at T = 1e5, D = 32, for 85 GFLOP of work (2.3 ms on the fp64 pipe).  Here a warp owns a
// node: lane r keeps row r of the matrix it works on in registers, the other operand of every D x D product is a
// shared-memory tile read with 16-byte broadcast loads (one row of the tile per output column), and the only
// synchronisation is __syncwarp.  D <= 32 (padded to DP = 8, 16 or 32 with an identity / zero border).
#pragma once
#include "common.cuh"

#define BW_WARPS 4                 // warps (= nodes) per CTA

template <int DP> struct BwTile { static constexpr int LD = DP + 2; };      // 16-byte aligned rows, even pitch

// global (row-major D x D) -> tile, optionally transposed; the border up to DP is zero
template <int DP>
__device__ __forceinline__ void bw_load(double *tile, const double *__restrict__ src, int D, int lane, bool transpose) {
    constexpr int LD = BwTile<DP>::LD;
    if (D == DP) {
#pragma unroll 4
        for (int e = lane; e < DP * DP; e += 32) {
            const int r = e / DP, c = e % DP;
            tile[transpose ? c * LD + r : r * LD + c] = __ldg(src + e);
        }
    } else {
        for (int e = lane; e < DP * DP; e += 32) {
            const int r = e / DP, c = e % DP;
            const double v = (r < D && c < D) ? __ldg(src + r * D + c) : 0.0;
            tile[transpose ? c * LD + r : r * LD + c] = v;
        }
    }
    __syncwarp();
}
template <int DP>
__device__ __forceinline__ void bw_zero(double *tile, int lane) {
    constexpr int LD = BwTile<DP>::LD;
    for (int e = lane; e < DP * LD; e += 32) tile[e] = 0.0;
    __syncwarp();
}
// tile -> global (row-major D x D), optionally transposed, scaled
template <int DP>
__device__ __forceinline__ void bw_store(double *__restrict__ dst, const double *tile, int D, int lane, bool transpose, double scale) {
    constexpr int LD = BwTile<DP>::LD;
    __syncwarp();
    for (int e = lane; e < D * D; e += 32) {
        const int r = e / D, c = e - r * D;
        dst[e] = scale * tile[transpose ? c * LD + r : r * LD + c];
    }
    __syncwarp();
}
// symmetrised store: dst = (tile + tile^T) / 2
template <int DP>
__device__ __forceinline__ void bw_store_sym(double *__restrict__ dst, const double *tile, int D, int lane) {
    constexpr int LD = BwTile<DP>::LD;
    __syncwarp();
    for (int e = lane; e < D * D; e += 32) {
        const int r = e / D, c = e - r * D;
        dst[e] = 0.5 * (tile[r * LD + c] + tile[c * LD + r]);
    }
    __syncwarp();
}
// lane r <- row r of the tile (lanes >= DP get zeros)
template <int DP>
__device__ __forceinline__ void bw_row_get(double (&a)[DP], const double *tile, int lane) {
    constexpr int LD = BwTile<DP>::LD;
    const double2 *row = reinterpret_cast<const double2 *>(tile + (lane < DP ? lane : 0) * LD);
#pragma unroll
    for (int k = 0; k < DP / 2; ++k) {
        const double2 v = row[k];
        a[2 * k] = lane < DP ? v.x : 0.0;
        a[2 * k + 1] = lane < DP ? v.y : 0.0;
    }
}
template <int DP>
__device__ __forceinline__ void bw_row_put(double *tile, const double (&a)[DP], int lane) {
    constexpr int LD = BwTile<DP>::LD;
    __syncwarp();
    if (lane < DP) {
        double2 *row = reinterpret_cast<double2 *>(tile + lane * LD);
#pragma unroll
        for (int k = 0; k < DP / 2; ++k) row[k] = make_double2(a[2 * k], a[2 * k + 1]);
    }
    __syncwarp();
}
// out[c] (+)= sum_k a[k] * Bt[c][k]   — row r of (A B) with B given TRANSPOSED in the tile (Bt[c][k] = B[k][c])
template <int DP, bool ACC>
__device__ __forceinline__ void bw_mul(double (&out)[DP], const double (&a)[DP], const double *Bt) {
    constexpr int LD = BwTile<DP>::LD;
#pragma unroll
    for (int c = 0; c < DP; ++c) {
        const double2 *row = reinterpret_cast<const double2 *>(Bt + c * LD);
        double s0 = ACC ? out[c] : 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < DP / 2; ++k) {
            const double2 b = row[k];
            s0 = fma(a[2 * k], b.x, s0);
            s1 = fma(a[2 * k + 1], b.y, s1);
        }
        out[c] = s0 + s1;
    }
}
// sum_k a[k] * v[k] with v a shared-memory vector (16-byte aligned, zero padded to DP)
template <int DP>
__device__ __forceinline__ double bw_dot(const double (&a)[DP], const double *v) {
    const double2 *p = reinterpret_cast<const double2 *>(v);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < DP / 2; ++k) {
        const double2 b = p[k];
        s0 = fma(a[2 * k], b.x, s0);
        s1 = fma(a[2 * k + 1], b.y, s1);
    }
    return s0 + s1;
}

// In-register inverse of a symmetric positive definite matrix by the symmetric sweep operator: lane r holds row r.
//   SWP(k):  b_kk = -1/a_kk,  b_ik = b_ki = a_ik / a_kk,  b_ij = a_ij - a_ik a_kj / a_kk     (stays symmetric)
// after k = 0..DP-1 the matrix is -A^-1.  The pivot COLUMN a_.k (= row k by symmetry) is exchanged through `col`.
// Returns sum_k log(pivot_k) = log det A; raises `bad` for a non-positive or non-finite pivot.
template <int DP>
__device__ __forceinline__ double bw_spd_inverse(double (&a)[DP], double *col, int lane, int &bad) {
    double mypiv = 1.0;                       // lane k remembers pivot k: one log per lane at the end
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        __syncwarp();
        if (lane < DP) col[lane] = a[k];
        __syncwarp();
        const double d = col[k];
        if (!(d > 0.0) || !isfinite(d)) bad = 1;
        mypiv = lane == k ? d : mypiv;
        const double invd = 1.0 / d;
        const bool me = lane == k;
        const double f = a[k] * invd;
        const double2 *cp = reinterpret_cast<const double2 *>(col);
#pragma unroll
        for (int j2 = 0; j2 < DP / 2; ++j2) {
            const double2 cj = cp[j2];
            // other rows: a_ij - (a_ik / d) a_kj;   pivot row: a_kj / d
            a[2 * j2] = me ? cj.x * invd : fma(-f, cj.x, a[2 * j2]);
            a[2 * j2 + 1] = me ? cj.y * invd : fma(-f, cj.y, a[2 * j2 + 1]);
        }
        a[k] = me ? -invd : f;
    }
#pragma unroll
    for (int j = 0; j < DP; ++j) a[j] = -a[j];
    return warp_sum(log(mypiv));
}

// ---- eliminated node: Ainv_j, G1_j = Ainv_j P_ij^T, G2_j = Ainv_j P_jk, v_j = Ainv_j y_j ----------------------------
template <int DP>
__global__ void __launch_bounds__(BW_WARPS * 32, 2) gmc_bcr3_elim(BcrArgs a, int base_only, int64_t nnodes) {
    constexpr int LD = BwTile<DP>::LD;
    extern __shared__ __align__(16) double sm3[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, D = a.D;
    const int64_t node = (int64_t)blockIdx.x * BW_WARPS + w;
    if (node >= nnodes) return;
    double *TA = sm3 + (size_t)w * (2 * DP * LD + 2 * DP), *TB = TA + DP * LD, *vec = TB + DP * LD, *col = vec + DP;
    const int64_t j = base_only ? 0 : a.s * (2 * node + 1);
    double *Vj = a.V + j * D * D;
    bw_load<DP>(TA, Vj, D, lane, false);
    double r[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        // symmetrised pivot block; identity on the padded border
        const double v = lane < DP ? 0.5 * (TA[lane * LD + k] + TA[k * LD + lane]) : 0.0;
        r[k] = (lane < D && k < D) ? v : ((lane == k && lane < DP) ? 1.0 : 0.0);
    }
    if (lane < DP) vec[lane] = lane < D ? a.x[j * D + lane] : 0.0;
    int bad = 0;
    const double ldet = bw_spd_inverse<DP>(r, col, lane, bad);
    if (bad && lane == 0) atomicOr(a.flag, BPK_FLAG_NOTSPD);
    if (lane == 0) a.ldnode[j] = ldet;
    __syncwarp();
    const double vj = bw_dot<DP>(r, vec);
    if (lane < D) a.x[j * D + lane] = vj;
    bw_row_put<DP>(TA, r, lane);
    bw_store<DP>(Vj, TA, D, lane, false, 1.0);
    if (base_only) return;
    const bool hasr = j + a.s < a.T;
    double o[DP];
    bw_load<DP>(TB, bcr_P(a, j - a.s), D, lane, false);          // Bt = P_ij  (B = P_ij^T)
    bw_mul<DP, false>(o, r, TB);
    bw_row_put<DP>(TA, o, lane);
    bw_store<DP>(a.G1 + j * D * D, TA, D, lane, false, 1.0);
    if (hasr) {
        bw_load<DP>(TB, bcr_P(a, j), D, lane, true);             // Bt = P_jk^T
        bw_mul<DP, false>(o, r, TB);
        bw_row_put<DP>(TA, o, lane);
        bw_store<DP>(a.G2 + j * D * D, TA, D, lane, false, 1.0);
    } else {
        for (int e = lane; e < D * D; e += 32) a.G2[j * D * D + e] = 0.0;
    }
}

// ---- kept node: A_i <- A_i - P_li^T G2_l - P_ir G1_r,  y_i likewise,  P'_{i,i+2s} = -P_ir G2_r ------------------------
template <int DP>
__global__ void __launch_bounds__(BW_WARPS * 32, 2) gmc_bcr3_keep(BcrArgs a, int64_t nnodes) {
    constexpr int LD = BwTile<DP>::LD;
    extern __shared__ __align__(16) double sm3[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, D = a.D;
    const int64_t node = (int64_t)blockIdx.x * BW_WARPS + w;
    if (node >= nnodes) return;
    double *TA = sm3 + (size_t)w * (2 * DP * LD + 2 * DP), *TB = TA + DP * LD, *vec = TB + DP * LD;
    const int64_t i = 2 * a.s * node, l = i - a.s, rn = i + a.s;
    const bool hasl = l >= 0, hasr = rn < a.T, hasn = rn + a.s < a.T;
    double acc[DP], p[DP];
#pragma unroll
    for (int c = 0; c < DP; ++c) acc[c] = 0.0;
    double xs = 0.0;
    if (hasl) {
        bw_load<DP>(TA, bcr_P(a, l), D, lane, true);             // rows of P_li^T
        bw_row_get<DP>(p, TA, lane);
        bw_load<DP>(TB, a.G2 + l * D * D, D, lane, true);        // Bt = G2_l^T
        bw_mul<DP, true>(acc, p, TB);
        __syncwarp();
        if (lane < DP) vec[lane] = lane < D ? a.x[l * D + lane] : 0.0;
        __syncwarp();
        xs += bw_dot<DP>(p, vec);
    }
    if (hasr) {
        bw_load<DP>(TA, bcr_P(a, i), D, lane, false);            // rows of P_ir
        bw_row_get<DP>(p, TA, lane);
        bw_load<DP>(TB, a.G1 + rn * D * D, D, lane, true);       // Bt = G1_r^T
        bw_mul<DP, true>(acc, p, TB);
        __syncwarp();
        if (lane < DP) vec[lane] = lane < D ? a.x[rn * D + lane] : 0.0;
        __syncwarp();
        xs += bw_dot<DP>(p, vec);
        if (hasn) {
            double u[DP];
            bw_load<DP>(TB, a.G2 + rn * D * D, D, lane, true);   // Bt = G2_r^T
            bw_mul<DP, false>(u, p, TB);
            bw_row_put<DP>(TA, u, lane);
            bw_store<DP>(a.lev + (a.off_next + i / (2 * a.s)) * D * D, TA, D, lane, false, -1.0);
        }
    }
    if (lane < D) a.x[i * D + lane] -= xs;
    double *Vi = a.V + i * D * D;
    bw_load<DP>(TA, Vi, D, lane, false);
    bw_row_get<DP>(p, TA, lane);
#pragma unroll
    for (int c = 0; c < DP; ++c) p[c] -= acc[c];
    bw_row_put<DP>(TA, p, lane);
    bw_store_sym<DP>(Vi, TA, D, lane);
}

// ---- back-substitution for an eliminated node (Takahashi recurrences) ------------------------------------------------
//   x_j = v_j - G1 x_i - G2 x_k;  S_ji = -(G1 S_ii + G2 S_ik^T);  S_jk = -(G1 S_ik + G2 S_kk);
//   S_jj = Ainv_j - S_ji G1^T - S_jk G2^T  (symmetrised)
template <int DP>
__global__ void __launch_bounds__(BW_WARPS * 32, 2) gmc_bcr3_back(BcrArgs a, int64_t nnodes) {
    constexpr int LD = BwTile<DP>::LD;
    extern __shared__ __align__(16) double sm3[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, D = a.D;
    const int64_t node = (int64_t)blockIdx.x * BW_WARPS + w;
    if (node >= nnodes) return;
    double *TA = sm3 + (size_t)w * (3 * DP * LD + 2 * DP), *TB = TA + DP * LD, *TC = TB + DP * LD, *vec = TC + DP * LD;
    const int64_t j = a.s * (2 * node + 1), i = j - a.s, k = j + a.s;
    const bool hask = k < a.T;
    const double *Sik_g = a.lev + (a.off_next + i / (2 * a.s)) * D * D;       // S_{i,k} from the coarser level
    double *Cf_i = a.s == 1 ? a.C + i * D * D : a.lev + (a.off_this + i / a.s) * D * D;     // S_{i,j}
    double *Cf_j = a.s == 1 ? a.C + j * D * D : a.lev + (a.off_this + j / a.s) * D * D;     // S_{j,k}
    double g[DP], s[DP], acc[DP];
    double xs = 0.0;
    // ---- S_ji ----
    bw_load<DP>(TA, a.G1 + j * D * D, D, lane, false);
    bw_row_get<DP>(g, TA, lane);                                  // g = row of G1 (TA keeps G1 for the S_jj product)
    bw_load<DP>(TB, a.V + i * D * D, D, lane, false);             // Bt = S_ii^T = S_ii
    bw_mul<DP, false>(s, g, TB);
    __syncwarp();
    if (lane < DP) vec[lane] = lane < D ? a.x[i * D + lane] : 0.0;
    __syncwarp();
    xs += bw_dot<DP>(g, vec);
    if (hask) {
        bw_load<DP>(TC, a.G2 + j * D * D, D, lane, false);        // TC keeps G2
        bw_row_get<DP>(g, TC, lane);                              // g = row of G2
        bw_load<DP>(TB, Sik_g, D, lane, false);                   // Bt = S_ik  (B = S_ik^T)
        bw_mul<DP, true>(s, g, TB);
        __syncwarp();
        if (lane < DP) vec[lane] = lane < D ? a.x[k * D + lane] : 0.0;
        __syncwarp();
        xs += bw_dot<DP>(g, vec);
    }
    if (lane < D) a.x[j * D + lane] -= xs;
#pragma unroll
    for (int c = 0; c < DP; ++c) s[c] = -s[c];                    // s = row of S_ji
    bw_mul<DP, false>(acc, s, TA);                                // S_ji G1^T   (Bt = G1)
    bw_row_put<DP>(TB, s, lane);
    bw_store<DP>(Cf_i, TB, D, lane, true, 1.0);                   // S_ij = S_ji^T
    // ---- S_jk ----
    if (hask) {
        bw_row_get<DP>(g, TA, lane);                              // row of G1 again
        bw_load<DP>(TB, Sik_g, D, lane, true);                    // Bt = S_ik^T
        bw_mul<DP, false>(s, g, TB);
        bw_row_get<DP>(g, TC, lane);                              // row of G2
        bw_load<DP>(TB, a.V + k * D * D, D, lane, false);         // Bt = S_kk
        bw_mul<DP, true>(s, g, TB);
#pragma unroll
        for (int c = 0; c < DP; ++c) s[c] = -s[c];                // s = row of S_jk
        bw_mul<DP, true>(acc, s, TC);                             // + S_jk G2^T  (Bt = G2)
        bw_row_put<DP>(TB, s, lane);
        bw_store<DP>(Cf_j, TB, D, lane, false, 1.0);
    }
    // ---- S_jj ----
    double *Vj = a.V + j * D * D;
    bw_load<DP>(TB, Vj, D, lane, false);                          // Ainv_j
    bw_row_get<DP>(g, TB, lane);
#pragma unroll
    for (int c = 0; c < DP; ++c) g[c] -= acc[c];
    bw_row_put<DP>(TB, g, lane);
    bw_store_sym<DP>(Vj, TB, D, lane);
}
