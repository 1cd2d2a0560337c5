// gmc_bcr3.cuh — block cyclic reduction, version 3: one WARP per node, matrix rows in registers.
//
// Same algorithm, workspace and outputs as the CTA-per-node kernels of gmc.cu (gmc_bcr_elim / _keep / _back:
// linalg.py:468-575 restated as a parallel-in-time elimination with selected inversion).  Those spend their time in
// block-wide barriers (a 32-step Gauss-Jordan with two __syncthreads per pivot, scalar shared-memory dot products):
// 22 ms per smoother pass at T = 1e5, D = 32, for 85 GFLOP of work (2.3 ms on the fp64 pipe).  Here a warp owns a
// node and the only synchronisation is __syncwarp.  Every D x D product runs on the fp64 tensor pipe
// (mma.sync.m8n8k4: A fragments from a row-major tile, B fragments from a tile holding B transposed, both
// bank-conflict free at a pitch of DP + 4) — a first version with lane-owns-row FMA products and 16-byte broadcast
// reads of the other operand was bound by shared-memory bandwidth (512 LDS.128 per product: 7.3 ms for level 0 at
// T = 1e5 against 11.3 ms of the CTA-per-node kernels).  The pivot inverse keeps lane r = row r in registers
// (symmetric sweep operator, pivot column exchanged through shared memory).  D <= 32, padded to DP = 8, 16 or 32
// with an identity / zero border.
#pragma once
#include "common.cuh"

#define BW_WARPS 4                 // warps (= nodes) per CTA

template <int DP> struct BwTile { static constexpr int LD = DP + 4; };      // 16-byte aligned rows; fragment loads conflict free

// global (row-major D x D) -> tile, optionally transposed; the border up to DP is zero
template <int DP>
__device__ __forceinline__ void bw_load(double *tile, const double *__restrict__ src, int D, int lane, bool transpose) {
    constexpr int LD = BwTile<DP>::LD;
    __syncwarp();                             // every lane is done with the previous contents of the tile
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
__device__ __forceinline__ void bw_dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// acc += A B with A a row-major tile ([m][k]) and B given TRANSPOSED (Bt[n][k] = B[k][n]); acc in the accumulator
// layout of m8n8k4: tile (i, j), lane (gr, tg) holds C[i*8+gr][j*8+2tg], C[i*8+gr][j*8+2tg+1]
template <int DP>
__device__ __forceinline__ void bw_mma(double (&acc)[DP / 8][DP / 8][2], const double *A, const double *Bt, int lane) {
    constexpr int LD = BwTile<DP>::LD, NB = DP / 8;
    const int gr = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int ks = 0; ks < DP / 4; ++ks) {
        double af[NB], bf[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) af[i] = A[(i * 8 + gr) * LD + ks * 4 + tg];
#pragma unroll
        for (int j = 0; j < NB; ++j) bf[j] = Bt[(j * 8 + gr) * LD + ks * 4 + tg];
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) bw_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
}
template <int DP>
__device__ __forceinline__ void bw_acc_zero(double (&acc)[DP / 8][DP / 8][2]) {
#pragma unroll
    for (int i = 0; i < DP / 8; ++i)
#pragma unroll
        for (int j = 0; j < DP / 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}
// tile = scale * acc   /   tile -= acc
template <int DP, bool SUB>
__device__ __forceinline__ void bw_acc_put(double *tile, const double (&acc)[DP / 8][DP / 8][2], int lane, double scale) {
    constexpr int LD = BwTile<DP>::LD;
    const int gr = lane >> 2, tg = lane & 3;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < DP / 8; ++i)
#pragma unroll
        for (int j = 0; j < DP / 8; ++j) {
            double2 *p = reinterpret_cast<double2 *>(tile + (i * 8 + gr) * LD + j * 8 + 2 * tg);
            if (SUB) { double2 v = *p; v.x -= acc[i][j][0]; v.y -= acc[i][j][1]; *p = v; }
            else *p = make_double2(scale * acc[i][j][0], scale * acc[i][j][1]);
        }
    __syncwarp();
}
// lane r: sum_k tile[r][k] * v[k]  (v: shared-memory vector, zero padded to DP); lanes >= DP return 0
template <int DP>
__device__ __forceinline__ double bw_rowdot(const double *tile, const double *v, int lane) {
    constexpr int LD = BwTile<DP>::LD;
    const double2 *row = reinterpret_cast<const double2 *>(tile + (lane < DP ? lane : 0) * LD);
    const double2 *p = reinterpret_cast<const double2 *>(v);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < DP / 2; ++k) {
        const double2 a = row[k], b = p[k];
        s0 = fma(a.x, b.x, s0);
        s1 = fma(a.y, b.y, s1);
    }
    return lane < DP ? s0 + s1 : 0.0;
}
template <int DP>
__device__ __forceinline__ void bw_vec_load(double *vec, const double *__restrict__ src, int D, int lane) {
    __syncwarp();
    if (lane < DP) vec[lane] = lane < D ? src[lane] : 0.0;
    __syncwarp();
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
    {
        double r[DP];
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            // symmetrised pivot block; identity on the padded border
            const double v = lane < DP ? 0.5 * (TA[lane * LD + k] + TA[k * LD + lane]) : 0.0;
            r[k] = (lane < D && k < D) ? v : ((lane == k && lane < DP) ? 1.0 : 0.0);
        }
        bw_vec_load<DP>(vec, a.x + j * D, D, lane);
        int bad = 0;
        const double ldet = bw_spd_inverse<DP>(r, col, lane, bad);
        if (bad && lane == 0) atomicOr(a.flag, BPK_FLAG_NOTSPD);
        if (lane == 0) a.ldnode[j] = ldet;
        __syncwarp();
        const double vj = bw_dot<DP>(r, vec);
        if (lane < D) a.x[j * D + lane] = vj;
        bw_row_put<DP>(TA, r, lane);                              // TA = Ainv_j (stays for the products)
    }
    bw_store<DP>(Vj, TA, D, lane, false, 1.0);
    if (base_only) return;
    const bool hasr = j + a.s < a.T;
    double acc[DP / 8][DP / 8][2];
    bw_acc_zero<DP>(acc);
    bw_load<DP>(TB, bcr_P(a, j - a.s), D, lane, false);           // Bt = P_ij  (B = P_ij^T)
    bw_mma<DP>(acc, TA, TB, lane);
    bw_acc_put<DP, false>(TB, acc, lane, 1.0);
    bw_store<DP>(a.G1 + j * D * D, TB, D, lane, false, 1.0);
    if (hasr) {
        bw_acc_zero<DP>(acc);
        bw_load<DP>(TB, bcr_P(a, j), D, lane, true);              // Bt = P_jk^T
        bw_mma<DP>(acc, TA, TB, lane);
        bw_acc_put<DP, false>(TB, acc, lane, 1.0);
        bw_store<DP>(a.G2 + j * D * D, TB, D, lane, false, 1.0);
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
    double acc[DP / 8][DP / 8][2];
    bw_acc_zero<DP>(acc);
    double xs = 0.0;
    if (hasl) {
        bw_load<DP>(TA, bcr_P(a, l), D, lane, true);              // A = P_li^T
        bw_vec_load<DP>(vec, a.x + l * D, D, lane);
        xs += bw_rowdot<DP>(TA, vec, lane);
        bw_load<DP>(TB, a.G2 + l * D * D, D, lane, true);         // Bt = G2_l^T
        bw_mma<DP>(acc, TA, TB, lane);
    }
    if (hasr) {
        bw_load<DP>(TA, bcr_P(a, i), D, lane, false);             // A = P_ir
        bw_vec_load<DP>(vec, a.x + rn * D, D, lane);
        xs += bw_rowdot<DP>(TA, vec, lane);
        bw_load<DP>(TB, a.G1 + rn * D * D, D, lane, true);        // Bt = G1_r^T
        bw_mma<DP>(acc, TA, TB, lane);
        if (hasn) {
            double u[DP / 8][DP / 8][2];
            bw_acc_zero<DP>(u);
            bw_load<DP>(TB, a.G2 + rn * D * D, D, lane, true);    // Bt = G2_r^T
            bw_mma<DP>(u, TA, TB, lane);
            bw_acc_put<DP, false>(TB, u, lane, -1.0);
            bw_store<DP>(a.lev + (a.off_next + i / (2 * a.s)) * D * D, TB, D, lane, false, 1.0);
        }
    }
    if (lane < D) a.x[i * D + lane] -= xs;
    double *Vi = a.V + i * D * D;
    bw_load<DP>(TA, Vi, D, lane, false);
    bw_acc_put<DP, true>(TA, acc, lane, 1.0);                     // A_i - (...)
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
    double s[DP / 8][DP / 8][2], acc[DP / 8][DP / 8][2];
    double xs = 0.0;
    // ---- S_ji ----
    bw_acc_zero<DP>(s);
    bw_load<DP>(TA, a.G1 + j * D * D, D, lane, false);            // TA = G1 (stays)
    bw_vec_load<DP>(vec, a.x + i * D, D, lane);
    xs += bw_rowdot<DP>(TA, vec, lane);
    bw_load<DP>(TB, a.V + i * D * D, D, lane, false);             // Bt = S_ii (symmetric)
    bw_mma<DP>(s, TA, TB, lane);
    if (hask) {
        bw_load<DP>(TC, a.G2 + j * D * D, D, lane, false);        // TC = G2 (stays)
        bw_vec_load<DP>(vec, a.x + k * D, D, lane);
        xs += bw_rowdot<DP>(TC, vec, lane);
        bw_load<DP>(TB, Sik_g, D, lane, false);                   // Bt = S_ik  (B = S_ik^T)
        bw_mma<DP>(s, TC, TB, lane);
    }
    if (lane < D) a.x[j * D + lane] -= xs;
    bw_acc_put<DP, false>(TB, s, lane, -1.0);                     // TB = S_ji
    bw_store<DP>(Cf_i, TB, D, lane, true, 1.0);                   // S_ij = S_ji^T
    bw_acc_zero<DP>(acc);
    bw_mma<DP>(acc, TB, TA, lane);                                // S_ji G1^T   (Bt = G1)
    // ---- S_jk ----
    if (hask) {
        bw_acc_zero<DP>(s);
        bw_load<DP>(TB, Sik_g, D, lane, true);                    // Bt = S_ik^T
        bw_mma<DP>(s, TA, TB, lane);
        bw_load<DP>(TB, a.V + k * D * D, D, lane, false);         // Bt = S_kk
        bw_mma<DP>(s, TC, TB, lane);
        bw_acc_put<DP, false>(TB, s, lane, -1.0);                 // TB = S_jk
        bw_store<DP>(Cf_j, TB, D, lane, false, 1.0);
        bw_mma<DP>(acc, TB, TC, lane);                            // + S_jk G2^T  (Bt = G2)
    }
    // ---- S_jj ----
    double *Vj = a.V + j * D * D;
    bw_load<DP>(TB, Vj, D, lane, false);                          // Ainv_j
    bw_acc_put<DP, true>(TB, acc, lane, 1.0);
    bw_store_sym<DP>(Vj, TB, D, lane);
}
