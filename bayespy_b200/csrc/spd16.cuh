// spd16.cuh — one 16 x 16 SPD system per THREAD: inverse, solve and log-determinant, arithmetic in registers.
//
// Used by the masked factor-model sweep (pca_masked.cu): every data column n has its own precision
//   Lam_n = diag(alpha) + tau sum_m mask[m,n] <w_m w_m^T>,
// and the reference factors, inverts and solves each of them with three SciPy calls per column
// (linalg.py:31-63 chol, :66-171 chol_solve, :174-207 chol_inv, :209-223 chol_logdet, driven by
// gaussian.py:672-706).  Here one thread owns one column.  The packed upper triangle (136 doubles) does not fit a
// thread's registers, so the matrix is split into 8 x 8 blocks [P Q; Q^T R] and processed by block Cholesky:
//   P = L L^T,  V = L^-1 Q,  S = R - V^T V = M M^T                       (block factor [L 0; V^T M])
//   x:  y1 = L^-1 f1, y2 = M^-1 (f2 - V^T y1), x2 = M^-T y2, x1 = L^-T (y1 - V x2);   f^T Lam^-1 f = |y1|^2 + |y2|^2
//   Cov22 = S^-1,  Cov12 = -L^-T (V S^-1),  Cov11 = L^-T (I + V S^-1 V^T) L^-1,   log det = 2 sum log diag(L), diag(M)
// Only 8 x 8 triangles (36 doubles) and a few 8-vectors are live at any time; everything else stays in the column's
// scratch rows, which the accessor maps to memory (coalesced across threads: row-major [row][column]).
// All loops have compile-time bounds and are fully unrolled, so the small arrays are registers.
//
// The same code compiles for the host (tests/test_spd16.py validates it against NumPy through a tiny C wrapper).
#pragma once
#if defined(__CUDACC__)
#define SPD16_HD __host__ __device__ __forceinline__
#if defined(__CUDA_ARCH__)
// keeps the compiler from hoisting the next phase's loads over the current phase (register pressure)
#define SPD16_FENCE() asm volatile("" ::: "memory")
#else
#define SPD16_FENCE()
#endif
#else
#define SPD16_FENCE()
#define SPD16_HD inline
#include <math.h>
#endif

// scratch rows of one column
#define SPD16_NPACK 136          // rows 0..135: packed upper triangle of Lam (in) / Cov + x x^T (out)
#define SPD16_PHI 136            // rows 136..151: phi0 (in) / x (out)
#define SPD16_LSAVE 152          // rows 152..187: saved L
#define SPD16_ID1 188            // rows 188..195: 1 / diag(L)
#define SPD16_G 196              // rows 196..231: G = I + V S^-1 V^T (lower packed)
#define SPD16_ROWS 232

// packed index of element (i, j), i <= j < 16, row-major upper triangle
SPD16_HD constexpr int spd16_pu(int i, int j) { return i * 16 - (i * (i - 1)) / 2 + (j - i); }
// lower-packed index of (i, j), i >= j, inside an 8 x 8 triangle
SPD16_HD constexpr int spd16_lp(int i, int j) { return (i * (i + 1)) / 2 + j; }

// in-place lower Cholesky of a lower-packed SPD 8 x 8; id[i] = 1 / L(i,i); ld += log det; bad |= non-positive pivot
SPD16_HD void spd8_chol(double (&L)[36], double (&id)[8], double &ld, bool &bad) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double d = L[spd16_lp(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[spd16_lp(j, k)] * L[spd16_lp(j, k)];
        if (!(d > 0.0)) bad = true;
        ld += log(d);
        const double r = 1.0 / sqrt(d);
        id[j] = r;
        L[spd16_lp(j, j)] = d * r;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) {
            double s = L[spd16_lp(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[spd16_lp(i, k)] * L[spd16_lp(j, k)];
            L[spd16_lp(i, j)] = s * r;
        }
    }
}
// v <- L^-1 v
SPD16_HD void spd8_fwd(const double (&L)[36], const double (&id)[8], double (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double s = v[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[spd16_lp(i, k)] * v[k];
        v[i] = s * id[i];
    }
}
// v <- L^-T v
SPD16_HD void spd8_bwd(const double (&L)[36], const double (&id)[8], double (&v)[8]) {
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        double s = v[i];
#pragma unroll
        for (int k = i + 1; k < 8; ++k) s -= L[spd16_lp(k, i)] * v[k];
        v[i] = s * id[i];
    }
}
// L <- L^-1 (lower triangular, in place)
SPD16_HD void spd8_triinv(double (&L)[36], const double (&id)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < i; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k) s += L[spd16_lp(i, k)] * (k == j ? id[j] : L[spd16_lp(k, j)]);
            L[spd16_lp(i, j)] = -s * id[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) L[spd16_lp(i, i)] = id[i];
}
// K <- K^T K (lower-packed symmetric result, in place), K lower triangular
SPD16_HD void spd8_ata(double (&K)[36]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double kii = K[spd16_lp(i, i)];
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double s = kii * K[spd16_lp(i, j)];
#pragma unroll
            for (int k = i + 1; k < 8; ++k) s += K[spd16_lp(k, i)] * K[spd16_lp(k, j)];
            K[spd16_lp(i, j)] = s;
        }
    }
}

// Acc: double ld(int row) const;  void st(int row, double v);   rows as defined above.
// On return rows 0..135 hold Cov + x x^T (packed upper), rows 136..151 hold x; q = phi^T Lam^-1 phi; ld = log det Lam.
// Returns false if a pivot was not positive (matrix not SPD; outputs are then garbage).
template <class Acc>
SPD16_HD bool spd16_solve_inverse(Acc &a, double &q_out, double &ld_out) {
    bool bad = false;
    double q = 0.0, ld = 0.0;
    double x1[8], x2[8];
    {
        double L[36], id1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) L[spd16_lp(i, j)] = a.ld(spd16_pu(j, i));
        spd8_chol(L, id1, ld, bad);
        // y1 = L^-1 phi1
#pragma unroll
        for (int r = 0; r < 8; ++r) x1[r] = a.ld(SPD16_PHI + r);
        spd8_fwd(L, id1, x1);
#pragma unroll
        for (int r = 0; r < 8; ++r) q += x1[r] * x1[r];
        // V = L^-1 Q (column by column, in place) and w = phi2 - V^T y1
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r] = a.ld(spd16_pu(r, 8 + c));
            spd8_fwd(L, id1, v);
            double w = a.ld(SPD16_PHI + 8 + c);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                a.st(spd16_pu(r, 8 + c), v[r]);
                w -= v[r] * x1[r];
            }
            x2[c] = w;
            SPD16_FENCE();
        }
#pragma unroll
        for (int e = 0; e < 36; ++e) a.st(SPD16_LSAVE + e, L[e]);
#pragma unroll
        for (int r = 0; r < 8; ++r) a.st(SPD16_ID1 + r, id1[r]);
    }
    SPD16_FENCE();
    {
        // S = R - V^T V, S = M M^T
        double S[36], id2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) S[spd16_lp(i, j)] = a.ld(spd16_pu(8 + j, 8 + i));
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double vr[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) vr[c] = a.ld(spd16_pu(r, 8 + c));
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) S[spd16_lp(i, j)] -= vr[i] * vr[j];
        }
        spd8_chol(S, id2, ld, bad);
        // y2 = M^-1 w, x2 = M^-T y2
        spd8_fwd(S, id2, x2);
#pragma unroll
        for (int c = 0; c < 8; ++c) q += x2[c] * x2[c];
        spd8_bwd(S, id2, x2);
        // t = y1 - V x2 (x1 := t; the back-substitution with L follows once L is reloaded)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double s = x1[r];
#pragma unroll
            for (int c = 0; c < 8; ++c) s -= a.ld(spd16_pu(r, 8 + c)) * x2[c];
            x1[r] = s;
        }
        // S^-1 = M^-T M^-1
        spd8_triinv(S, id2);
        spd8_ata(S);
        // U = V S^-1 row by row (in place over V) and G = I + U V^T (lower packed, to scratch)
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            double vb[8], ub[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) vb[c] = a.ld(spd16_pu(b, 8 + c));
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double s = 0.0;
#pragma unroll
                for (int d = 0; d < 8; ++d) s += (c >= d ? S[spd16_lp(c, d)] : S[spd16_lp(d, c)]) * vb[d];
                ub[c] = s;
            }
#pragma unroll
            for (int aa = 0; aa < b; ++aa) {
                double g = 0.0;
#pragma unroll
                for (int c = 0; c < 8; ++c) g += a.ld(spd16_pu(aa, 8 + c)) * vb[c];      // row aa already holds U
                a.st(SPD16_G + spd16_lp(b, aa), g);
            }
            double gbb = 1.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) gbb += ub[c] * vb[c];
            a.st(SPD16_G + spd16_lp(b, b), gbb);
#pragma unroll
            for (int c = 0; c < 8; ++c) a.st(spd16_pu(b, 8 + c), ub[c]);
            SPD16_FENCE();
        }
        // Cov22 + x2 x2^T
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) a.st(spd16_pu(8 + j, 8 + i), S[spd16_lp(i, j)] + x2[i] * x2[j]);
#pragma unroll
        for (int c = 0; c < 8; ++c) a.st(SPD16_PHI + 8 + c, x2[c]);
    }
    SPD16_FENCE();
    {
        double L[36], id1[8];
#pragma unroll
        for (int e = 0; e < 36; ++e) L[e] = a.ld(SPD16_LSAVE + e);
#pragma unroll
        for (int r = 0; r < 8; ++r) id1[r] = a.ld(SPD16_ID1 + r);
        spd8_bwd(L, id1, x1);                       // x1 = L^-T (y1 - V x2)
#pragma unroll
        for (int r = 0; r < 8; ++r) a.st(SPD16_PHI + r, x1[r]);
        // Cov12 = -L^-T U  (+ x1 x2^T)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double u[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) u[r] = a.ld(spd16_pu(r, 8 + c));
            spd8_bwd(L, id1, u);
#pragma unroll
            for (int r = 0; r < 8; ++r) a.st(spd16_pu(r, 8 + c), x1[r] * x2[c] - u[r]);
            SPD16_FENCE();
        }
        // Cov11 = K^T G K, K = L^-1  (+ x1 x1^T)
        spd8_triinv(L, id1);
        double G[36];
#pragma unroll
        for (int e = 0; e < 36; ++e) G[e] = a.ld(SPD16_G + e);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            double t[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double s = 0.0;
#pragma unroll
                for (int d = b; d < 8; ++d) s += (c >= d ? G[spd16_lp(c, d)] : G[spd16_lp(d, c)]) * L[spd16_lp(d, b)];
                t[c] = s;
            }
#pragma unroll
            for (int aa = 0; aa <= b; ++aa) {
                double s = 0.0;
#pragma unroll
                for (int c = aa; c < 8; ++c) s += L[spd16_lp(c, aa)] * t[c];
                a.st(spd16_pu(aa, b), s + x1[aa] * x1[b]);
            }
        }
    }
    q_out = q;
    ld_out = ld;
    return !bad;
}
