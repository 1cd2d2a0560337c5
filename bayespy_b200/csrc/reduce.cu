// reduce.cu — bpk_sum_multiply: sum over chosen axes of the product of up to
// four broadcast operands (the restricted einsum of bayespy/utils/misc.py:851
// and the masked plate-sum of nodes/node.py:650).  Deterministic: fixed
// reduction trees, no atomics.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

struct SmArgs {
    int n_in, nk, ns;
    int64_t n_kept, n_sum;
    int64_t kshape[BPK_MAXD], sshape[BPK_MAXD];
    int64_t kout[BPK_MAXD];
    int64_t kin[BPK_MAXIN][BPK_MAXD], sin[BPK_MAXIN][BPK_MAXD];
    const void *in[BPK_MAXIN];
    int dtype[BPK_MAXIN];
    double *out;
    double scale;
    int accumulate;
    int nsplit;
    int64_t chunk;
    double *partial;
};

__device__ __forceinline__ double sm_load(const void *p, int dtype, int64_t off) {
    return dtype == BPK_U8 ? (double)((const uint8_t *)p)[off] : ((const double *)p)[off];
}

__device__ __forceinline__ void sm_decode_kept(const SmArgs &A, int64_t o, int64_t &oo, int64_t base[BPK_MAXIN]) {
    oo = 0;
#pragma unroll
    for (int k = 0; k < BPK_MAXIN; ++k) base[k] = 0;
    int64_t rem = o;
    for (int d = A.nk - 1; d >= 0; --d) {
        int64_t q = rem / A.kshape[d];
        int64_t r = rem - q * A.kshape[d];
        rem = q;
        oo += r * A.kout[d];
#pragma unroll
        for (int k = 0; k < BPK_MAXIN; ++k)
            if (k < A.n_in) base[k] += r * A.kin[k][d];
    }
}

__device__ __forceinline__ double sm_term(const SmArgs &A, int64_t s, const int64_t base[BPK_MAXIN]) {
    int64_t off[BPK_MAXIN];
#pragma unroll
    for (int k = 0; k < BPK_MAXIN; ++k) off[k] = base[k];
    int64_t rem = s;
    for (int d = A.ns - 1; d >= 0; --d) {
        int64_t q = rem / A.sshape[d];
        int64_t r = rem - q * A.sshape[d];
        rem = q;
#pragma unroll
        for (int k = 0; k < BPK_MAXIN; ++k)
            if (k < A.n_in) off[k] += r * A.sin[k][d];
    }
    double p = sm_load(A.in[0], A.dtype[0], off[0]);
#pragma unroll
    for (int k = 1; k < BPK_MAXIN; ++k)
        if (k < A.n_in) p *= sm_load(A.in[k], A.dtype[k], off[k]);
    return p;
}

// A: one thread per kept element, serial loop over the summed space.
__global__ void __launch_bounds__(256) sm_thread_kernel(SmArgs A) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; o < A.n_kept; o += step) {
        int64_t oo, base[BPK_MAXIN];
        sm_decode_kept(A, o, oo, base);
        double acc = 0.0;
        if (A.ns == 1) {   // common case: odometer-free inner loop
            for (int64_t s = 0; s < A.n_sum; ++s) {
                double p = sm_load(A.in[0], A.dtype[0], base[0] + s * A.sin[0][0]);
#pragma unroll
                for (int k = 1; k < BPK_MAXIN; ++k)
                    if (k < A.n_in) p *= sm_load(A.in[k], A.dtype[k], base[k] + s * A.sin[k][0]);
                acc += p;
            }
        } else {
            for (int64_t s = 0; s < A.n_sum; ++s) acc += sm_term(A, s, base);
        }
        double v = A.scale * acc;
        A.out[oo] = A.accumulate ? A.out[oo] + v : v;
    }
}

// B: one block per (kept element, split); threads stride over the summed space.
__global__ void __launch_bounds__(256) sm_block_kernel(SmArgs A) {
    __shared__ double red[8];
    int64_t b = blockIdx.x;
    for (; b < A.n_kept * A.nsplit; b += gridDim.x) {
        int64_t o = b / A.nsplit;
        int sp = (int)(b - o * A.nsplit);
        int64_t oo, base[BPK_MAXIN];
        sm_decode_kept(A, o, oo, base);
        int64_t s0 = sp * A.chunk;
        int64_t s1 = s0 + A.chunk;
        if (s1 > A.n_sum) s1 = A.n_sum;
        double acc = 0.0;
        for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) acc += sm_term(A, s, base);
        acc = warp_sum(acc);
        int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        if (nw > 1) {
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[w] = acc;
            __syncthreads();
            acc = 0.0;
            if (threadIdx.x < nw) acc = red[threadIdx.x];
            if (w == 0) acc = warp_sum(acc);
        }
        if (threadIdx.x == 0) {
            if (A.nsplit == 1) {
                double v = A.scale * acc;
                A.out[oo] = A.accumulate ? A.out[oo] + v : v;
            } else {
                A.partial[b] = acc;
            }
        }
    }
}

__global__ void __launch_bounds__(256) sm_final_kernel(SmArgs A) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; o < A.n_kept; o += step) {
        int64_t oo, base[BPK_MAXIN];
        sm_decode_kept(A, o, oo, base);
        double acc = 0.0;
        for (int sp = 0; sp < A.nsplit; ++sp) acc += A.partial[o * A.nsplit + sp];
        double v = A.scale * acc;
        A.out[oo] = A.accumulate ? A.out[oo] + v : v;
    }
}

// ---- GEMM-shaped contractions on the fp64 tensor pipe -------------------------------------------------------
// When a two-operand sum_multiply collapses to  C[m,n] (+)= scale * sum_k A[m,k] B[k,n]  (one axis only in
// operand 0, one only in operand 1, one summed axis in both — e.g. <f f> = sum_ij <cc>[m,ij] <xx>[n,ij] of
// dot.py:403, the messages of dot.py:581, the W^T Lambda W style products of the Gaussian messages), it runs
// as a DMMA GEMM with arbitrary element strides: 64 x 64 tile per CTA, 4 warps of 32 x 32, k-step 16 staged in
// shared memory k-contiguous (pitch 20 = 4 mod 16: conflict-free fragment loads).  Fixed summation order.
struct GemmArgs {
    const double *A, *B;
    double *C;
    int64_t M, N, K;
    int64_t sAm, sAk, sBk, sBn, sCm, sCn;
    double scale;
    int accumulate;
    int64_t kchunk;            // split-K: CTA z contracts k in [z*kchunk, (z+1)*kchunk) into partial slab z of `part`
    double *part;              // [nsplit][M][N] (NULL: single pass straight into C)
};
#define GM_BM 64
#define GM_BN 64
#define GM_BK 16
#define GM_LD 20

__device__ __forceinline__ void gm_dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128, 2) dgemm_dmma_kernel(GemmArgs g) {
    __shared__ double As[GM_BM * GM_LD], Bs[GM_BN * GM_LD];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5, gr = lane >> 2, tg = lane & 3;
    const int wm = (w >> 1) * 32, wn = (w & 1) * 32;
    const int64_t m0 = (int64_t)blockIdx.y * GM_BM, n0 = (int64_t)blockIdx.x * GM_BN;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const bool a_kfast = g.sAk == 1 || g.sAm != 1, b_kfast = g.sBk == 1 || g.sBn != 1;
    // each thread stages 8 + 8 elements of a k-tile; the NEXT tile is fetched into registers while the
    // current one is multiplied, so the HBM / L2 latency hides behind the 64 DMMA of the tile
    constexpr int PT = GM_BM * GM_BK / 128;
    // tile element e -> (row, k): k fastest when the operand is k-contiguous, row fastest otherwise
    auto arow = [&](int e) { return a_kfast ? e / GM_BK : e % GM_BM; };
    auto acol = [&](int e) { return a_kfast ? e % GM_BK : e / GM_BM; };
    auto brow = [&](int e) { return b_kfast ? e / GM_BK : e % GM_BN; };
    auto bcol = [&](int e) { return b_kfast ? e % GM_BK : e / GM_BN; };
    double ra[PT], rb[PT];
    const int64_t kbeg = (int64_t)blockIdx.z * g.kchunk;
    const int64_t kend = (kbeg + g.kchunk < g.K) ? kbeg + g.kchunk : g.K;
    auto fetch = [&](int64_t k0) {
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            const int e = t + q * 128;
            const int64_t mm = m0 + arow(e), ka = k0 + acol(e), nn = n0 + brow(e), kb = k0 + bcol(e);
            ra[q] = (mm < g.M && ka < kend) ? g.A[mm * g.sAm + ka * g.sAk] : 0.0;
            rb[q] = (nn < g.N && kb < kend) ? g.B[kb * g.sBk + nn * g.sBn] : 0.0;
        }
    };
    fetch(kbeg);
    for (int64_t k0 = kbeg; k0 < kend; k0 += GM_BK) {
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            const int e = t + q * 128;
            As[arow(e) * GM_LD + acol(e)] = ra[q];
            Bs[brow(e) * GM_LD + bcol(e)] = rb[q];
        }
        __syncthreads();
        if (k0 + GM_BK < kend) fetch(k0 + GM_BK);
#pragma unroll
        for (int ks = 0; ks < GM_BK; ks += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[(wm + i * 8 + gr) * GM_LD + ks + tg];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn + j * 8 + gr) * GM_LD + ks + tg];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) gm_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int64_t m = m0 + wm + i * 8 + gr, n = n0 + wn + j * 8 + 2 * tg + q;
                if (m < g.M && n < g.N) {
                    if (g.part) {
                        g.part[((int64_t)blockIdx.z * g.M + m) * g.N + n] = acc[i][j][q];
                    } else {
                        double *c = g.C + m * g.sCm + n * g.sCn;
                        const double v = g.scale * acc[i][j][q];
                        *c = g.accumulate ? *c + v : v;
                    }
                }
            }
}

// split-K epilogue: C (+)= scale * sum_z part[z] in a fixed order (deterministic)
__global__ void __launch_bounds__(256) dgemm_splitk_final_kernel(GemmArgs g, int nsplit) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = g.M * g.N;
    for (; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = e / g.N, n = e - m * g.N;
        double s = 0.0;
        for (int z = 0; z < nsplit; ++z) s += g.part[(int64_t)z * total + e];
        double *c = g.C + m * g.sCm + n * g.sCn;
        const double v = g.scale * s;
        *c = g.accumulate ? *c + v : v;
    }
}

extern "C" int bpk_sum_multiply(int nd, const int64_t *shape,
                                int n_in, const void *const *in, const int *in_dtype,
                                const int64_t *in_stride,
                                double *out, const int64_t *out_stride,
                                double scale, int accumulate) {
    BPK_REQUIRE_INIT();
    if (nd < 0 || nd > BPK_MAXD) return bpk_set_error(BPK_EINVAL, "bpk_sum_multiply: nd=%d exceeds %d", nd, BPK_MAXD);
    if (n_in < 1 || n_in > BPK_MAXIN) return bpk_set_error(BPK_EINVAL, "bpk_sum_multiply: n_in=%d (max %d)", n_in, BPK_MAXIN);
    SmArgs A;
    memset(&A, 0, sizeof(A));
    A.n_in = n_in; A.out = out; A.scale = scale; A.accumulate = accumulate;
    for (int k = 0; k < n_in; ++k) { A.in[k] = in[k]; A.dtype[k] = in_dtype[k]; }
    A.n_kept = 1; A.n_sum = 1;
    bool empty = false;
    for (int d = 0; d < nd; ++d) {
        int64_t e = shape[d];
        if (e < 0) return bpk_set_error(BPK_EINVAL, "bpk_sum_multiply: negative extent");
        if (e == 0) empty = true;
        if (e == 1) continue;                 // carries no index
        if (out_stride[d] != 0) {             // kept axis
            // merge with the previous kept axis when every array is contiguous across the pair
            bool merge = A.nk > 0 && A.kout[A.nk - 1] == out_stride[d] * e;
            for (int k = 0; merge && k < n_in; ++k)
                merge = A.kin[k][A.nk - 1] == in_stride[k * nd + d] * e;
            // only adjacent original axes may merge: guaranteed because axes are visited in order
            if (merge && d > 0) {
                int j = A.nk - 1;
                A.kshape[j] *= e;
                A.kout[j] = out_stride[d];
                for (int k = 0; k < n_in; ++k) A.kin[k][j] = in_stride[k * nd + d];
            } else {
                int j = A.nk++;
                A.kshape[j] = e;
                A.kout[j] = out_stride[d];
                for (int k = 0; k < n_in; ++k) A.kin[k][j] = in_stride[k * nd + d];
            }
            A.n_kept *= e;
        } else {                              // summed axis
            bool all_bcast = true;
            for (int k = 0; k < n_in; ++k) all_bcast = all_bcast && in_stride[k * nd + d] == 0;
            if (all_bcast) { A.scale *= (double)e; continue; }   // broadcasting multiplier (misc.py:761)
            bool merge = A.ns > 0;
            for (int k = 0; merge && k < n_in; ++k)
                merge = A.sin[k][A.ns - 1] == in_stride[k * nd + d] * e;
            if (merge) {
                int j = A.ns - 1;
                A.sshape[j] *= e;
                for (int k = 0; k < n_in; ++k) A.sin[k][j] = in_stride[k * nd + d];
            } else {
                int j = A.ns++;
                A.sshape[j] = e;
                for (int k = 0; k < n_in; ++k) A.sin[k][j] = in_stride[k * nd + d];
            }
            A.n_sum *= e;
        }
    }
    if (empty) {
        // empty sum: the result is zero (or unchanged when accumulating) over the kept space;
        // an empty kept space writes nothing.
        return BPK_OK;
    }
    // GEMM-shaped: two fp64 operands, kept axes {m: operand 0 only, n: operand 1 only}, one summed axis in both
    if (n_in == 2 && A.dtype[0] == BPK_F64 && A.dtype[1] == BPK_F64 && A.nk == 2 && A.ns == 1 &&
        A.sin[0][0] != 0 && A.sin[1][0] != 0) {
        int im = -1, in_ = -1;
        for (int d = 0; d < 2; ++d) {
            if (A.kin[0][d] != 0 && A.kin[1][d] == 0) im = d;
            else if (A.kin[0][d] == 0 && A.kin[1][d] != 0) in_ = d;
        }
        if (im >= 0 && in_ >= 0 && A.kshape[im] >= 8 && A.kshape[in_] >= 8 && A.sshape[0] >= 4 &&
            (double)A.n_kept * (double)A.n_sum >= 262144.0 && !getenv("BPK_NO_GEMM")) {
            GemmArgs g;
            g.A = (const double *)A.in[0]; g.B = (const double *)A.in[1]; g.C = out;
            g.M = A.kshape[im]; g.N = A.kshape[in_]; g.K = A.sshape[0];
            g.sAm = A.kin[0][im]; g.sAk = A.sin[0][0]; g.sBk = A.sin[1][0]; g.sBn = A.kin[1][in_];
            g.sCm = A.kout[im]; g.sCn = A.kout[in_];
            g.scale = A.scale; g.accumulate = accumulate;
            dim3 grid((unsigned)((g.N + GM_BN - 1) / GM_BN), (unsigned)((g.M + GM_BM - 1) / GM_BM));
            // split-K when the output tiles cannot fill the machine and the contraction is long (e.g. the
            // plate-summed messages of dot.py:581: 256 x 1024 outputs over 1e5 time steps): slabs of k, partial
            // results to scratch, fixed-order epilogue
            g.kchunk = g.K;
            g.part = nullptr;
            int nsplit = 1;
            const int64_t ctas = (int64_t)grid.x * grid.y, want = 2 * (int64_t)g_bpk.sm_count;
            if (ctas < want && g.K >= 4096 && !getenv("BPK_GEMM_NO_SPLITK")) {
                nsplit = (int)((want + ctas - 1) / ctas);
                const int64_t maxs = g.K / 1024;
                if (nsplit > maxs) nsplit = (int)maxs;
                if (nsplit > 64) nsplit = 64;
                if (nsplit > 1) {
                    int64_t kc = (g.K + nsplit - 1) / nsplit;
                    kc = ((kc + GM_BK - 1) / GM_BK) * GM_BK;
                    nsplit = (int)((g.K + kc - 1) / kc);
                    g.kchunk = kc;
                    g.part = bpk_scratch((size_t)nsplit * g.M * g.N * sizeof(double));
                    if (!g.part) return bpk_set_error(BPK_ECUDA, "bpk_sum_multiply: scratch allocation failed");
                    grid.z = nsplit;
                }
            }
            if (grid.y <= 65535u) {
                static bool carve = false;
                if (!carve) {     // several CTAs per SM: ask for the shared-memory carve-out up front
                    cudaFuncSetAttribute(dgemm_dmma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
                    carve = true;
                }
                dgemm_dmma_kernel<<<grid, 128, 0, g_bpk.stream>>>(g);
                g_bpk.launches++;
                cudaError_t e_ = cudaPeekAtLastError();
                if (e_ != cudaSuccess) return bpk_set_error(BPK_ECUDA, "launch of dgemm_dmma_kernel failed: %s", cudaGetErrorString(e_));
                if (nsplit > 1) {
                    int64_t fb = (g.M * g.N + 255) / 256;
                    if (fb > (int64_t)g_bpk.sm_count * 8) fb = (int64_t)g_bpk.sm_count * 8;
                    BPK_LAUNCH(dgemm_splitk_final_kernel, (unsigned)fb, 256, 0, g, nsplit);
                }
                return BPK_OK;
            }
        }
    }
    if (A.ns == 0) { A.ns = 1; A.sshape[0] = 1; }   // pure broadcast product

    bool kept_contig = false;   // some operand is unit-stride along the innermost kept axis
    if (A.nk > 0)
        for (int k = 0; k < n_in; ++k) kept_contig = kept_contig || A.kin[k][A.nk - 1] == 1;
    bool use_thread = A.n_sum <= 8 || (A.n_kept >= 32768 && (kept_contig || A.n_sum <= 64));
    if (use_thread) {
        int64_t blocks = (A.n_kept + 255) / 256;
        int64_t cap = (int64_t)g_bpk.sm_count * 32;
        if (blocks > cap) blocks = cap;
        BPK_LAUNCH(sm_thread_kernel, (unsigned)blocks, 256, 0, A);
        return BPK_OK;
    }
    // block-per-output path
    int threads = A.n_sum >= 4096 ? 256 : (A.n_sum >= 512 ? 128 : 32);
    int64_t target_blocks = (int64_t)g_bpk.sm_count * 16;
    int nsplit = 1;
    if (A.n_kept < target_blocks) {
        int64_t want = target_blocks / A.n_kept;
        int64_t maxsplit = A.n_sum / (threads * 8);   // keep >= 8 terms per thread
        if (maxsplit < 1) maxsplit = 1;
        nsplit = (int)(want < maxsplit ? want : maxsplit);
        if (nsplit > 4096) nsplit = 4096;
        if (nsplit < 1) nsplit = 1;
    }
    A.nsplit = nsplit;
    A.chunk = (A.n_sum + nsplit - 1) / nsplit;
    if (nsplit > 1) {
        A.partial = bpk_scratch((size_t)A.n_kept * nsplit * sizeof(double));
        if (!A.partial) return bpk_set_error(BPK_ECUDA, "bpk_sum_multiply: scratch allocation failed");
    }
    int64_t blocks = A.n_kept * nsplit;
    int64_t cap = (int64_t)g_bpk.sm_count * 64;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(sm_block_kernel, (unsigned)blocks, threads, 0, A);
    if (nsplit > 1) {
        int64_t fb = (A.n_kept + 255) / 256;
        BPK_LAUNCH(sm_final_kernel, (unsigned)fb, 256, 0, A);
    }
    return BPK_OK;
}
