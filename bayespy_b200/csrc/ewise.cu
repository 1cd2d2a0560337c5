// ewise.cu — generic broadcast elementwise kernel (fp64) of libbpk.
// HBM-bound streaming: 1 thread per output element, grid-stride, index
// decomposition only when the caller could not collapse the index space to 1-D.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

struct EwArgs {
    int op, nd, n_in;
    int64_t total;
    int64_t shape[BPK_MAXD];
    int64_t ostride[BPK_MAXD];
    int64_t istride[3][BPK_MAXD];
    const void *in[3];
    int dtype[3];
    double *out;
    double alpha, beta;
};

__device__ __forceinline__ double ew_load(const void *p, int dtype, int64_t off) {
    return dtype == BPK_U8 ? (double)((const uint8_t *)p)[off] : ((const double *)p)[off];
}

__device__ __forceinline__ double ew_apply(int op, double a, double b, double c, double alpha, double beta) {
    switch (op) {
    case BPK_OP_COPY: return a;
    case BPK_OP_ADD: return a + b;
    case BPK_OP_SUB: return a - b;
    case BPK_OP_MUL: return a * b;
    case BPK_OP_DIV: return a / b;
    case BPK_OP_AXPBY: return alpha * a + beta * b;
    case BPK_OP_AFFINE: return alpha * a + beta;
    case BPK_OP_FMA: return alpha * a * b + beta * c;
    case BPK_OP_WHERE: return a != 0.0 ? b : c;
    case BPK_OP_LOG: return log(a);
    case BPK_OP_EXP: return exp(a);
    case BPK_OP_RECIP: return alpha / a;
    case BPK_OP_SQUARE: return a * a;
    case BPK_OP_SQRT: return sqrt(a);
    case BPK_OP_LGAMMA: return lgamma(a);
    case BPK_OP_DIGAMMA: return bpk_digamma(a);
    case BPK_OP_MVLGAMMA: return bpk_mvlgamma(a, (int)alpha);
    case BPK_OP_MVDIGAMMA: return bpk_mvdigamma(a, (int)alpha);
    case BPK_OP_NONZERO: return a != 0.0 ? b : 0.0;
    case BPK_OP_TRIGAMMA: return bpk_trigamma(a);
    }
    return nan("");
}

__global__ void __launch_bounds__(256) ewise_kernel(EwArgs A) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; i < A.total; i += step) {
        int64_t rem = i, oo = 0, o0 = 0, o1 = 0, o2 = 0;
#pragma unroll 1
        for (int d = A.nd - 1; d >= 0; --d) {
            int64_t q = rem / A.shape[d];
            int64_t r = rem - q * A.shape[d];
            rem = q;
            oo += r * A.ostride[d];
            o0 += r * A.istride[0][d];
            o1 += r * A.istride[1][d];
            o2 += r * A.istride[2][d];
        }
        double a = A.n_in > 0 ? ew_load(A.in[0], A.dtype[0], o0) : 0.0;
        double b = A.n_in > 1 ? ew_load(A.in[1], A.dtype[1], o1) : 0.0;
        double c = A.n_in > 2 ? ew_load(A.in[2], A.dtype[2], o2) : 0.0;
        A.out[oo] = ew_apply(A.op, a, b, c, A.alpha, A.beta);
    }
}

// Fast path (the plate-sized arrays of the per-node paths land here): at most one outer axis, the inner axis
// contiguous in the output and, in every input, contiguous (stride 1) or broadcast (stride 0).  Two elements per
// thread per step as 16-byte accesses, one 32-bit division per step (none for a single row).
template <bool WIDE>
__global__ void __launch_bounds__(256) ewise_rows_kernel(EwArgs A, int64_t half_cols) {
    const int64_t total2 = A.total >> 1;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const bool one_row = A.nd == 1;
    const double *p0 = (const double *)A.in[0], *p1 = (const double *)A.in[1], *p2 = (const double *)A.in[2];
    const int64_t si0 = A.istride[0][A.nd - 1], si1 = A.istride[1][A.nd - 1], si2 = A.istride[2][A.nd - 1];
    const int64_t so0 = one_row ? 0 : A.istride[0][0], so1 = one_row ? 0 : A.istride[1][0], so2 = one_row ? 0 : A.istride[2][0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += step) {
        int64_t row = 0, c2 = i;
        if (!one_row) {
            if (WIDE) { row = i / half_cols; c2 = i - row * half_cols; }
            else { const unsigned r = (unsigned)i / (unsigned)half_cols; row = r; c2 = (unsigned)i - r * (unsigned)half_cols; }
        }
        const int64_t col = 2 * c2;
        double2 a = make_double2(0.0, 0.0), b = a, c = a;
        if (si0) a = *(const double2 *)(p0 + row * so0 + col); else { const double v = p0[row * so0]; a = make_double2(v, v); }
        if (A.n_in > 1) {
            if (si1) b = *(const double2 *)(p1 + row * so1 + col); else { const double v = p1[row * so1]; b = make_double2(v, v); }
        }
        if (A.n_in > 2) {
            if (si2) c = *(const double2 *)(p2 + row * so2 + col); else { const double v = p2[row * so2]; c = make_double2(v, v); }
        }
        double2 r;
        r.x = ew_apply(A.op, a.x, b.x, c.x, A.alpha, A.beta);
        r.y = ew_apply(A.op, a.y, b.y, c.y, A.alpha, A.beta);
        *(double2 *)(A.out + 2 * i) = r;
    }
}

static bool ew_rows_ok(const EwArgs &A) {
    if (A.nd < 1 || A.nd > 2 || getenv("BPK_EWISE_GENERIC")) return false;
    const int in = A.nd - 1;
    const int64_t C = A.shape[in];
    if ((C & 1) || A.ostride[in] != 1 || ((uintptr_t)A.out & 15)) return false;
    if (A.nd == 2 && A.ostride[0] != C) return false;
    for (int k = 0; k < A.n_in; ++k) {
        if (A.dtype[k] != BPK_F64) return false;
        const int64_t si = A.istride[k][in], so = A.nd == 2 ? A.istride[k][0] : 0;
        if (si != 0 && si != 1) return false;
        if (so < 0) return false;
        if (si == 1 && (((uintptr_t)A.in[k] & 15) || (so & 1))) return false;
    }
    return true;
}

extern "C" int bpk_ewise(int op, int nd, const int64_t *shape,
                         double *out, const int64_t *out_stride,
                         int n_in, const void *const *in, const int *in_dtype,
                         const int64_t *in_stride, double alpha, double beta) {
    BPK_REQUIRE_INIT();
    if (op < 0 || op >= BPK_OP_COUNT_) return bpk_set_error(BPK_EINVAL, "bpk_ewise: bad op %d", op);
    if (nd < 0 || nd > BPK_MAXD) return bpk_set_error(BPK_EINVAL, "bpk_ewise: nd=%d exceeds %d", nd, BPK_MAXD);
    if (n_in < 1 || n_in > 3) return bpk_set_error(BPK_EINVAL, "bpk_ewise: n_in=%d", n_in);
    EwArgs A;
    memset(&A, 0, sizeof(A));
    A.op = op; A.nd = nd; A.n_in = n_in; A.out = out; A.alpha = alpha; A.beta = beta;
    A.total = 1;
    for (int d = 0; d < nd; ++d) {
        if (shape[d] < 0) return bpk_set_error(BPK_EINVAL, "bpk_ewise: negative extent");
        A.shape[d] = shape[d];
        A.ostride[d] = out_stride[d];
        A.total *= shape[d];
        for (int k = 0; k < n_in; ++k) A.istride[k][d] = in_stride[k * nd + d];
    }
    for (int k = 0; k < n_in; ++k) { A.in[k] = in[k]; A.dtype[k] = in_dtype[k]; }
    if (A.total == 0) return BPK_OK;
    int64_t cap = (int64_t)g_bpk.sm_count * 32;
    if (ew_rows_ok(A)) {
        const int64_t total2 = A.total >> 1;
        int64_t blocks = (total2 + 255) / 256;
        if (blocks > cap) blocks = cap;
        const int64_t half_cols = A.shape[nd - 1] >> 1;
        if (total2 < (1ll << 32) && half_cols < (1ll << 32)) BPK_LAUNCH(ewise_rows_kernel<false>, (unsigned)blocks, 256, 0, A, half_cols);
        else BPK_LAUNCH(ewise_rows_kernel<true>, (unsigned)blocks, 256, 0, A, half_cols);
        return BPK_OK;
    }
    int64_t blocks = (A.total + 255) / 256;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(ewise_kernel, (unsigned)blocks, 256, 0, A);
    return BPK_OK;
}

// ---- index gathers along one plate axis (nodes/take.py) ---------------------------------------------------------
// out[a][j][c] = in[a][idx[j]][c]; pure data movement: bit-exact
__global__ void __launch_bounds__(256) take_kernel(const double *__restrict__ in, int64_t pre, int64_t L, int64_t post,
                                                   const int64_t *__restrict__ idx, int64_t J, double *__restrict__ out) {
    const int64_t total = pre * J * post;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = e % post, j = (e / post) % J, a = e / (post * J);
        out[e] = in[(a * L + idx[j]) * post + c];
    }
}
// out[a][i][c] = sum over the j with idx[j] == i of in[a][j][c], in increasing j (order/start: stable sort of idx, CSR)
__global__ void __launch_bounds__(256) put_add_kernel(const double *__restrict__ in, int64_t pre, int64_t J, int64_t post,
                                                      const int64_t *__restrict__ order, const int64_t *__restrict__ start,
                                                      int64_t L, double *__restrict__ out) {
    const int64_t total = pre * L * post;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = e % post, i = (e / post) % L, a = e / (post * L);
        double s = 0.0;
        for (int64_t q = start[i]; q < start[i + 1]; ++q) s += in[(a * J + order[q]) * post + c];
        out[e] = s;
    }
}
static unsigned take_grid(int64_t total) {
    int64_t blocks = (total + 255) / 256, cap = (int64_t)g_bpk.sm_count * 32;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks < 1 ? 1 : blocks);
}
extern "C" int bpk_take(const double *in, int64_t pre, int64_t L, int64_t post, const int64_t *idx, int64_t J, double *out) {
    BPK_REQUIRE_INIT();
    if (pre < 0 || L < 1 || post < 0 || J < 0) return bpk_set_error(BPK_EINVAL, "bpk_take: bad shape");
    if (pre * J * post == 0) return BPK_OK;
    BPK_LAUNCH(take_kernel, take_grid(pre * J * post), 256, 0, in, pre, L, post, idx, J, out);
    return BPK_OK;
}
extern "C" int bpk_put_add(const double *in, int64_t pre, int64_t J, int64_t post, const int64_t *order,
                           const int64_t *start, int64_t L, double *out) {
    BPK_REQUIRE_INIT();
    if (pre < 0 || L < 1 || post < 0 || J < 0) return bpk_set_error(BPK_EINVAL, "bpk_put_add: bad shape");
    if (pre * L * post == 0) return BPK_OK;
    BPK_LAUNCH(put_add_kernel, take_grid(pre * L * post), 256, 0, in, pre, J, post, order, start, L, out);
    return BPK_OK;
}
