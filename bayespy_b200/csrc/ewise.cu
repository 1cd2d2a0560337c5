// ewise.cu — generic broadcast elementwise kernel (fp64) of libbpk.
// HBM-bound streaming: 1 thread per output element, grid-stride, index
// decomposition only when the caller could not collapse the index space to 1-D.
#include "common.cuh"
#include <string.h>

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
    int64_t blocks = (A.total + 255) / 256;
    int64_t cap = (int64_t)g_bpk.sm_count * 32;
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
