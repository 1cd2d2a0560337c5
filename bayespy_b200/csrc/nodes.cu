// nodes.cu — fused moment kernels of the scalar / simplex exponential families:
// Gamma (gamma.py:124-148), Dirichlet (dirichlet.py:130-160),
// Categorical/Multinomial softmax (multinomial.py:101-121, misc.py:1366-1401),
// one-hot encoding (categorical.py:30-47).
#include "common.cuh"

__global__ void gamma_kernel(const double *__restrict__ phi0, int64_t n0, const double *__restrict__ phi1, int64_t n1,
                             int64_t n, double *__restrict__ u0, double *__restrict__ u1, double *__restrict__ g, int *flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double b = -phi0[n0 == 1 ? 0 : i];
        double a = phi1[n1 == 1 ? 0 : i];
        if (!(b > 0.0) || !(a > 0.0)) atomicOr(flag, BPK_FLAG_DOMAIN);   // np.errstate(raise) gamma.py:142
        double logb = log(b);
        if (u0) u0[i] = a / b;
        if (u1) u1[i] = bpk_digamma(a) - logb;
        if (g) g[i] = a * logb - lgamma(a);
    }
}

extern "C" int bpk_gamma_moments(const double *phi0, int64_t n0, const double *phi1, int64_t n1,
                                 int64_t n, double *u0, double *u1, double *g, int check) {
    BPK_REQUIRE_INIT();
    if ((n0 != 1 && n0 != n) || (n1 != 1 && n1 != n))
        return bpk_set_error(BPK_EINVAL, "bpk_gamma_moments: n0,n1 must be 1 or n");
    if (n <= 0) return BPK_OK;
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)g_bpk.sm_count * 16;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(gamma_kernel, (unsigned)blocks, 256, 0, phi0, n0, phi1, n1, n, u0, u1, g, g_bpk.d_flag);
    if (check) return bpk_check_flag(BPK_EDOMAIN);
    return BPK_OK;
}

// one warp per row of K concentration parameters
__global__ void dirichlet_kernel(const double *__restrict__ phi, int64_t n, int K,
                                 double *__restrict__ u, double *__restrict__ g, int *flag) {
    const int lane = threadIdx.x & 31;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t step = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; r < n; r += step) {
        const double *p = phi + r * K;
        double s = 0.0, lg = 0.0;
        int bad = 0;
        for (int k = lane; k < K; k += 32) {
            double a = p[k];
            if (!(a > 0.0)) bad = 1;      // "Natural parameters should be positive" dirichlet.py:147
            s += a;
            lg += lgamma(a);
        }
        s = warp_sum(s);
        lg = warp_sum(lg);
        if (bad) atomicOr(flag, BPK_FLAG_DOMAIN);
        double ps = bpk_digamma(s);
        if (u)
            for (int k = lane; k < K; k += 32) u[r * K + k] = bpk_digamma(p[k]) - ps;
        if (g && lane == 0) g[r] = lgamma(s) - lg;
    }
}

extern "C" int bpk_dirichlet_moments(const double *phi, int64_t n, int K, double *u, double *g, int check) {
    BPK_REQUIRE_INIT();
    if (K < 1) return bpk_set_error(BPK_EINVAL, "bpk_dirichlet_moments: K=%d", K);
    if (n <= 0) return BPK_OK;
    int64_t blocks = (n + 7) / 8;
    int64_t cap = (int64_t)g_bpk.sm_count * 16;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(dirichlet_kernel, (unsigned)blocks, 256, 0, phi, n, K, u, g, g_bpk.d_flag);
    if (check) return bpk_check_flag(BPK_EDOMAIN);
    return BPK_OK;
}

// softmax with the reference's second renormalisation (misc.py:1398-1401):
//   m = max phi; e = exp(phi-m); lse = log(sum e) + m; p = exp(phi - lse); p /= sum p
// one warp per row.
__global__ void softmax_kernel(const double *__restrict__ phi, int64_t n, int K,
                               double *__restrict__ u, double *__restrict__ g) {
    const int lane = threadIdx.x & 31;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t step = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; r < n; r += step) {
        const double *p = phi + r * K;
        double m = -INFINITY;
        for (int k = lane; k < K; k += 32) m = fmax(m, p[k]);
        m = warp_max(m);
        double mm = isfinite(m) ? m : 0.0;
        double s = 0.0;
        for (int k = lane; k < K; k += 32) s += exp(p[k] - mm);
        s = warp_sum(s);
        double lse = log(s) + mm;
        double s2 = 0.0;
        for (int k = lane; k < K; k += 32) s2 += exp(p[k] - lse);
        s2 = warp_sum(s2);
        if (u)
            for (int k = lane; k < K; k += 32) u[r * K + k] = exp(p[k] - lse) / s2;
        if (g && lane == 0) g[r] = -lse;
    }
}

extern "C" int bpk_softmax_moments(const double *phi, int64_t n, int K, double *u, double *g) {
    BPK_REQUIRE_INIT();
    if (K < 1) return bpk_set_error(BPK_EINVAL, "bpk_softmax_moments: K=%d", K);
    if (n <= 0) return BPK_OK;
    int64_t blocks = (n + 7) / 8;
    int64_t cap = (int64_t)g_bpk.sm_count * 32;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(softmax_kernel, (unsigned)blocks, 256, 0, phi, n, K, u, g);
    return BPK_OK;
}

__global__ void one_hot_kernel(const int64_t *__restrict__ labels, int64_t n, int K, double *__restrict__ u, int *flag) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; e < n * K; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e / K;
        int k = (int)(e - r * K);
        int64_t l = labels[r];
        if (k == 0 && (l < 0 || l >= K)) atomicOr(flag, BPK_FLAG_DOMAIN);
        u[e] = (l == k) ? 1.0 : 0.0;
    }
}

extern "C" int bpk_one_hot(const int64_t *labels, int64_t n, int K, double *u, int check) {
    BPK_REQUIRE_INIT();
    if (K < 1) return bpk_set_error(BPK_EINVAL, "bpk_one_hot: K=%d", K);
    if (n <= 0) return BPK_OK;
    int64_t total = n * K;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)g_bpk.sm_count * 32;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(one_hot_kernel, (unsigned)blocks, 256, 0, labels, n, K, u, g_bpk.d_flag);
    if (check) {
        int rc = bpk_check_flag(BPK_EDOMAIN);
        if (rc == BPK_EDOMAIN) return bpk_set_error(BPK_EINVAL, "Invalid category index");
        return rc;
    }
    return BPK_OK;
}
