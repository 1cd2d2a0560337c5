/*
 * bpk.h — C ABI of libbpk.so, the B200 (sm_100a) kernel library behind the
 * BayesPy plate-batched VMP update path.
 *
 * Boundary.  The reference (pure Python) has no FFI; the seams
 * this ABI replaces are, citing /root/reference paths:
 *   seam 1  bayespy/utils/misc.py:851 sum_multiply, :935 sum_product,
 *           :805 sum_multiply_to_plates            -> bpk_sum_multiply
 *           bayespy/utils/linalg.py:31 chol, :66 chol_solve, :174 chol_inv,
 *           :209 chol_logdet                       -> bpk_chol*
 *   seam 2  the Distribution protocol methods
 *           (bayespy/inference/vmp/nodes/expfamily.py:26-43) of
 *           gaussian.py:672 (GaussianARD), gaussian.py:397 (Gaussian),
 *           gamma.py:124, wishart.py:165, dirichlet.py:130,
 *           multinomial.py:101, mixture.py:53    -> bpk_*_moments, bpk_mix_*
 *           and dot.py:355,403,581 (SumMultiply)   -> bpk_sum_multiply / bpk_pca_*
 *   seam 3  VB.update (bayespy/inference/vmp/vmp.py:132) -> fused sweeps
 *           bpk_pca_* / bpk_gmm_* scheduled by bayespy_b200.inference.VB
 *
 * Conventions: plain C, no torch types.  All entry points return 0 on success
 * or a BPK_E* code (text via bpk_last_error()).  All arrays are fp64 unless a
 * dtype is given; "dev" pointers are device pointers returned by bpk_malloc.
 * Everything is ordered on the library's single compute stream; only bpk_d2h,
 * bpk_sync and the *_check variants block the host.  One process drives one
 * GPU (one rank per GPU under torchrun); the caller owns every buffer.
 */
#ifndef BPK_H
#define BPK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPK_MAXD 8          /* max broadcast rank of the generic kernels     */
#define BPK_MAXIN 4         /* max inputs of bpk_ewise / bpk_sum_multiply    */
#define BPK_MAXDIM 64       /* max matrix dimension of the batched linalg    */

/* status codes */
#define BPK_OK 0
#define BPK_ECUDA 1         /* CUDA runtime error                            */
#define BPK_EINVAL 2        /* bad argument (-> ValueError)                  */
#define BPK_ENOTSPD 3       /* "Matrix not positive definite" linalg.py:58   */
#define BPK_EDOMAIN 4       /* non-positive natural parameter (gamma.py:142, dirichlet.py:147) */
#define BPK_ENCCL 5
#define BPK_ENOGPU 6

/* dtypes of generic-kernel inputs */
#define BPK_F64 0
#define BPK_U8 1            /* bool masks                                    */

/* ---- lifecycle / memory ------------------------------------------------ */
int bpk_init(int device);
int bpk_shutdown(void);
const char *bpk_last_error(void);
int bpk_device_info(int *sm_count, int *cc_major, int *cc_minor,
                    uint64_t *hbm_total, uint64_t *hbm_free);
/* PCI bus id ("0000:1b:00.0") of the bound GPU: lets the host pin a rank's threads and pinned buffers to the
 * GPU's NUMA node (parallel.bind_to_gpu_numa).  The reference is single-process NumPy and has no such notion.   */
int bpk_device_pci_bus_id(char *buf, int len);
int bpk_sync(void);
uint64_t bpk_launch_count(void);      /* kernels launched since bpk_init    */
int bpk_malloc(void **dev, uint64_t bytes);
int bpk_free(void *dev);
int bpk_h2d(void *dev, const void *host, uint64_t bytes);
int bpk_d2h(void *host, const void *dev, uint64_t bytes);   /* blocks */
int bpk_d2d(void *dst, const void *src, uint64_t bytes);
int bpk_memset(void *dev, int byte, uint64_t bytes);
int bpk_host_alloc(void **host, uint64_t bytes);            /* pinned */
int bpk_host_free(void *host);
/* CUDA-event timers on the compute stream (bench.py) */
int bpk_timer_create(int *id);
int bpk_timer_record(int id, int which /*0=start,1=stop*/);
int bpk_timer_elapsed_ms(int id, double *ms);               /* blocks */
int bpk_flush_l2(void);               /* overwrite a >L2-sized scratch buffer */

/* ---- multi-GPU: one NCCL communicator per process -------- */
int bpk_comm_unique_id(char id[128]);
int bpk_comm_init(const char id[128], int nranks, int rank);
int bpk_comm_size(int *nranks, int *rank);
int bpk_allreduce_sum_f64(double *dev, uint64_t count);     /* in place, on the compute stream */
/* out of place: dst = sum over ranks of src (dst may equal src)                */
int bpk_allreduce_sum_f64_oop(const double *src, double *dst, uint64_t count);
int bpk_comm_destroy(void);
/* Peer-memory exchange window for the in-kernel all-reduce of the resident loop: each rank
 * creates one window, the 64-byte CUDA IPC handles are exchanged by the host (any side
 * channel), and every rank opens all of them.  Up to 8 ranks on one NVLink/NVSwitch box.     */
int bpk_xchg_create(char handle[64]);
int bpk_xchg_open(const char *handles /* [nranks][64] */, int nranks, int rank);
int bpk_xchg_close(void);

/* ---- generic broadcast kernels (seam 1) -------------------------------- */
/* Elementwise out[i] = op(in0[i], in1[i], in2[i]; alpha, beta) over an
 * nd-dimensional index space `shape`.  Strides are in elements; 0 = broadcast.
 * The output may be strided too (e.g. the diagonal of a KxK tile).          */
enum {
    BPK_OP_COPY = 0,    /* a                                   */
    BPK_OP_ADD,         /* a + b                               */
    BPK_OP_SUB,         /* a - b                               */
    BPK_OP_MUL,         /* a * b                               */
    BPK_OP_DIV,         /* a / b                               */
    BPK_OP_AXPBY,       /* alpha*a + beta*b                    */
    BPK_OP_AFFINE,      /* alpha*a + beta                      */
    BPK_OP_FMA,         /* alpha*a*b + beta*c                  */
    BPK_OP_WHERE,       /* a(mask) ? b : c                     */
    BPK_OP_LOG,         /* log(a)                              */
    BPK_OP_EXP,         /* exp(a)                              */
    BPK_OP_RECIP,       /* alpha / a                           */
    BPK_OP_SQUARE,      /* a*a                                 */
    BPK_OP_SQRT,        /* sqrt(a)                             */
    BPK_OP_LGAMMA,      /* lgamma(a)      scipy.special.gammaln */
    BPK_OP_DIGAMMA,     /* psi(a)         scipy.special.psi     */
    BPK_OP_MVLGAMMA,    /* multigammaln(a, d=(int)alpha)        */
    BPK_OP_MVDIGAMMA,   /* sum_{i<d} psi(a - i/2), d=(int)alpha  misc.py:1146 */
    BPK_OP_NONZERO,     /* a != 0 ? b : 0   (expfamily.py:463 "0 * -inf") */
    BPK_OP_TRIGAMMA,    /* psi'(a)        scipy.special.polygamma(1, a), gamma.py:210 */
    BPK_OP_COUNT_
};
int bpk_ewise(int op, int nd, const int64_t *shape,
              double *out, const int64_t *out_stride,
              int n_in, const void *const *in, const int *in_dtype,
              const int64_t *in_stride /* [n_in][nd] */,
              double alpha, double beta);

/* out[kept] = scale * sum_{summed} prod_i in_i[...]  (+ out if accumulate).
 * Axes with out_stride==0 are summed.  This is misc.sum_multiply (misc.py:851,
 * an np.einsum over broadcast operands) and the masked plate-sum of
 * Node._message_to_parent (node.py:650).                                    */
int bpk_sum_multiply(int nd, const int64_t *shape,
                     int n_in, const void *const *in, const int *in_dtype,
                     const int64_t *in_stride /* [n_in][nd] */,
                     double *out, const int64_t *out_stride,
                     double scale, int accumulate);

/* ---- batched dense SPD linear algebra (linalg.py:31-223) --------------- */
/* A,U: [batch][D][D] row-major.  U is upper-triangular with A = U^T U
 * (scipy cho_factor(lower=False)); the strict lower triangle is written as 0
 * (the reference leaves it undefined).  check!=0: block and return
 * BPK_ENOTSPD if any matrix was not positive definite.                      */
int bpk_chol(const double *A, double *U, int64_t batch, int D, int check);
/* X = A^{-1} B for B: [batchB][D][nrhs]; batchU, batchB in {1, batch}.      */
int bpk_chol_solve(const double *U, int64_t batchU, const double *B, int64_t batchB,
                   double *X, int64_t batch, int D, int nrhs);
int bpk_chol_inv(const double *U, double *Ainv, int64_t batch, int D);
int bpk_chol_logdet(const double *U, double *out, int64_t batch, int D);

/* Block-tridiagonal SPD system of a Gaussian Markov chain (linalg.py:468-575 block_banded_solve,
 * called from gaussian_markov_chain.py:89-123): A [batch][T][D][D] diagonal blocks, B [batch][T-1][D][D]
 * super-diagonal blocks (sub-diagonal = transposes), y [batch][T][D].  Returns the diagonal blocks V
 * [batch][T][D][D] and super-diagonal blocks C [batch][T-1][D][D] of the INVERSE, the solution x
 * [batch][T][D] and log det [batch] — i.e. the Kalman/RTS smoother in information form.            */
int bpk_block_banded_solve(const double *A, const double *B, const double *y,
                           int64_t batch, int64_t T, int D,
                           double *V, double *C, double *x, double *logdet, int check);

/* ---- fused per-node moment kernels (seam 2) ---------------------------- */
/* Gaussian / GaussianARD (gaussian.py:397-446, :672-706):
 *   Lambda = -2 phi1,  Cov = Lambda^-1,  u0 = Cov phi0,
 *   g = -1/2 u0.phi0 + 1/2 log|Lambda|.   phi0: [n0][K], phi1: [n1][K][K],
 *   n0,n1 in {1,N}.  Outputs u0 [N][K], cov [n1][K][K] (second moment is
 *   cov + u0 u0^T, materialised on demand by bpk_outer_add), g [N].
 *   Outputs may be NULL to skip.                                            */
int bpk_gaussian_moments(const double *phi0, int64_t n0, const double *phi1, int64_t n1,
                         int64_t N, int K, double *u0, double *cov, double *g,
                         double *logdet /* [n1] or NULL */, int check);
/* u1[n] = cov[n or 0] + u0[n] u0[n]^T                                        */
int bpk_outer_add(const double *u0, const double *cov, int64_t ncov,
                  int64_t N, int K, double *u1);
/* Gamma (gamma.py:124-148): a = phi1, b = -phi0: u0=a/b, u1=psi(a)-log b,
 * g = a log b - lgamma(a).  n0,n1 in {1,n}.                                 */
int bpk_gamma_moments(const double *phi0, int64_t n0, const double *phi1, int64_t n1,
                      int64_t n, double *u0, double *u1, double *g, int check);
/* Wishart (wishart.py:165-188): V=-phi0 [n][D][D], nu/2 = phi1 [n1]:
 * u0 = phi1 V^-1, u1 = -log|V| + psi_D(phi1), g = phi1 log|V| - lgamma_D(phi1) */
int bpk_wishart_moments(const double *phi0, const double *phi1, int64_t n1,
                        int64_t n, int D, double *u0, double *u1, double *g, int check);
/* Dirichlet (dirichlet.py:130-160): u = psi(phi) - psi(sum phi),
 * g = lgamma(sum phi) - sum lgamma(phi); phi: [n][K]                         */
int bpk_dirichlet_moments(const double *phi, int64_t n, int K,
                          double *u, double *g, int check);
/* Categorical/Multinomial(1) (multinomial.py:101-121, misc.py:1366-1401):
 * u = softmax(phi) (renormalised), g = -logsumexp(phi); phi: [n][K]          */
int bpk_softmax_moments(const double *phi, int64_t n, int K, double *u, double *g);
/* one-hot encode integer labels (categorical.py:30-47); bit-exact           */
int bpk_one_hot(const int64_t *labels, int64_t n, int K, double *u, int check);

/* index gathers along one plate axis (nodes/take.py:63-88, misc.py:549-585 put_simple); idx on the device, already
 * normalised to [0, L).  bpk_take: out[a][j][c] = in[a][idx[j]][c] (bit-exact data movement).  bpk_put_add: the
 * inverse with accumulation, out[a][i][c] = sum_{j: idx[j]==i} in[a][j][c] in increasing j; order = stable argsort
 * of idx, start = its CSR offsets (L+1 entries), both built by the host once per node.                              */
int bpk_take(const double *in, int64_t pre, int64_t L, int64_t post, const int64_t *idx, int64_t J, double *out);
int bpk_put_add(const double *in, int64_t pre, int64_t J, int64_t post, const int64_t *order,
                const int64_t *start, int64_t L, double *out);

/* ---- fused sweep kernels (seam 3) -------------------------- */
/* PCA / factor model  y[m,n] ~ N(w_m . x_n, 1/tau), fully observed.
 * One pass over Y: x_n = A y_n + b for every column n (A: [K][M], b: [K]),
 * writes X [N][K] and accumulates the sufficient statistics
 *   stats = [ S_yx (M*K) | S_xx (K*K) | s_x (K) ]  (sum_n y_n x_n^T, x_n x_n^T, x_n)
 * Y: [M][N] row-major (n fastest).  stats must be zeroed by the caller
 * (or hold a partial to accumulate onto).  640 B/col at M=64, K=16.         */
int bpk_pca_xsweep(const double *Y, int64_t M, int64_t N, int K,
                   const double *A, const double *b, double *X, double *stats);
/* statistics only, for an X that did not come from bpk_pca_xsweep           */
int bpk_pca_stats(const double *Y, int64_t M, int64_t N, int K,
                  const double *X, double *stats);
/* masked variant: per-column precision.  mask: [M][N] u8.
 * Lam_n = diag(alpha) + tau sum_m mask[m,n] WW[m];  phi0_n = tau sum_m mask y w_m + amu
 * writes X [N][K], optional COV [N][K][K], g [N], and stats
 *   [ S_yx (M*K) | S_xx (M*K*K, per row m incl. covariance) ]               */
int bpk_pca_xsweep_masked(const double *Y, const uint8_t *mask, int64_t M, int64_t N, int K,
                          const double *W /*[M][K]*/, const double *WW /*[M][K][K]*/,
                          double tau, const double *alpha /*[K]*/, const double *amu /*[K]*/,
                          double *X, double *COV /*nullable*/, double *g /*[N]*/,
                          double *stats, int check);
/* The same sweep FUSED (csrc/pca_masked.cu; M <= 64, K <= 16): per-column precisions are built, inverted and
 * contracted chunk by chunk through a fixed-size scratch; nothing of size (N,K,K) is stored.  Replaces, per sweep,
 * the per-plate SciPy loops of linalg.py:50-59,111-146,185-195 driven by gaussian.py:672-706 and the two masked
 * plate sums of dot.py:581 / node.py:650.  Writes X [N][K], g [N] (nullable: cgf of q(x_n)), and ADDS to
 *   stats = [ S_yx (M*K) | S_xx (M*K*K: sum_n mask[m,n] <x_n x_n^T>) | sum_n <x_n x_n^T> (K*K) | sum_n x_n (K)
 *             | sum_n phi_n.x_n | sum_n log det Lam_n ]          (caller zeroes it).  704 B/col of HBM traffic.   */
int bpk_pca_xsweep_masked_fused(const double *Y, const uint8_t *mask, int64_t M, int64_t N, int K,
                                const double *W /*[M][K]*/, const double *WW /*[M][K][K]*/,
                                double tau, const double *alpha /*[K]*/, const double *amu /*[K] or NULL*/,
                                double *X, double *g /*nullable*/, double *stats, int check);
/* sum_{m,n} mask*y^2 and count (constants of the tau update)                */
int bpk_sumsq(const double *Y, const uint8_t *mask /*nullable*/, int64_t count,
              double *out2 /* [2]: sum y^2, #observed */);

/* Gaussian mixture E-step + statistics in one pass over y [N][D]:
 *   L[n,k] = c[k] + y_n.h[k] - 1/2 y_n^T Lam[k] y_n + logpi[k]
 *   p[n,:] = softmax(L[n,:]);  writes P [N][K] (nullable), g [N] = -logsumexp
 *   stats = [ sum_n p (K) | sum_n p y (K*D) | sum_n p y y^T (K*D*D) | sum_n logsumexp (1) ] */
int bpk_gmm_sweep(const double *Y, int64_t N, int D, int K,
                  const double *c, const double *h, const double *Lam, const double *logpi,
                  double *P, double *g, double *stats);

/* statistics only, for responsibilities P [N][K] that did not come from bpk_gmm_sweep
 * (e.g. a random initialisation): stats[0..K+K*D+K*D*D) += sums, stats[last] unchanged.   */
int bpk_gmm_stats(const double *Y, int64_t N, int D, int K, const double *P, double *stats);

/* ---- device-resident VB loop of the factor model (seam 3: VB.update, vmp.py:132-172) ----
 * Model: X=GaussianARD(mu_x,a_x,plates=(1,N),shape=(K,)), C=GaussianARD(mu_c,alpha,plates=(M,1),
 * shape=(K,)), alpha~Gamma(a0,b0) plates (K,) or constant, tau~Gamma(ta0,tb0) or constant,
 * Y=GaussianARD(SumMultiply(X,C),tau) fully observed (doc/source/examples/pca.rst:40-66).
 * `ops` is ONE iteration of the user's update order as opcodes; it is replayed `niter` times
 * with no host round trip: the big op is the one-pass sweep kernel, every run of small ops is
 * one single-CTA launch.  BOUND evaluates all lower-bound terms (expfamily.py:400-480), appends
 * [L_Y, L_X, L_C, L_alpha, L_tau, L] to Lhist, and applies the convergence test of vmp.py:738-747
 * on device (tol < 0 disables it); once it fires, ctrl[1] is raised and every later kernel of the
 * run is a no-op, so the state is exactly the one the reference would have stopped at.
 * state: fp64 vector laid out per bpk_pca_vb_layout (hyper-parameters, q(C), q(alpha), q(tau),
 * shared part of q(X), plate-summed statistics); X: [N][K] posterior means, rewritten every sweep.
 * ctrl: device int[4] = {iterations finished, stop, error bits (1 = not SPD, 2 = domain, 4 = peer exchange
 * timed out, 8 = grid barrier timed out), 0}.
 * With a communicator (bpk_comm_init) STATS is followed by the sweep's one all-reduce.          */
enum {
    BPK_VBOP_XSWEEP = 1,  /* X.update(): one pass over Y (pca_xsweep_kernel)                    */
    BPK_VBOP_STATS = 2,   /* grid reduction of the sweep's partial statistics (+ all-reduce)    */
    BPK_VBOP_SXXT = 3,    /* sum_n <x x^T> = N Cov_x + S_xx                                     */
    BPK_VBOP_XPRE = 4,    /* shared part of q(X): Cov_x, A = tau Cov_x <W>^T, b                 */
    BPK_VBOP_ROW = 5,     /* C.update()                                                         */
    BPK_VBOP_ALPHA = 6,   /* alpha.update()                                                     */
    BPK_VBOP_TAU = 7,     /* tau.update()                                                       */
    BPK_VBOP_BOUND = 8    /* lower bound + convergence test                                     */
};
int bpk_pca_vb_layout(int M, int K, int64_t *offsets /* [nfields+1] */, int *nfields);
const char *bpk_pca_vb_field_name(int i);
int bpk_pca_vb_run(const double *Y, int64_t M, int64_t N, int K, double *X, double *state,
                   const int *ops, int nops, int niter, int has_alpha, int has_tau, double tol,
                   double *Lhist, int cap, int *ctrl);
/* Multi-rank runs: every rank must drive the in-kernel exchange the same way.  The persistent multi-sweep launch
 * needs 16-byte aligned rows on THIS rank (even N); no_loop = 1 makes this rank use one launch per sweep with the
 * full statistics vector on the wire, which is what ranks with unaligned shards do anyway (the host agrees on the
 * flag collectively, plans.FactorModelPlan).  Returns the previous value through *prev (nullable).              */
int bpk_pca_vb_set_mode(int no_loop, int *prev);
/* bench.py: record these timers (bpk_timer_create ids) around the next n sweep-kernel launches */
int bpk_pca_vb_set_timers(const int *ids, int n);
int bpk_pca_vb_timers_used(void);
/* diagnostics: device-clock stamps of the last fused sweep launch (needs BPK_VB_DEBUG=1 in the environment) */
int bpk_debug_stamps(uint64_t *out, int n);

/* ---- device-resident VB loop of the Gaussian mixture model --------------------------------------
 * Replaces, for  Y = Mixture(Z, Gaussian, mu, Lambda)  with  Z ~ Categorical(alpha), alpha ~ Dirichlet,
 * mu ~ Gaussian, Lambda ~ Wishart (doc/source/examples/gmm.rst:71-98), the scheduler loop of VB.update
 * (bayespy/inference/vmp/vmp.py:132-172) together with the node updates and lower-bound terms it calls:
 * nodes/mixture.py:108-225, gaussian.py:341-463, wishart.py:136-205, dirichlet.py:120-170,
 * multinomial.py:83-130, expfamily.py:400-480, and the convergence test vmp.py:717-747.
 * `ops` is ONE iteration's program in the user's update order (one opcode per node, BOUND last); `state`
 * is an fp64 vector laid out by bpk_gmm_vb_layout (prior natural parameters in, posterior moments out);
 * P (N x K responsibilities) and gz (N, nullable) are written by every Z sweep.  ctrl / Lhist / tol / cap
 * as for bpk_pca_vb_run; a Lhist row is [Y, Z, mu, Lambda, alpha, total].  With a communicator the
 * statistics of every sweep are all-reduced on the library stream.                                   */
enum {
    BPK_GMMOP_Z = 1,       /* Z.update(): responsibilities + plate-summed statistics (bpk_gmm_sweep) */
    BPK_GMMOP_MU = 2,      /* mu.update()                                                            */
    BPK_GMMOP_LAMBDA = 3,  /* Lambda.update()                                                        */
    BPK_GMMOP_ALPHA = 4,   /* alpha.update()                                                         */
    BPK_GMMOP_BOUND = 5    /* lower bound + convergence test                                         */
};
int bpk_gmm_vb_layout(int D, int K, int64_t *offsets /* [nfields+1] */, int *nfields);
const char *bpk_gmm_vb_field_name(int i);
int bpk_gmm_vb_run(const double *Y, int64_t N, int D, int K, double *P, double *gz, double *state,
                   const int *ops, int nops, int niter, double tol, double *Lhist, int cap, int *ctrl);
/* bench: record these timers around the next n responsibilities-kernel launches of bpk_gmm_vb_run */
int bpk_gmm_vb_set_timers(const int *ids, int n);

#ifdef __cplusplus
}
#endif
#endif /* BPK_H */
