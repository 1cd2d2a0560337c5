// pca_vb.cu — the device-resident VB loop of the plated Gaussian factor model (Bayesian PCA)
//
//   X = GaussianARD(mu_x, a_x, plates=(1,N), shape=(K,))        "col" factor
//   C = GaussianARD(mu_c, alpha, plates=(M,1), shape=(K,))      "row" factor, alpha ~ Gamma(a0,b0) plates (K,)
//   Y = GaussianARD(SumMultiply('k,k->', X, C), tau),  tau ~ Gamma(ta0,tb0),  Y fully observed
//
// What it replaces: the Python scheduler loop of VB.update (vmp.py:132-172) together with the
// per-node update / lower-bound calls it makes for this model — X: gaussian.py:649-706 fed by
// dot.py:581; C: the same with the roles swapped; alpha, tau: gamma.py:96-148 fed by
// gaussian.py:609-637 and :2351-2371; bound: expfamily.py:400-480 for every node, then the
// convergence test of vmp.py:717-747.  The reference runs ~10^2 NumPy calls per node per sweep;
// the first GPU version of this engine issued ~100 tiny kernels per sweep from Python and was
// host-bound (3.2 ms of dispatch around a 2.1 ms sweep kernel).  Here one sweep is
//       [pca_xsweep_kernel]  ->  [pca_vb_small_kernel: STATS, C, alpha, tau, BOUND, XPRE]
// with nothing read back until a chunk of sweeps has been enqueued.  The update ORDER is the
// user's: the host passes the per-iteration program as a list of opcodes (one per node), and
// runs of small ops between two X sweeps are executed by one single-CTA launch.
//
// Convergence is decided ON DEVICE (same test as vmp.py:738-747) and raises a stop word that
// every later kernel of the chunk checks on entry, so the posterior freezes at exactly the
// iteration the reference would have returned at.
//
// All persistent quantities live in one fp64 state vector (layout: pca_vb_offsets below,
// exported through bpk_pca_vb_layout); after the run the node objects of the Python graph
// simply view slices of it.
#include "pca_vb_ops.cuh"
#include <vector>
#include <stdlib.h>

__global__ void __launch_bounds__(VB_THREADS, 1) pca_vb_small_kernel(PcaVbArgs p, size_t sm_doubles) {
    extern __shared__ double sm[];
    pca_vb_ops(p, sm, sm_doubles);
}

extern "C" int bpk_pca_vb_layout(int M, int K, int64_t *offsets, int *nfields) {
    if (M < 1 || K < 1) return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_layout: bad shape");
    if (offsets) pca_vb_offsets(M, K, offsets);
    if (nfields) *nfields = F_COUNT;
    return BPK_OK;
}
extern "C" const char *bpk_pca_vb_field_name(int i) { return (i >= 0 && i < F_COUNT) ? kFieldNames[i] : nullptr; }

#define BPK_VB_NDBG 192
static unsigned long long *g_vb_dbg = nullptr;
// BPK_VB_DEBUG=1: %globaltimer stamps of the LAST fused sweep launch: [0] kernel start, [1] data pass done
// (CTA 0), [2] after grid barrier 1, [3] CTA 0's share of the reduction done, [4] after grid barrier 2,
// [8+i] start of tail op i, [5] tail done; [64+it] the service CTA passing grid barrier 1 of sweep `it` of the launch.
extern "C" int bpk_debug_stamps(uint64_t *out, int n) {
    BPK_REQUIRE_INIT();
    if (!g_vb_dbg) return bpk_set_error(BPK_EINVAL, "no debug stamps (set BPK_VB_DEBUG=1 before the first run)");
    if (n > BPK_VB_NDBG) n = BPK_VB_NDBG;
    BPK_CUDA(cudaMemcpyAsync(out, g_vb_dbg, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, g_bpk.stream));
    BPK_CUDA(cudaStreamSynchronize(g_bpk.stream));
    return BPK_OK;
}
static int g_vb_timers[64];
static int g_vb_ntimers = 0, g_vb_timer_pos = 0;
extern "C" int bpk_pca_vb_set_timers(const int *ids, int n) {
    if (n < 0 || n > 64) return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_set_timers: at most 64 timers");
    for (int i = 0; i < n; ++i) g_vb_timers[i] = ids[i];
    g_vb_ntimers = n;
    g_vb_timer_pos = 0;
    return BPK_OK;
}
extern "C" int bpk_pca_vb_timers_used(void) { return g_vb_timer_pos; }

static int g_vb_no_loop = 0;
extern "C" int bpk_pca_vb_set_mode(int no_loop, int *prev) {
    if (prev) *prev = g_vb_no_loop;
    g_vb_no_loop = no_loop ? 1 : 0;
    return BPK_OK;
}

static int vb_launch_small(PcaVbArgs &a, size_t smem) {
    if (a.nops == 0) return BPK_OK;
    BPK_LAUNCH(pca_vb_small_kernel, 1, VB_THREADS, smem, a, smem / sizeof(double));
    a.nops = 0;
    a.partial = nullptr;
    a.nparts = 0;
    return BPK_OK;
}

extern "C" int bpk_pca_vb_run(const double *Y, int64_t M, int64_t N, int K, double *X, double *state,
                              const int *ops, int nops, int niter, int has_alpha, int has_tau, double tol,
                              double *Lhist, int cap, int *ctrl) {
    BPK_REQUIRE_INIT();
    if (M < 1 || K < 1 || K > BPK_MAXDIM || N < 0 || M > (1 << 20))
        return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: bad shape");
    if (nops < 1 || niter < 0) return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: empty program");
    int nranks = 1, rank = 0;
    bpk_comm_size(&nranks, &rank);
    int64_t off[F_COUNT + 1];
    pca_vb_offsets((int)M, K, off);
    size_t smem = pca_vb_smem_doubles(K) * sizeof(double);
    if (smem + (size_t)off[F_COUNT] * sizeof(double) <= (200u << 10)) smem += (size_t)off[F_COUNT] * sizeof(double);   // stage the state
    if (smem > (48u << 10))
        BPK_CUDA(cudaFuncSetAttribute(pca_vb_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool fast = (M <= PCA_MP && K <= PCA_KP);
    PcaVbArgs a;
    a.M = (int)M; a.K = K; a.has_alpha = has_alpha; a.has_tau = has_tau; a.tol = tol;
    a.st = state; a.partial = nullptr; a.nparts = 0; a.local_stats = nranks > 1;
    a.Lhist = Lhist; a.cap = cap; a.ctrl = ctrl; a.nops = 0;
    a.xranks = 1; a.xrank = 0;
    a.niter = 1; a.nops_last = 0; a.derive_sxx = 0;
    a.ll = 0;
    a.gj2 = getenv("BPK_VB_GJ1") ? 0 : 1;
    a.par = getenv("BPK_VB_SERIAL") ? 0 : 1;
    a.dry_every = getenv("BPK_VB_DRY_FIRST_ONLY") ? 0 : 1;
    a.dbg = nullptr;
    if (getenv("BPK_VB_DEBUG")) {
        if (!g_vb_dbg) {
            BPK_CUDA(cudaMalloc((void **)&g_vb_dbg, BPK_VB_NDBG * sizeof(unsigned long long)));
            BPK_CUDA(cudaMemsetAsync(g_vb_dbg, 0, BPK_VB_NDBG * sizeof(unsigned long long), g_bpk.stream));
            if (getenv("BPK_VB_DRYTEST")) {
                unsigned long long one = 1;
                BPK_CUDA(cudaMemcpyAsync(g_vb_dbg + 63, &one, 8, cudaMemcpyHostToDevice, g_bpk.stream));
                BPK_CUDA(cudaStreamSynchronize(g_bpk.stream));
            }
        }
        a.dbg = g_vb_dbg;
    }
    for (int r = 0; r < BPK_XCHG_MAXRANKS; ++r) a.xwin[r] = nullptr;
    a.xown = nullptr;
    const int64_t nstat = M * K + (int64_t)K * K + K;
    double *stats_dst = state + (nranks > 1 ? off[F_STATS_LOCAL] : off[F_STATS]);
    if (fast && N > 0) {
        const bool p2p = g_xchg.ready && g_xchg.nranks > 1 && PCA_NSTAT <= BPK_XCHG_CAP;
        if (p2p) {
            a.xranks = g_xchg.nranks; a.xrank = g_xchg.rank; a.local_stats = 0;
            for (int r = 0; r < g_xchg.nranks; ++r) a.xwin[r] = g_xchg.win[r];
            a.xown = g_xchg.win[g_xchg.rank];
        }
        // Fused launches hand the reduced statistics to CTA 0 as LL packets through a window (the peers' windows
        // with p2p, this GPU's own otherwise): no second grid barrier.  BPK_VB_NO_LL=1 (single rank only) keeps the
        // scratch buffer + grid barrier hand-off for A/B measurements.
        bool ll_ok = p2p;
        if (nranks == 1 && !getenv("BPK_VB_NO_LL")) {
            int rc = bpk_xchg_local();
            if (rc) return rc;
            a.xwin[0] = g_xchg.own;
            a.xown = g_xchg.own;
            ll_ok = true;
        }
        // One launch per sweep: the small ops that follow an XSWEEP (up to the next one) ride in the
        // tail of the sweep kernel; only what precedes the first sweep of the run is a launch of its own.
        // Whole chunk in ONE launch when the program has a single XSWEEP per iteration and the exchange (if any)
        // happens inside the kernel: ops after the sweep (+ the ops that precede the next one) form the tail.
        {
            int nx = 0, kx = -1;
            for (int i = 0; i < nops; ++i)
                if (ops[i] == BPK_VBOP_XSWEEP) { ++nx; kx = i; }
            const bool loopable = nx == 1 && (nranks == 1 || p2p) && nops - 1 <= VB_MAXOPS && niter >= 1 &&
                                  pca_ws_available(Y, N) && !g_vb_no_loop && !getenv("BPK_VB_NO_LOOP");
            if (loopable) {
                for (int i = 0; i < nops; ++i)
                    if (ops[i] < BPK_VBOP_XSWEEP || ops[i] > BPK_VBOP_BOUND || (ops[i] == BPK_VBOP_STATS && i < kx))
                        return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: bad opcode %d at %d", ops[i], i);
                for (int i = 0; i < kx; ++i) a.ops[a.nops++] = ops[i];          // head of the first sweep
                int rc = vb_launch_small(a, smem);
                if (rc) return rc;
                PcaVbArgs tail = a;
                tail.nops = 0;
                for (int i = kx + 1; i < nops; ++i) tail.ops[tail.nops++] = ops[i];
                tail.nops_last = tail.nops;
                for (int i = 0; i < kx; ++i) tail.ops[tail.nops++] = ops[i];
                tail.niter = niter;
                tail.derive_sxx = getenv("BPK_PCA_SXX_DMMA") ? 0 : 1;
                tail.ll = (ll_ok && tail.nops_last > 0 && tail.ops[0] == BPK_VBOP_STATS) ? 1 : 0;
                if (p2p && !tail.ll) return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: STATS must follow XSWEEP");
                int tid = -1;
                if (g_vb_timer_pos < g_vb_ntimers) tid = g_vb_timers[g_vb_timer_pos++];
                if (tid >= 0) bpk_timer_record(tid, 0);
                double *partial = nullptr;
                int nparts = 0, tail_done = 0;
                rc = pca_xsweep_partials(Y, M, N, K, state + off[F_A], state + off[F_BX], X, ctrl + 1, &partial, &nparts,
                                         &tail, &tail_done);
                if (rc) return rc;
                if (tid >= 0) bpk_timer_record(tid, 1);
                if (!tail_done) return bpk_set_error(BPK_ECUDA, "bpk_pca_vb_run: fused sweep kernel was not used");
                return BPK_OK;
            }
        }
        std::vector<int> seq;
        seq.reserve((size_t)niter * nops);
        for (int it = 0; it < niter; ++it)
            for (int i = 0; i < nops; ++i) {
                const int op = ops[i];
                if (op < BPK_VBOP_XSWEEP || op > BPK_VBOP_BOUND)
                    return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: unknown opcode %d", op);
                seq.push_back(op);
            }
        size_t i = 0;
        while (i < seq.size()) {
            const int op = seq[i];
            if (op != BPK_VBOP_XSWEEP) {
                if (op == BPK_VBOP_STATS) return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: STATS without a preceding XSWEEP");
                a.ops[a.nops++] = op;
                ++i;
                if (a.nops == VB_MAXOPS) { int rc = vb_launch_small(a, smem); if (rc) return rc; }
                continue;
            }
            int rc = vb_launch_small(a, smem);
            if (rc) return rc;
            PcaVbArgs tail = a;
            tail.nops = 0;
            size_t j = i + 1;
            bool exchange = false;
            while (j < seq.size() && seq[j] != BPK_VBOP_XSWEEP && tail.nops < VB_MAXOPS) {
                tail.ops[tail.nops++] = seq[j];
                ++j;
                if (seq[j - 1] == BPK_VBOP_STATS && nranks > 1 && !p2p) { exchange = true; break; }
            }
            tail.nops_last = tail.nops;
            tail.niter = 1;
            tail.ll = (ll_ok && !exchange && tail.nops > 0 && tail.ops[0] == BPK_VBOP_STATS && pca_ws_available(Y, N)) ? 1 : 0;
            int tid = -1;
            if (g_vb_timer_pos < g_vb_ntimers) tid = g_vb_timers[g_vb_timer_pos++];
            if (tid >= 0) bpk_timer_record(tid, 0);
            double *partial = nullptr;
            int nparts = 0, tail_done = 0;
            rc = pca_xsweep_partials(Y, M, N, K, state + off[F_A], state + off[F_BX], X, ctrl + 1, &partial, &nparts,
                                     tail.nops ? &tail : nullptr, &tail_done);
            if (rc) return rc;
            if (tid >= 0) bpk_timer_record(tid, 1);
            if (!tail_done && tail.nops) {       // unaligned input: the 8-warp kernel has no tail
                a = tail;
                a.partial = partial;
                a.nparts = nparts;
                rc = vb_launch_small(a, smem);
                if (rc) return rc;
            }
            if (exchange) {
                rc = bpk_allreduce_sum_f64_oop(state + off[F_STATS_LOCAL], state + off[F_STATS], (uint64_t)nstat);
                if (rc) return rc;
            }
            i = j;
        }
        return vb_launch_small(a, smem);
    }
    for (int it = 0; it < niter; ++it) {
        for (int i = 0; i < nops; ++i) {
            const int op = ops[i];
            if (op == BPK_VBOP_XSWEEP) {
                int rc = vb_launch_small(a, smem);
                if (rc) return rc;
                if (N == 0) {
                    BPK_CUDA(cudaMemsetAsync(stats_dst, 0, nstat * sizeof(double), g_bpk.stream));
                    continue;
                }
                int tid = -1;
                if (g_vb_timer_pos < g_vb_ntimers) tid = g_vb_timers[g_vb_timer_pos++];
                if (tid >= 0) bpk_timer_record(tid, 0);
                if (fast) {
                    double *partial = nullptr;
                    int nparts = 0;
                    rc = pca_xsweep_partials(Y, M, N, K, state + off[F_A], state + off[F_BX], X, ctrl + 1,
                                             &partial, &nparts, nullptr, nullptr);
                    if (rc) return rc;
                    a.partial = partial;
                    a.nparts = nparts;
                } else {
                    // generic shapes: host decides per iteration (the caller uses niter == 1)
                    BPK_CUDA(cudaMemsetAsync(stats_dst, 0, nstat * sizeof(double), g_bpk.stream));
                    rc = bpk_pca_xsweep(Y, M, N, K, state + off[F_A], state + off[F_BX], X, stats_dst);
                    if (rc) return rc;
                }
                if (tid >= 0) bpk_timer_record(tid, 1);
            } else if (op == BPK_VBOP_STATS) {
                if (fast && N > 0) a.ops[a.nops++] = op;
                if (nranks > 1) {
                    int rc = vb_launch_small(a, smem);
                    if (rc) return rc;
                    rc = bpk_allreduce_sum_f64_oop(state + off[F_STATS_LOCAL], state + off[F_STATS], (uint64_t)nstat);
                    if (rc) return rc;
                }
            } else if (op >= BPK_VBOP_SXXT && op <= BPK_VBOP_BOUND) {
                a.ops[a.nops++] = op;
            } else {
                return bpk_set_error(BPK_EINVAL, "bpk_pca_vb_run: unknown opcode %d", op);
            }
            if (a.nops == VB_MAXOPS) {
                int rc = vb_launch_small(a, smem);
                if (rc) return rc;
            }
        }
    }
    return vb_launch_small(a, smem);
}
