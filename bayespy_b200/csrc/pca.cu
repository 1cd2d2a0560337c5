// pca.cu — the fused one-pass sweep of the plated Gaussian factor model
//   y[m,n] ~ N(w_m . x_n, 1/tau)      (Bayesian PCA)
//
// Reference path being replaced (per VB sweep, all Python/NumPy):
//   X.update():   dot.py:581 einsum messages -> gaussian.py:672-706 moments
//   C.update():   dot.py:581 again with the roles swapped (sum over n)
//   tau.update(): dot.py:355,403 <f>,<f^2> over (M,N) -> gaussian.py:2361-2369
//   bound:        expfamily.py:400-480 over (M,N) again
// Here: ONE pass over Y per sweep.  For every column n
//   x_n = A y_n + b                       (A = tau Cov_x <W>^T, K x M)
// is written out, and the plate-summed sufficient statistics that every other
// node of the model needs (the einsum reductions over n of dot.py:581 and
// the <f>,<f^2> sums of dot.py:355,403) are accumulated on the fly:
//   S_yx = sum_n y_n x_n^T (M x K),  S_xx = sum_n x_n x_n^T (K x K),  s_x = sum_n x_n.
// Algorithmic traffic: M*8 B read + K*8 B written per column (640 B at 64x16).
//
// Two kernels implement it (both: persistent grid, one CTA per SM, both GEMM-shaped
// contractions on the fp64 tensor pipe with mma.sync.m8n8k4.f64 — tcgen05 has no fp64
// kind —, per-CTA partial statistics reduced in a fixed order: deterministic, no atomics):
//   pca_xsweep_ws_kernel  (16-byte aligned inputs, the normal case; further down) 16 warps,
//       warp-specialised X / S roles, TMA tensor-map loads with 128B swizzle, mbarrier ring;
//       as <FUSED> it is the whole device-resident VB loop (pca_vb_ops.cuh tail, grid barriers).
//   pca_xsweep_kernel     (v1, kept for odd N / unaligned Y) 8 warps of 220 registers, Y tiles of
//       64 x T doubles through a cp.async ring (padded pitch T+4), A as 32 register fragments,
//       each warp a private S_yx/S_xx accumulator set, grid reduction by pca_stats_final_kernel.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include "pca_common.cuh"
#include "spd.cuh"
#include "pca_vb_ops.cuh"

#define PCA_WARPS 8
#define PCA_LDX 20     // pitch of the per-warp X staging tile

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *g, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *g, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int NT, bool ALIGN16>
__device__ __forceinline__ void pca_issue_tile(double *Ys, const double *__restrict__ Y, int64_t M, int64_t N,
                                               int64_t n0) {
    constexpr int T = PCA_WARPS * NT, LDY = T + 4;
    if (ALIGN16) {
        constexpr int CH = T / 2;   // 16-byte chunks per row
        for (int e = threadIdx.x; e < (int)M * CH; e += PCA_WARPS * 32) {
            int m = e / CH, c = e - m * CH;
            int64_t n = n0 + 2 * c;
            const double *src = Y + (int64_t)m * N + (n < N ? n : 0);
            cp_async16(Ys + m * LDY + 2 * c, src, n < N ? 16 : 0);
        }
    } else {
        for (int e = threadIdx.x; e < (int)M * T; e += PCA_WARPS * 32) {
            int m = e / T, c = e - m * T;
            int64_t n = n0 + c;
            const double *src = Y + (int64_t)m * N + (n < N ? n : 0);
            cp_async8(Ys + m * LDY + c, src, n < N ? 8 : 0);
        }
    }
}

template <int NT, int STAGES, bool ALIGN16, bool COMPUTE_X>
__global__ void __launch_bounds__(PCA_WARPS * 32, 1)
pca_xsweep_kernel(const double *__restrict__ Y, int64_t M, int64_t N, int K,
                  const double *__restrict__ A, const double *__restrict__ bvec,
                  double *__restrict__ X, double *__restrict__ partial, int64_t ntiles,
                  const int *__restrict__ stop) {
    constexpr int T = PCA_WARPS * NT, LDY = T + 4, CB = NT / 8, NS = NT / 4;
    if (stop && *stop) return;      // resident VB loop: converged, leave state frozen
    extern __shared__ __align__(16) double smem[];
    double *Ysm = smem;                                            // [STAGES][64][LDY]
    double *Xsm = smem + (size_t)STAGES * PCA_MP * LDY;            // [WARPS][NT][LDX]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gr = lane >> 2, tg = lane & 3;                       // mma groupID / threadID_in_group
    double *Xs = Xsm + (size_t)w * NT * PCA_LDX;

    // zero the whole ring once: padded rows m >= M stay zero for the kernel's lifetime
    for (int e = threadIdx.x; e < STAGES * PCA_MP * LDY; e += PCA_WARPS * 32) Ysm[e] = 0.0;
    __syncthreads();

    // A fragments (row-major 8x4 blocks of the K x M matrix), zero padded
    double afrag[2][16];
    if (COMPUTE_X) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int ms = 0; ms < 16; ++ms) {
                int k = kb * 8 + gr, m = ms * 4 + tg;
                afrag[kb][ms] = (k < K && m < M) ? A[(int64_t)k * M + m] : 0.0;
            }
    }
    double bk[2] = {0.0, 0.0};
    if (COMPUTE_X && bvec) {
        if (gr < K) bk[0] = bvec[gr];
        if (8 + gr < K) bk[1] = bvec[8 + gr];
    }

    double syx[8][2][2];   // [mb][kb][j]
    double sxx[3][2];      // blocks (0,0),(0,1),(1,1)
    double sx = 0.0;       // lanes 0..15: sum_n x[n][lane]
#pragma unroll
    for (int mb = 0; mb < 8; ++mb)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) syx[mb][kb][0] = syx[mb][kb][1] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) sxx[i][0] = sxx[i][1] = 0.0;

    // prologue: STAGES-1 tiles in flight
    const int64_t first = blockIdx.x, stride = gridDim.x;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        int64_t t = first + (int64_t)s * stride;
        if (t < ntiles) pca_issue_tile<NT, ALIGN16>(Ysm + (size_t)s * PCA_MP * LDY, Y, M, N, t * T);
        cp_async_commit();
    }

    int it = 0;
    for (int64_t tile = first; tile < ntiles; tile += stride, ++it) {
        {   // keep the ring full
            int64_t tn = tile + (int64_t)(STAGES - 1) * stride;
            if (tn < ntiles)
                pca_issue_tile<NT, ALIGN16>(Ysm + (size_t)((it + STAGES - 1) % STAGES) * PCA_MP * LDY, Y, M, N, tn * T);
            cp_async_commit();
        }
        cp_async_wait<STAGES - 1>();
        __syncthreads();
        const double *Ys = Ysm + (size_t)(it % STAGES) * PCA_MP * LDY;
        const int c0 = w * NT;                     // this warp's first column within the tile
        const int64_t nbase = tile * T + c0;       // global column of Xs row 0

        if (COMPUTE_X) {
            double acc[2][CB][2];
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) acc[kb][cb][0] = acc[kb][cb][1] = bk[kb];
#pragma unroll
            for (int ms = 0; ms < 16; ++ms) {
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) {
                    double yb = Ys[(ms * 4 + tg) * LDY + c0 + cb * 8 + gr];
                    dmma884(acc[0][cb][0], acc[0][cb][1], afrag[0][ms], yb);
                    dmma884(acc[1][cb][0], acc[1][cb][1], afrag[1][ms], yb);
                }
            }
            // stage x (zeroed beyond N so that padded columns add nothing to the statistics)
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        int nl = cb * 8 + 2 * tg + j;
                        double v = (nbase + nl < N) ? acc[kb][cb][j] : 0.0;
                        Xs[nl * PCA_LDX + kb * 8 + gr] = v;
                    }
            __syncwarp();
            // coalesced store of the NT x K block
            for (int e = lane; e < NT * K; e += 32) {
                int nl = e / K, k = e - nl * K;
                if (nbase + nl < N) X[(nbase + nl) * K + k] = Xs[nl * PCA_LDX + k];
            }
        } else {
            for (int e = lane; e < NT * PCA_KP; e += 32) {
                int nl = e >> 4, k = e & 15;
                Xs[nl * PCA_LDX + k] = (nbase + nl < N && k < K) ? X[(nbase + nl) * K + k] : 0.0;
            }
            __syncwarp();
        }

        if (lane < PCA_KP) {
#pragma unroll
            for (int nl = 0; nl < NT; ++nl) sx += Xs[nl * PCA_LDX + lane];
        }

#pragma unroll
        for (int ns = 0; ns < NS; ++ns) {
            double xf0 = Xs[(ns * 4 + tg) * PCA_LDX + gr];
            double xf1 = Xs[(ns * 4 + tg) * PCA_LDX + 8 + gr];
#pragma unroll
            for (int mb = 0; mb < 8; ++mb) {
                double ya = Ys[(mb * 8 + gr) * LDY + c0 + ns * 4 + tg];
                dmma884(syx[mb][0][0], syx[mb][0][1], ya, xf0);
                dmma884(syx[mb][1][0], syx[mb][1][1], ya, xf1);
            }
            dmma884(sxx[0][0], sxx[0][1], xf0, xf0);
            dmma884(sxx[1][0], sxx[1][1], xf0, xf1);
            dmma884(sxx[2][0], sxx[2][1], xf1, xf1);
        }
        __syncthreads();   // everyone is done with this stage before it is refilled
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- warp -> CTA reduction through smem (reuses the ring) ----
    double *red = smem;    // [WARPS][NSTAT]
    double *mine = red + (size_t)w * PCA_NSTAT;
#pragma unroll
    for (int mb = 0; mb < 8; ++mb)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                mine[(mb * 8 + gr) * PCA_KP + kb * 8 + 2 * tg + j] = syx[mb][kb][j];
    double *mxx = mine + PCA_MP * PCA_KP;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        mxx[gr * PCA_KP + 2 * tg + j] = sxx[0][j];
        mxx[gr * PCA_KP + 8 + 2 * tg + j] = sxx[1][j];
        mxx[(8 + 2 * tg + j) * PCA_KP + gr] = sxx[1][j];      // symmetric block
        mxx[(8 + gr) * PCA_KP + 8 + 2 * tg + j] = sxx[2][j];
    }
    if (lane < PCA_KP) mine[PCA_MP * PCA_KP + PCA_KP * PCA_KP + lane] = sx;
    __syncthreads();
    double *pout = partial + (size_t)blockIdx.x * PCA_NSTAT;
    for (int e = threadIdx.x; e < PCA_NSTAT; e += PCA_WARPS * 32) {
        double s = 0.0;
#pragma unroll
        for (int ww = 0; ww < PCA_WARPS; ++ww) s += red[(size_t)ww * PCA_NSTAT + e];
        pout[e] = s;
    }
}

// =====================================================================================
// v2: warp-specialised sweep fed by TMA (aligned inputs).  16 warps per SM instead of 8:
//   X-warps 0..7  hold the A fragments, compute x = A y + b for their 1/8 of the tile's
//                 columns on the fp64 tensor pipe, store x to HBM straight from the
//                 accumulators, and hand the tile to their partner through shared memory;
//   S-warps 8..15 hold the S_yx / S_xx accumulators and run the second contraction.
// Each role fits 128 registers.  Y tiles (64 rows x T columns) arrive as T/16 TMA boxes of
// 64 x 16 doubles (cp.async.bulk.tensor.2d, SWIZZLE_128B, zero fill out of bounds) that
// complete on per-stage "full" mbarriers; stages are released through "empty" mbarriers and
// each X/S warp pair hands x tiles over through its own double-buffered mbarrier pair — no
// CTA-wide barrier inside the main loop.  The shared-memory image of a box is dense
// (row pitch 128 B, 16-byte chunk index XOR row%8); the k-slot -> row map of the first
// contraction (m = 8*(ms>>1) + 2*tg + (ms&1)) and the row map of the second
// (m = 8*mb + 2*(gr&3) + (gr>>2)) are chosen so that every fragment load hits 32 distinct
// banks per half-warp under that swizzle.
// (A first version used 64 cp.async.bulk row copies of 512 B per tile into padded rows: it
// was limited by the per-copy cost of the TMA unit, 2.8 ms; profiles/r01_*.)
// =====================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nWAIT_LOOP:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

#define WS_PAIRS 8
#define WS_SKIP 64                       // tiles CTA 0 gives away in the fused kernel (~1.4 us each)
#define WS_BOXC 16                       // columns per TMA box (128 B: the SWIZZLE_128B span)
#define WS_BOX (PCA_MP * WS_BOXC)        // doubles per box

// offset (in doubles) of element (row m, column c) inside a stage of T/16 swizzled boxes
__device__ __forceinline__ int ws_off(int m, int c) {
    return (c >> 4) * WS_BOX + m * WS_BOXC + ((((c & 15) >> 1) ^ (m & 7)) << 1) + (c & 1);
}

// Self-resetting grid barrier (all CTAs are co-resident: one persistent CTA per SM).
// bar[0] = arrival count, bar[1] = epoch.
// Spin with a watchdog: if the grid is not fully co-resident (another context on the GPU) or a CTA died, give up
// after ~60 s, raise error bit 8 in *err and let the kernel terminate instead of hanging the device.
__device__ __forceinline__ void grid_spin(unsigned int *bar, unsigned int target, int *err) {
    const long long t0 = clock64();
    while (*(volatile unsigned int *)&bar[1] != target) {
        __nanosleep(32);
        if (clock64() - t0 > 120000000000ll) {     // ~60 s: longer than the peer-exchange watchdog of CTA 0
            if (err) atomicOr(err, 8);
            break;
        }
    }
}
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &epoch, int *err = nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int target = epoch + 1;
        if (atomicAdd(&bar[0], 1u) == gridDim.x - 1) {
            atomicExch(&bar[0], 0u);
            __threadfence();
            atomicExch(&bar[1], target);
        } else {
            grid_spin(bar, target, err);
        }
        __threadfence();
    }
    epoch += 1;
    __syncthreads();
}

__device__ __forceinline__ void grid_arrive(unsigned int *bar, unsigned int epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&bar[0], 1u) == gridDim.x - 1) {
            atomicExch(&bar[0], 0u);
            __threadfence();
            atomicExch(&bar[1], epoch + 1);
        }
    }
}
__device__ __forceinline__ void grid_wait(unsigned int *bar, unsigned int &epoch, int *err = nullptr) {
    if (threadIdx.x == 0) {
        grid_spin(bar, epoch + 1, err);
        __threadfence();
    }
    epoch += 1;
    __syncthreads();
}

// FUSED: the launch is one whole VB sweep of the resident loop — after the data pass the grid
// reduces its partial statistics in place (two grid barriers, fixed order) and CTA 0 runs the
// sweep's small ops (pca_vb_ops: STATS .. BOUND, XPRE of the next sweep) before the kernel ends.
// DERIVE_SXX (fused launches only): the S-warps skip the three S_xx DMMAs per 4 columns; since x_n = A y_n + b
// exactly, sum x x^T = A (sum y x^T) + b (sum x)^T is formed from S_yx and s_x in the tail (pca_vb_ops, STATS),
// after the cross-rank exchange — 8.6 % less work on the binding fp64 pipe, still one pass over Y per sweep.
template <int NT, int STAGES, int DIST, bool COMPUTE_X, bool FUSED, bool DERIVE_SXX>
__global__ void __launch_bounds__(2 * WS_PAIRS * 32, 1)
pca_xsweep_ws_kernel(const __grid_constant__ CUtensorMap tmapY, int64_t M, int64_t N, int K,
                     const double *A, const double *bvec,
                     double *__restrict__ X, double *partial, int64_t ntiles,
                     const int *stop, unsigned int *gbar, const __grid_constant__ PcaVbArgs vb,
                     size_t vb_sm_doubles) {
    constexpr int T = WS_PAIRS * NT, NS = NT / 4, CB = NT / 8, NBOX = T / WS_BOXC, STG = NBOX * WS_BOX;
    static_assert(DIST >= 1 && DIST < STAGES, "prefetch distance must leave one stage for the consumers");
    if (stop && *stop) return;
    if (FUSED && blockIdx.x == 0) vb_stamp(vb.dbg, 0);
    extern __shared__ __align__(1024) double smem_ws[];
    double *smem = smem_ws;
    double *Ysm = smem;                                              // [STAGES][NBOX][64][16] swizzled
    double *Xsm = Ysm + (size_t)STAGES * STG;                        // [PAIRS][2][NT][LDX]
    uint64_t *bars = (uint64_t *)(Xsm + (size_t)WS_PAIRS * 2 * NT * PCA_LDX);
    uint64_t *full = bars, *empty = bars + STAGES;
    uint64_t *xfull = bars + 2 * STAGES, *xfree = xfull + 2 * WS_PAIRS;   // [pair][2]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const bool is_x = w < WS_PAIRS;
    const int p = w & (WS_PAIRS - 1);

    unsigned int epoch = 0;                       // grid-barrier epoch (FUSED); thread 0's copy is the one used
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2 * WS_PAIRS); }
        for (int i = 0; i < 2 * WS_PAIRS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xfree[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmapY) : "memory");
        if (FUSED) epoch = *(volatile unsigned int *)&gbar[1];
    }
    __syncthreads();

    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t ntl_own = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;   // round-robin share
    // FUSED: CTA 0 runs the sweep's small ops after the grid reduction.  That code executes once per launch
    // from a cold instruction cache (60 us cold, 30 us warm), so CTA 0 hands its last `skip` tiles to the
    // CTAs at the other end of the grid (one each), finishes early and spends the head start on a dry run
    // of the ops (no side effects) while everybody else is still streaming.  Static, hence deterministic.
    int64_t skip = 0;
    if (FUSED && vb.nops > 0 && stride > 1) {
        const int64_t ntl0 = (ntiles + stride - 1) / stride;
        skip = ntl0 / 2 < WS_SKIP ? ntl0 / 2 : WS_SKIP;
        if (skip > stride - 1) skip = stride - 1;
    }
    const int64_t xtra = stride - 1 - first;          // CTA G-1 takes CTA 0's first surplus tile, G-2 the next ...
    const bool has_xtra = skip > 0 && first != 0 && xtra < skip;
    const int64_t ntl = (first == 0 ? ntl_own - skip : ntl_own) + (has_xtra ? 1 : 0);
    const int64_t xtra_tile = ((ntiles + stride - 1) / stride - skip + xtra) * stride;     // a tile of CTA 0's list
    auto tile_of = [&](int64_t i) -> int64_t { return (has_xtra && i == ntl - 1) ? xtra_tile : first + i * stride; };

    // FUSED: the launch runs vb.niter whole sweeps.  The mbarrier phases simply keep counting across sweeps
    // (tbase = tiles this CTA has consumed so far), so nothing is re-initialised between them.
    const int niter = FUSED ? vb.niter : 1;
    int64_t tbase = 0;
    auto issue = [&](int64_t jl) {       // lanes 0..NBOX-1 of warp 0: one TMA box each
        const int64_t j = tbase + jl;
        const int slot = (int)(j % STAGES);
        if (j >= STAGES) mbar_wait(&empty[slot], (uint32_t)((j / STAGES - 1) & 1));
        const int64_t n0 = tile_of(jl) * T;
        if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)(STG * sizeof(double)));
        __syncwarp();
        if (lane < NBOX)
            tma_load_2d(Ysm + (size_t)slot * STG + lane * WS_BOX, &tmapY, (int)(n0 + lane * WS_BOXC), 0, &full[slot]);
    };

    // BPK_VB_DEBUG: the last CTA of the grid (a full data pass plus one of CTA 0's tiles) stamps the phases of the
    // second-to-last sweep of the launch into dbg[48..57] (tools/vb_tail_timing.py)
    const bool probe = FUSED && vb.dbg != nullptr && blockIdx.x == gridDim.x - 1;
    for (int it = 0; it < niter; ++it, tbase += ntl) {
    const bool pr = probe && (it == niter - 2 || niter == 1);
    if (probe && it == niter - 1 && niter > 1) vb_stamp(vb.dbg, 57);
    if (pr) vb_stamp(vb.dbg, 48);
    if (it > 0) {
        if (*(volatile const int *)stop) break;                       // converged inside this launch (uniform)
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");    // ring was scratch for the tail: order before TMA
    }
    if (is_x) {
        double afrag[2][16];
        double bk[2] = {0.0, 0.0};
        if (COMPUTE_X) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int ms = 0; ms < 16; ++ms) {
                    int k = kb * 8 + gr, m = (ms >> 1) * 8 + 2 * tg + (ms & 1);
                    afrag[kb][ms] = (k < K && m < M) ? __ldcg(A + (int64_t)k * M + m) : 0.0;   // rewritten by the tail of every sweep
                }
            if (bvec) {
                if (gr < K) bk[0] = __ldcg(bvec + gr);
                if (8 + gr < K) bk[1] = __ldcg(bvec + 8 + gr);
            }
        }
        if (pr && w == 0) vb_stamp_lane(vb.dbg, 49);
        if (w == 0)
            for (int64_t j = 0; j < DIST && j < ntl; ++j) issue(j);
        for (int64_t i = 0; i < ntl; ++i) {
            if (w == 0 && i + DIST < ntl) issue(i + DIST);
            const int64_t ig = tbase + i;                 // tiles consumed by this CTA since the launch began
            const int slot = (int)(ig % STAGES), b = (int)(ig & 1);
            const int c0 = p * NT;
            const int64_t nbase = tile_of(i) * T + c0;
            double *Xs = Xsm + ((size_t)(p * 2 + b) * NT) * PCA_LDX;
            if (COMPUTE_X) {
                mbar_wait(&full[slot], (uint32_t)((ig / STAGES) & 1));
                if (pr && w == 0 && i == 0) vb_stamp_lane(vb.dbg, 50);
                const double *Ys = Ysm + (size_t)slot * STG;
                // two partial sums over the m blocks per accumulator: 4*CB independent DMMA chains
                double acc[2][2][CB][2];
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) {
                        acc[0][kb][cb][0] = acc[0][kb][cb][1] = bk[kb];
                        acc[1][kb][cb][0] = acc[1][kb][cb][1] = 0.0;
                    }
#pragma unroll
                for (int ms = 0; ms < 16; ++ms) {
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) {
                        double yb = Ys[ws_off((ms >> 1) * 8 + 2 * tg + (ms & 1), c0 + cb * 8 + gr)];
                        dmma884(acc[ms & 1][0][cb][0], acc[ms & 1][0][cb][1], afrag[0][ms], yb);
                        dmma884(acc[ms & 1][1][cb][0], acc[ms & 1][1][cb][1], afrag[1][ms], yb);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);           // this warp no longer reads the Y stage
                if (ig >= 2) mbar_wait(&xfree[p * 2 + b], (uint32_t)(((ig >> 1) - 1) & 1));
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int nl = cb * 8 + 2 * tg + j, k = kb * 8 + gr;
                            const bool inb = nbase + nl < N;
                            double v = inb ? acc[0][kb][cb][j] + acc[1][kb][cb][j] : 0.0;   // padded columns add nothing
                            Xs[nl * PCA_LDX + k] = v;
                            if (inb && k < K) X[(nbase + nl) * K + k] = v;
                        }
            } else {
                if (lane == 0) {
                    mbar_wait(&full[slot], (uint32_t)((ig / STAGES) & 1));
                    mbar_arrive(&empty[slot]);
                }
                if (ig >= 2) mbar_wait(&xfree[p * 2 + b], (uint32_t)(((ig >> 1) - 1) & 1));
                for (int e = lane; e < NT * PCA_KP; e += 32) {
                    int nl = e >> 4, k = e & 15;
                    Xs[nl * PCA_LDX + k] = (nbase + nl < N && k < K) ? X[(nbase + nl) * K + k] : 0.0;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&xfull[p * 2 + b]);
        }
        if (pr && w == 0) vb_stamp_lane(vb.dbg, 51);
    } else {
        const int rr = 2 * (gr & 3) + (gr >> 2);      // S_yx row within an 8-row block held by this lane group
        double syx[8][2][2];
        double sxx[3][2];
        double sx = 0.0;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb)
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) syx[mb][kb][0] = syx[mb][kb][1] = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) sxx[i][0] = sxx[i][1] = 0.0;
        for (int64_t i = 0; i < ntl; ++i) {
            const int64_t ig = tbase + i;
            const int slot = (int)(ig % STAGES), b = (int)(ig & 1);
            const int c0 = p * NT;
            const double *Ys = Ysm + (size_t)slot * STG;
            const double *Xs = Xsm + ((size_t)(p * 2 + b) * NT) * PCA_LDX;
            mbar_wait(&full[slot], (uint32_t)((ig / STAGES) & 1));
            mbar_wait(&xfull[p * 2 + b], (uint32_t)((ig >> 1) & 1));
            if (lane < PCA_KP) {
#pragma unroll
                for (int nl = 0; nl < NT; ++nl) sx += Xs[nl * PCA_LDX + lane];
            }
#pragma unroll
            for (int ns = 0; ns < NS; ++ns) {
                double xf0 = Xs[(ns * 4 + tg) * PCA_LDX + gr];
                double xf1 = Xs[(ns * 4 + tg) * PCA_LDX + 8 + gr];
#pragma unroll
                for (int mb = 0; mb < 8; ++mb) {
                    double ya = Ys[ws_off(mb * 8 + rr, c0 + ns * 4 + tg)];
                    dmma884(syx[mb][0][0], syx[mb][0][1], ya, xf0);
                    dmma884(syx[mb][1][0], syx[mb][1][1], ya, xf1);
                }
                if (!DERIVE_SXX) {
                    dmma884(sxx[0][0], sxx[0][1], xf0, xf0);
                    dmma884(sxx[1][0], sxx[1][1], xf0, xf1);
                    dmma884(sxx[2][0], sxx[2][1], xf1, xf1);
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&xfree[p * 2 + b]); mbar_arrive(&empty[slot]); }
        }
        if (pr && w == WS_PAIRS) vb_stamp_lane(vb.dbg, 52);
        // every issued tile has been consumed by every warp once all warps pass the barrier below
        __syncthreads();
        double *mine = smem + (size_t)p * PCA_NSTAT;
#pragma unroll
        for (int mb = 0; mb < 8; ++mb)
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    mine[(mb * 8 + rr) * PCA_KP + kb * 8 + 2 * tg + j] = syx[mb][kb][j];
        double *mxx = mine + PCA_MP * PCA_KP;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            mxx[gr * PCA_KP + 2 * tg + j] = sxx[0][j];
            mxx[gr * PCA_KP + 8 + 2 * tg + j] = sxx[1][j];
            mxx[(8 + 2 * tg + j) * PCA_KP + gr] = sxx[1][j];
            mxx[(8 + gr) * PCA_KP + 8 + 2 * tg + j] = sxx[2][j];
        }
        if (lane < PCA_KP) mine[PCA_MP * PCA_KP + PCA_KP * PCA_KP + lane] = sx;
    }
    if (is_x) __syncthreads();     // pairs with the S-warps' barrier above
    __syncthreads();
    double *pout = partial + (size_t)blockIdx.x * PCA_NSTAT;
    for (int e = threadIdx.x; e < PCA_NSTAT; e += blockDim.x) {
        double s = 0.0;
#pragma unroll
        for (int ww = 0; ww < WS_PAIRS; ++ww) s += smem[(size_t)ww * PCA_NSTAT + e];
        pout[e] = s;
    }
    if (FUSED) {
        if (blockIdx.x == 0) vb_stamp(vb.dbg, 1);
        if (pr) vb_stamp(vb.dbg, 53);
        const int nops_it = (it == niter - 1) ? vb.nops_last : vb.nops;
        grid_arrive(gbar, epoch);
        if (blockIdx.x == 0 && skip > 0 && (it == 0 || vb.dry_every)) pca_vb_ops(vb, smem, vb_sm_doubles, true, nops_it);   // warm the instruction cache
        grid_wait(gbar, epoch, vb.ctrl + 2);
        if (blockIdx.x == 0) vb_stamp(vb.dbg, 2);
        if (pr) vb_stamp(vb.dbg, 54);
        // distributed, fixed-order reduction over the CTAs: CTA c owns elements [c*per, (c+1)*per)
        const int per = (PCA_NSTAT + gridDim.x - 1) / gridDim.x;
        const int e0 = blockIdx.x * per;
        if (vb.ll) {
            // ... and the sweep's one exchange rides on it: the owner of a slice pushes its sum straight into every
            // rank's peer-memory window (lane r -> rank r, this GPU's own window included) as self-tagged LL packets.
            // No second grid barrier and no flag: CTA 0 (here and on every peer) gathers the packets in STATS as they
            // land, so the exchange costs one NVLink one-way latency instead of a single-CTA deposit + fence + flag
            // round trip.
            // exchanges this rank has completed so far = word 0 of its own window, advanced by CTA 0's STATS before the
            // grid barrier that ended the previous sweep (or the previous launch)
            const unsigned long long xseq = *(volatile const unsigned long long *)vb.xown + 1ull;
            const int par = (int)(xseq & 1ull);
            const unsigned int seq = (unsigned int)xseq;
            double *mywin = vb.xwin[0];            // lane r serves rank r (static selects: no dynamic index into the parameter block)
#pragma unroll
            for (int r = 1; r < BPK_XCHG_MAXRANKS; ++r)
                if (lane == r) mywin = vb.xwin[r];
            const int xr = vb.xranks;
            for (int ee = w; ee < per; ee += 2 * WS_PAIRS) {
                const int e = e0 + ee;
                if (e >= PCA_NSTAT) continue;
                if (DERIVE_SXX && e >= PCA_MP * PCA_KP && e < PCA_MP * PCA_KP + PCA_KP * PCA_KP) continue;   // S_xx is formed in the tail
                double s = 0.0;
                for (int bb = lane; bb < (int)gridDim.x; bb += 32) s += __ldcg(partial + (size_t)bb * PCA_NSTAT + e);
                s = warp_sum(s);
                if (xr > 1) {
                    if (lane < xr) ll_store(ll_slot(mywin, par, vb.xrank, e), s, seq);          // my share -> every rank
                    double v = 0.0;
                    if (lane < xr) v = ll_wait(ll_slot(vb.xown, par, lane, e), seq, vb.ctrl + 2);   // every rank's share -> me
                    s = 0.0;
#pragma unroll
                    for (int r = 0; r < BPK_XCHG_MAXRANKS; ++r) {                                // rank order: same bits everywhere
                        const double vr = __shfl_sync(0xffffffffu, v, r);
                        if (r < xr) s += vr;
                    }
                }
                if (lane == 0) ll_store(ll_total_slot(vb.xown, par, e), s, seq);
            }
            if (pr) vb_stamp(vb.dbg, 55);
            if (blockIdx.x == 0) {
                vb_stamp(vb.dbg, 3);
                vb_stamp(vb.dbg, 4);
                pca_vb_ops(vb, smem, vb_sm_doubles, false, nops_it, xseq);
                vb_stamp(vb.dbg, 5);
            }
        } else {
            double *fin = partial + (size_t)gridDim.x * PCA_NSTAT;
            for (int ee = w; ee < per; ee += 2 * WS_PAIRS) {
                const int e = e0 + ee;
                if (e < PCA_NSTAT) {
                    double s = 0.0;
                    for (int bb = lane; bb < (int)gridDim.x; bb += 32) s += __ldcg(partial + (size_t)bb * PCA_NSTAT + e);
                    s = warp_sum(s);
                    if (lane == 0) fin[e] = s;
                }
            }
            if (blockIdx.x == 0) vb_stamp(vb.dbg, 3);
            grid_barrier(gbar, epoch, vb.ctrl + 2);
            if (blockIdx.x == 0) {
                vb_stamp(vb.dbg, 4);
                pca_vb_ops(vb, smem, vb_sm_doubles, false, nops_it);
                vb_stamp(vb.dbg, 5);
            }
        }
        if (it + 1 < niter) grid_barrier(gbar, epoch, vb.ctrl + 2);      // the next sweep's A, b (and the stop word) are visible to every CTA
        if (pr) vb_stamp(vb.dbg, 56);
    }
    }   // sweeps of this launch
}

// =====================================================================================
// v3: the persistent VB loop with a SERVICE CTA and DYNAMIC tile scheduling (what the resident loop launches when the
// shard is large enough).  Measured on one steady-state sweep of the v2 loop at 1.25 M columns per GPU (the per-rank size
// of the 8-GPU strong-scaling run; profiles/r02_probe_*): 10 us until the A fragments had arrived (148 x 8 warps
// fetching the same 8 KB from L2), 12-18 us between the first and the last CTA reaching the grid barrier (equal tile
// counts, unequal SM speed), 74 us of tail + barrier on CTA 0 (state staged in and out of shared memory every sweep,
// code re-warmed by a dry run) — 287 us per sweep of which 184 us stream data.  Here:
//   * CTA 0 does no data pass at all: it keeps the state vector resident in shared memory for the whole launch, its
//     instruction cache holds nothing but the small ops, and it only publishes A, b for the next sweep;
//   * the other CTAs take their tiles from a global counter (the producer warp grabs the next tile when it issues the
//     TMA load and hands the tile index to the consumers through shared memory; a sentinel ends the sweep), so the grid
//     reaches the barrier within about one tile time.  X does not depend on the assignment; the per-CTA partial sums do
//     (fixed reduction order, but which tiles a CTA summed varies): statistics agree to rounding between runs, and
//     every rank still ends up with bit-identical totals after the exchange.  BPK_PCA_STATIC=1 keeps the v2 loop.
//   * A and b enter shared memory once per CTA and sweep (coalesced), after the first TMA loads have been issued.
// =====================================================================================
#define VL_LDA 65                        // pitch of the A tile in shared memory

// distributed, fixed-order reduction of the per-CTA partials + LL push/gather of the owner's slice (see the v2 kernel)
template <bool DERIVE_SXX>
__device__ __forceinline__ void vl_reduce_push(const PcaVbArgs &vb, const double *partial, int w, int lane) {
    // called by the worker CTAs only (blocks 1..): the service CTA takes no slice and goes straight to collecting the
    // totals, so that its wait for the peers' packets overlaps the workers' instead of preceding it
    const int nb = (int)gridDim.x - 1;
    const int per = (PCA_NSTAT + nb - 1) / nb;
    const int e0 = ((int)blockIdx.x - 1) * per;
    const unsigned long long xseq = *(volatile const unsigned long long *)vb.xown + 1ull;
    const int par = (int)(xseq & 1ull);
    const unsigned int seq = (unsigned int)xseq;
    double *mywin = vb.xwin[0];
#pragma unroll
    for (int r = 1; r < BPK_XCHG_MAXRANKS; ++r)
        if (lane == r) mywin = vb.xwin[r];
    const int xr = vb.xranks;
    for (int ee = w; ee < per; ee += 2 * WS_PAIRS) {
        const int e = e0 + ee;
        if (e >= PCA_NSTAT) continue;
        if (DERIVE_SXX && e >= PCA_MP * PCA_KP && e < PCA_MP * PCA_KP + PCA_KP * PCA_KP) continue;
        double s = 0.0;
        for (int bb = 1 + lane; bb < (int)gridDim.x; bb += 32) s += __ldcg(partial + (size_t)bb * PCA_NSTAT + e);   // row 0 (service CTA) is empty
        s = warp_sum(s);
        if (xr > 1) {
            if (lane < xr) ll_store(ll_slot(mywin, par, vb.xrank, e), s, seq);
            double v = 0.0;
            if (lane < xr) v = ll_wait(ll_slot(vb.xown, par, lane, e), seq, vb.ctrl + 2);
            s = 0.0;
#pragma unroll
            for (int r = 0; r < BPK_XCHG_MAXRANKS; ++r) {
                const double vr = __shfl_sync(0xffffffffu, v, r);
                if (r < xr) s += vr;
            }
        }
        if (lane == 0) ll_store(ll_total_slot(vb.xown, par, e), s, seq);
    }
}

template <int NT, int STAGES, int DIST, bool DERIVE_SXX>
__global__ void __launch_bounds__(2 * WS_PAIRS * 32, 1)
pca_vbloop_kernel(const __grid_constant__ CUtensorMap tmapY, int64_t M, int64_t N, int K,
                  const double *A, const double *bvec, double *__restrict__ X, double *partial, int64_t ntiles,
                  const int *stop, unsigned int *gbar, unsigned long long *tctr, const __grid_constant__ PcaVbArgs vb,
                  size_t vb_sm_doubles) {
    constexpr int T = WS_PAIRS * NT, NS = NT / 4, CB = NT / 8, NBOX = T / WS_BOXC, STG = NBOX * WS_BOX;
    static_assert(DIST >= 1 && DIST < STAGES, "prefetch distance must leave one stage for the consumers");
    static_assert(CB == 1, "the loop kernel is written for 8 columns per X-warp");
    if (*stop) return;
    extern __shared__ __align__(1024) double smem_ws[];
    double *smem = smem_ws;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gr = lane >> 2, tg = lane & 3;
    const int niter = vb.niter;
    unsigned int epoch = 0;
    if (threadIdx.x == 0) epoch = *(volatile unsigned int *)&gbar[1];

    // ------------------------------------------------------------------------------------------------------------
    if (blockIdx.x == 0) {
        // SERVICE CTA: grid reduction share, the sweep's small ops on the shared-memory resident state, nothing else
        vb_stamp(vb.dbg, 0);
        for (int e = threadIdx.x; e < PCA_NSTAT; e += blockDim.x) partial[e] = 0.0;       // it sums no tiles
        pca_vb_state_load(vb, smem);
        // The small ops run once per sweep; in between this SM idles at the grid barrier for the whole data pass and the
        // ops' code (transcendentals, two K x K inversions, the bound) falls out of the instruction caches: measured
        // 78 us per tail cold against ~50 us right after a run.  So the ops are DRY-RUN (scratch copy of the state, no
        // side effects) while the grid streams data, timed from the previous sweep's data-pass duration to finish just
        // before the grid arrives.
        __shared__ unsigned long long svc_t[4];       // [0] sweep start, [1] data-pass duration of the previous sweep, [2] dry-run duration
        if (threadIdx.x == 0) {
            unsigned long long t0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            svc_t[0] = t0; svc_t[1] = 0ull; svc_t[2] = 0ull;
        }
        __syncthreads();
        for (int it = 0; it < niter; ++it) {
            if (it > 0 && *(volatile const int *)stop) break;
            const int nops_it = (it == niter - 1) ? vb.nops_last : vb.nops;
            if (vb.dry_every || it == 0) {
                if (threadIdx.x == 0 && svc_t[1] > 0ull) {
                    // start so that the run ends ~5 us before the expected arrival of the grid
                    const unsigned long long lead = svc_t[2] + (svc_t[2] >> 2) + 5000ull;
                    if (svc_t[1] > lead) {
                        const unsigned long long until = svc_t[0] + svc_t[1] - lead;
                        unsigned long long now;
                        for (;;) {
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                            if (now >= until || *(volatile unsigned int *)&gbar[0] > 0u) break;   // or somebody is already there
                            __nanosleep(256);
                        }
                    }
                }
                __syncthreads();
                unsigned long long d0 = 0ull;
                if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(d0));
                pca_vb_ops(vb, smem, vb_sm_doubles, true, nops_it, 0ull, 1);
                __syncthreads();
                if (threadIdx.x == 0) {
                    unsigned long long d1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(d1));
                    svc_t[2] = d1 - d0;
                }
            }
            vb_stamp(vb.dbg, 1);
            grid_arrive(gbar, epoch);
            grid_wait(gbar, epoch, vb.ctrl + 2);
            if (threadIdx.x == 0) {
                unsigned long long tb;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tb));
                svc_t[1] = tb - svc_t[0];
            }
            vb_stamp(vb.dbg, 2);
            if (it < 128) vb_stamp(vb.dbg, 64 + it);
            vb_stamp(vb.dbg, 3);
            const unsigned long long xseq = *(volatile const unsigned long long *)vb.xown + 1ull;
            vb_stamp(vb.dbg, 4);
            pca_vb_ops(vb, smem, vb_sm_doubles, false, nops_it, xseq, 1);
            vb_stamp(vb.dbg, 5);
            if (threadIdx.x == 0) *(volatile unsigned long long *)tctr = 0ull;    // next sweep's tile counter
            if (it + 1 < niter) grid_barrier(gbar, epoch, vb.ctrl + 2);
            if (threadIdx.x == 0) {
                unsigned long long ts;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
                svc_t[0] = ts;
            }
        }
        pca_vb_state_store(vb, smem);
        return;
    }

    // ------------------------------------------------------------------------------------------------------------
    // WORKER CTAs
    double *Ysm = smem;                                              // [STAGES][NBOX][64][16] swizzled
    double *Xsm = Ysm + (size_t)STAGES * STG;                        // [PAIRS][2][NT][LDX]
    double *Asm = Xsm + (size_t)WS_PAIRS * 2 * NT * PCA_LDX;         // [16][VL_LDA] + b[16]
    uint64_t *bars = (uint64_t *)(Asm + PCA_KP * VL_LDA + PCA_KP);
    uint64_t *full = bars, *empty = bars + STAGES;
    uint64_t *xfull = bars + 2 * STAGES, *xfree = xfull + 2 * WS_PAIRS;   // [pair][2]
    long long *stile = (long long *)(xfree + 2 * WS_PAIRS);          // [STAGES] tile index of the stage, -1 = end of sweep
    const bool is_x = w < WS_PAIRS;
    const int p = w & (WS_PAIRS - 1);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2 * WS_PAIRS); }
        for (int i = 0; i < 2 * WS_PAIRS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xfree[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tmapY) : "memory");
    }
    __syncthreads();
    const bool probe = vb.dbg != nullptr && blockIdx.x == gridDim.x - 1;

    int ig = 0;                  // stages consumed by this CTA since the launch began (every warp counts the same)
    int xg = 0;                  // X -> S hand-offs (stages that carried a tile)
    int issued = 0;              // producer (warp 0): stages issued
    // producer: take the next tile from the global counter and start its load (or post the end-of-sweep sentinel)
    // The grab is software-pipelined: the counter is read one issue ahead (`pending`), so the L2 round trip of the atomic
    // overlaps the tile in between instead of sitting on X-warp 0's critical path; no tile is lost, because a producer
    // processes every valid index it ever drew (after the first index >= ntiles every later one is >= ntiles too).
    bool ended = false, have_pending = false;
    long long pending = 0;
    auto issue = [&]() {
        const int pos = issued;
        const int slot = pos % STAGES;
        long long tile = 0;
        if (lane == 0) {
            tile = have_pending ? pending : (long long)atomicAdd(tctr, 1ull);
            if (tile < ntiles) pending = (long long)atomicAdd(tctr, 1ull);     // needed at the next issue only
        }
        have_pending = true;
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (pos >= STAGES) mbar_wait(&empty[slot], (uint32_t)((pos / STAGES - 1) & 1));
        if (tile < ntiles) {
            if (lane == 0) {
                stile[slot] = tile;
                mbar_expect_tx(&full[slot], (uint32_t)(STG * sizeof(double)));
            }
            __syncwarp();
            if (lane < NBOX)
                tma_load_2d(Ysm + (size_t)slot * STG + lane * WS_BOX, &tmapY, (int)(tile * T + lane * WS_BOXC), 0, &full[slot]);
        } else {
            if (lane == 0) {
                stile[slot] = -1;
                mbar_arrive(&full[slot]);
            }
            __syncwarp();
            ended = true;
        }
        issued = pos + 1;
    };

    for (int it = 0; it < niter; ++it) {
        const bool pr = probe && (it == niter - 2 || niter == 1);
        if (probe && it == niter - 1 && niter > 1) vb_stamp(vb.dbg, 57);
        if (pr) vb_stamp(vb.dbg, 48);
        if (it > 0) {
            if (*(volatile const int *)stop) break;                       // converged inside this launch (uniform)
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        }
        ended = false;
        have_pending = false;                                             // the counter was reset for this sweep
        if (w == 0)
            while (!ended && issued < ig + DIST) issue();                 // loads first ...
        // ... then this sweep's A, b -> shared memory (coalesced; one L2 fetch per CTA instead of one per warp)
        for (int e = threadIdx.x; e < PCA_KP * PCA_MP; e += blockDim.x) {
            const int k = e / PCA_MP, m = e - k * PCA_MP;
            Asm[k * VL_LDA + m] = (k < K && m < M) ? __ldcg(A + (int64_t)k * M + m) : 0.0;
        }
        if (threadIdx.x < PCA_KP) Asm[PCA_KP * VL_LDA + threadIdx.x] = (bvec && (int)threadIdx.x < K) ? __ldcg(bvec + threadIdx.x) : 0.0;
        __syncthreads();
        if (is_x) {
            double afrag[2][16];
            double bk[2];
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int ms = 0; ms < 16; ++ms)
                    afrag[kb][ms] = Asm[(kb * 8 + gr) * VL_LDA + (ms >> 1) * 8 + 2 * tg + (ms & 1)];
            bk[0] = Asm[PCA_KP * VL_LDA + gr];
            bk[1] = Asm[PCA_KP * VL_LDA + 8 + gr];
            if (pr && w == 0) vb_stamp_lane(vb.dbg, 49);
            bool first = true;
            for (;;) {
                if (w == 0)
                    while (!ended && issued < ig + DIST + 1) issue();
                const int slot = (int)(ig % STAGES);
                mbar_wait(&full[slot], (uint32_t)((ig / STAGES) & 1));
                const long long tile = *(volatile long long *)&stile[slot];
                if (tile < 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[slot]);
                    ++ig;
                    break;
                }
                if (pr && w == 0 && first) vb_stamp_lane(vb.dbg, 50);
                first = false;
                const int b = (int)(xg & 1);
                const int c0 = p * NT;
                const int64_t nbase = tile * T + c0;
                double *Xs = Xsm + ((size_t)(p * 2 + b) * NT) * PCA_LDX;
                const double *Ys = Ysm + (size_t)slot * STG;
                double acc[2][2][2];
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    acc[0][kb][0] = acc[0][kb][1] = bk[kb];
                    acc[1][kb][0] = acc[1][kb][1] = 0.0;
                }
#pragma unroll
                for (int ms = 0; ms < 16; ++ms) {
                    const double yb = Ys[ws_off((ms >> 1) * 8 + 2 * tg + (ms & 1), c0 + gr)];
                    dmma884(acc[ms & 1][0][0], acc[ms & 1][0][1], afrag[0][ms], yb);
                    dmma884(acc[ms & 1][1][0], acc[ms & 1][1][1], afrag[1][ms], yb);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);           // this warp no longer reads the Y stage
                if (xg >= 2) mbar_wait(&xfree[p * 2 + b], (uint32_t)(((xg >> 1) - 1) & 1));
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int nl = 2 * tg + j, k = kb * 8 + gr;
                        const bool inb = nbase + nl < N;
                        const double v = inb ? acc[0][kb][j] + acc[1][kb][j] : 0.0;   // padded columns add nothing
                        Xs[nl * PCA_LDX + k] = v;
                        if (inb && k < K) X[(nbase + nl) * K + k] = v;
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(&xfull[p * 2 + b]);
                ++ig;
                ++xg;
            }
            if (pr && w == 0) vb_stamp_lane(vb.dbg, 51);
        } else {
            const int rr = 2 * (gr & 3) + (gr >> 2);      // S_yx row within an 8-row block held by this lane group
            double syx[8][2][2];
            double sxx[3][2];
            double sx = 0.0;
#pragma unroll
            for (int mb = 0; mb < 8; ++mb)
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) syx[mb][kb][0] = syx[mb][kb][1] = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) sxx[i][0] = sxx[i][1] = 0.0;
            for (;;) {
                const int slot = (int)(ig % STAGES);
                mbar_wait(&full[slot], (uint32_t)((ig / STAGES) & 1));
                const long long tile = *(volatile long long *)&stile[slot];
                if (tile < 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[slot]);
                    ++ig;
                    break;
                }
                const int b = (int)(xg & 1);
                const int c0 = p * NT;
                const double *Ys = Ysm + (size_t)slot * STG;
                const double *Xs = Xsm + ((size_t)(p * 2 + b) * NT) * PCA_LDX;
                mbar_wait(&xfull[p * 2 + b], (uint32_t)((xg >> 1) & 1));
                if (lane < PCA_KP) {
#pragma unroll
                    for (int nl = 0; nl < NT; ++nl) sx += Xs[nl * PCA_LDX + lane];
                }
#pragma unroll
                for (int ns = 0; ns < NS; ++ns) {
                    const double xf0 = Xs[(ns * 4 + tg) * PCA_LDX + gr];
                    const double xf1 = Xs[(ns * 4 + tg) * PCA_LDX + 8 + gr];
#pragma unroll
                    for (int mb = 0; mb < 8; ++mb) {
                        const double ya = Ys[ws_off(mb * 8 + rr, c0 + ns * 4 + tg)];
                        dmma884(syx[mb][0][0], syx[mb][0][1], ya, xf0);
                        dmma884(syx[mb][1][0], syx[mb][1][1], ya, xf1);
                    }
                    if (!DERIVE_SXX) {
                        dmma884(sxx[0][0], sxx[0][1], xf0, xf0);
                        dmma884(sxx[1][0], sxx[1][1], xf0, xf1);
                        dmma884(sxx[2][0], sxx[2][1], xf1, xf1);
                    }
                }
                __syncwarp();
                if (lane == 0) { mbar_arrive(&xfree[p * 2 + b]); mbar_arrive(&empty[slot]); }
                ++ig;
                ++xg;
            }
            if (pr && w == WS_PAIRS) vb_stamp_lane(vb.dbg, 52);
            // every stage has been consumed by every warp once all warps pass the barrier below
            __syncthreads();
            double *mine = smem + (size_t)p * PCA_NSTAT;
#pragma unroll
            for (int mb = 0; mb < 8; ++mb)
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        mine[(mb * 8 + rr) * PCA_KP + kb * 8 + 2 * tg + j] = syx[mb][kb][j];
            double *mxx = mine + PCA_MP * PCA_KP;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                mxx[gr * PCA_KP + 2 * tg + j] = sxx[0][j];
                mxx[gr * PCA_KP + 8 + 2 * tg + j] = sxx[1][j];
                mxx[(8 + 2 * tg + j) * PCA_KP + gr] = sxx[1][j];
                mxx[(8 + gr) * PCA_KP + 8 + 2 * tg + j] = sxx[2][j];
            }
            if (lane < PCA_KP) mine[PCA_MP * PCA_KP + PCA_KP * PCA_KP + lane] = sx;
        }
        if (is_x) __syncthreads();     // pairs with the S-warps' barrier above
        __syncthreads();
        double *pout = partial + (size_t)blockIdx.x * PCA_NSTAT;
        for (int e = threadIdx.x; e < PCA_NSTAT; e += blockDim.x) {
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < WS_PAIRS; ++ww) s += smem[(size_t)ww * PCA_NSTAT + e];
            pout[e] = s;
        }
        if (pr) vb_stamp(vb.dbg, 53);
        grid_arrive(gbar, epoch);
        grid_wait(gbar, epoch, vb.ctrl + 2);
        if (pr) vb_stamp(vb.dbg, 54);
        vl_reduce_push<DERIVE_SXX>(vb, partial, w, lane);
        if (pr) vb_stamp(vb.dbg, 55);
        if (it + 1 < niter) grid_barrier(gbar, epoch, vb.ctrl + 2);      // the next sweep's A, b, the stop word and the tile counter
        if (pr) vb_stamp(vb.dbg, 56);
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
static unsigned int *g_pca_gbar = nullptr;     // grid barrier of the fused sweep (count, epoch)
static unsigned long long *g_pca_tctr = nullptr;   // tile counter of the persistent loop kernel (dynamic scheduling)

typedef CUresult (*bpk_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                         const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                         CUtensorMapFloatOOBfill);
static int pca_make_tmap(const double *Y, int64_t M, int64_t N, CUtensorMap *out) {
    static bpk_encode_tiled_fn enc = nullptr;
    if (!enc) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess)
            return bpk_set_error(BPK_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
        enc = (bpk_encode_tiled_fn)fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)N * sizeof(double)};
    cuuint32_t box[2] = {WS_BOXC, PCA_MP};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)Y, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return bpk_set_error(BPK_ECUDA, "cuTensorMapEncodeTiled failed (%d) for Y[%lld][%lld]", (int)r,
                                                (long long)M, (long long)N);
    return BPK_OK;
}

template <int NT, int STAGES, int DIST, bool COMPUTE_X>
static int pca_launch_ws(const double *Y, int64_t M, int64_t N, int K, const double *A, const double *b,
                         double *X, const int *stop, double **partial_out, int *nparts_out,
                         const PcaVbArgs *tail = nullptr) {
    constexpr int T = WS_PAIRS * NT;
    size_t ring = (size_t)STAGES * (T / WS_BOXC) * WS_BOX * sizeof(double);
    size_t xs = (size_t)WS_PAIRS * 2 * NT * PCA_LDX * sizeof(double);
    size_t bars = (size_t)(2 * STAGES + 4 * WS_PAIRS) * sizeof(uint64_t);
    size_t redb = (size_t)WS_PAIRS * PCA_NSTAT * sizeof(double);
    size_t smem = ring + xs + bars;
    if (smem < redb) smem = redb;
    int64_t ntiles = (N + T - 1) / T;
    int grid = g_bpk.sm_count;
    if (ntiles < grid) grid = (int)ntiles;
    double *partial = bpk_scratch((size_t)(grid + 1) * PCA_NSTAT * sizeof(double));
    if (!partial) return bpk_set_error(BPK_ECUDA, "pca: scratch allocation failed");
    // the descriptor depends on (Y, M, N) only: resident runs launch on the same data over and over
    static CUtensorMap tmap;
    static const double *tmap_Y = nullptr;
    static int64_t tmap_M = -1, tmap_N = -1;
    if (tmap_Y != Y || tmap_M != M || tmap_N != N) {
        int rc = pca_make_tmap(Y, M, N, &tmap);
        if (rc) { tmap_Y = nullptr; return rc; }
        tmap_Y = Y; tmap_M = M; tmap_N = N;
    }
    if (tail && COMPUTE_X) {
        // one launch = data pass + grid reduction + the sweep's small ops
        if (!g_pca_gbar) {
            BPK_CUDA(cudaMalloc(&g_pca_gbar, 2 * sizeof(unsigned int)));
            BPK_CUDA(cudaMemsetAsync(g_pca_gbar, 0, 2 * sizeof(unsigned int), g_bpk.stream));
        }
        // the arrival count is 0 between launches; re-zeroing it heals the barrier after a watchdog exit
        BPK_CUDA(cudaMemsetAsync(g_pca_gbar, 0, sizeof(unsigned int), g_bpk.stream));
        PcaVbArgs vb = *tail;
        vb.partial = partial + (size_t)grid * PCA_NSTAT;     // the grid-reduced statistics
        vb.nparts = 1;
        if constexpr (NT == 8) {
            // the persistent loop with a service CTA and dynamic tile scheduling: large shards, exchange through the windows
            if (vb.derive_sxx && vb.ll && grid >= 8 && ntiles >= 8 * (int64_t)grid && !getenv("BPK_PCA_STATIC")) {
                if (!g_pca_tctr) BPK_CUDA(cudaMalloc(&g_pca_tctr, sizeof(unsigned long long)));
                BPK_CUDA(cudaMemsetAsync(g_pca_tctr, 0, sizeof(unsigned long long), g_bpk.stream));
                size_t smem3 = ring + xs + (size_t)(PCA_KP * VL_LDA + PCA_KP) * sizeof(double) +
                               (size_t)(2 * STAGES + 4 * WS_PAIRS) * sizeof(uint64_t) + (size_t)STAGES * sizeof(long long);
                if (smem3 < redb) smem3 = redb;
                auto kern = pca_vbloop_kernel<NT, STAGES, DIST, true>;
                static size_t attr_set = 0;
                if (attr_set != smem3) {
                    BPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
                    attr_set = smem3;
                }
                BPK_LAUNCH(kern, grid, 2 * WS_PAIRS * 32, smem3, tmap, M, N, K, A, b, X, partial, ntiles, stop, g_pca_gbar,
                           g_pca_tctr, vb, smem3 / sizeof(double));
                *partial_out = partial + (size_t)grid * PCA_NSTAT;
                *nparts_out = 1;
                return BPK_OK;
            }
        }
        if (vb.derive_sxx) {
            auto kern = pca_xsweep_ws_kernel<NT, STAGES, DIST, true, true, true>;
            BPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            BPK_LAUNCH(kern, grid, 2 * WS_PAIRS * 32, smem, tmap, M, N, K, A, b, X, partial, ntiles, stop, g_pca_gbar, vb, smem / sizeof(double));
        } else {
            auto kern = pca_xsweep_ws_kernel<NT, STAGES, DIST, true, true, false>;
            BPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            BPK_LAUNCH(kern, grid, 2 * WS_PAIRS * 32, smem, tmap, M, N, K, A, b, X, partial, ntiles, stop, g_pca_gbar, vb, smem / sizeof(double));
        }
        *partial_out = partial + (size_t)grid * PCA_NSTAT;
        *nparts_out = 1;
        return BPK_OK;
    }
    PcaVbArgs none;
    none.nops = 0; none.nops_last = 0; none.niter = 1; none.dry_every = 0; none.derive_sxx = 0; none.ll = 0; none.gj2 = 0; none.par = 0;
    none.xranks = 1; none.xrank = 0; none.dbg = nullptr; none.xown = nullptr;
    auto kern = pca_xsweep_ws_kernel<NT, STAGES, DIST, COMPUTE_X, false, false>;
    BPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BPK_LAUNCH(kern, grid, 2 * WS_PAIRS * 32, smem, tmap, M, N, K, A, b, X, partial, ntiles, stop, (unsigned int *)nullptr, none, (size_t)0);
    *partial_out = partial;
    *nparts_out = grid;
    return BPK_OK;
}

// grid-level reduction of the per-CTA partials into the caller's (unpadded) stats
__global__ void pca_stats_final_kernel(const double *__restrict__ partial, int nblocks, int M, int K,
                                       double *__restrict__ stats) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    int total = M * K + K * K + K;
    if (e >= total) return;
    int pe;
    if (e < M * K) { int m = e / K, k = e - m * K; pe = m * PCA_KP + k; }
    else if (e < M * K + K * K) { int r = e - M * K; int i = r / K, j = r - i * K; pe = PCA_MP * PCA_KP + i * PCA_KP + j; }
    else { pe = PCA_MP * PCA_KP + PCA_KP * PCA_KP + (e - M * K - K * K); }
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * PCA_NSTAT + pe];
    stats[e] += s;
}

// BPK_PCA_VARIANT (A/B measurements, tools/bench_xsweep.py): -1 = the 8-warp v1 kernel,
// 0..3 = tile/ring shapes of the warp-specialised kernel; default 0.
static int pca_variant() {
    static int v = -2;
    if (v == -2) {
        const char *e = getenv("BPK_PCA_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <int NT, int STAGES, bool COMPUTE_X>
static int pca_launch_partials(const double *Y, int64_t M, int64_t N, int K, const double *A, const double *b,
                               double *X, const int *stop, double **partial_out, int *nparts_out,
                               const PcaVbArgs *tail = nullptr, int *tail_done = nullptr) {
    constexpr int T = PCA_WARPS * NT, LDY = T + 4;
    size_t ring = (size_t)STAGES * PCA_MP * LDY * sizeof(double);
    size_t xs = (size_t)PCA_WARPS * NT * PCA_LDX * sizeof(double);
    size_t redb = (size_t)PCA_WARPS * PCA_NSTAT * sizeof(double);
    size_t smem = ring + xs;
    if (smem < redb) smem = redb;
    int64_t ntiles = (N + T - 1) / T;
    int grid = g_bpk.sm_count;
    if (ntiles < grid) grid = (int)ntiles;
    double *partial = bpk_scratch((size_t)grid * PCA_NSTAT * sizeof(double));
    if (!partial) return bpk_set_error(BPK_ECUDA, "pca: scratch allocation failed");
    bool al = (N % 2 == 0) && (((uintptr_t)Y & 15u) == 0);
    if (al && pca_variant() >= 0) {
        if (tail_done) *tail_done = (tail != nullptr && COMPUTE_X);
        switch (pca_variant()) {
        case 1: return pca_launch_ws<8, 6, 4, COMPUTE_X>(Y, M, N, K, A, b, X, stop, partial_out, nparts_out, tail);
        case 2: return pca_launch_ws<16, 2, 1, COMPUTE_X>(Y, M, N, K, A, b, X, stop, partial_out, nparts_out, tail);
        case 3: return pca_launch_ws<8, 4, 2, COMPUTE_X>(Y, M, N, K, A, b, X, stop, partial_out, nparts_out, tail);
        default: return pca_launch_ws<8, 5, 3, COMPUTE_X>(Y, M, N, K, A, b, X, stop, partial_out, nparts_out, tail);
        }
    }
    if (al) {
        auto kern = pca_xsweep_kernel<NT, STAGES, true, COMPUTE_X>;
        BPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BPK_LAUNCH(kern, grid, PCA_WARPS * 32, smem, Y, M, N, K, A, b, X, partial, ntiles, stop);
    } else {
        auto kern = pca_xsweep_kernel<NT, STAGES, false, COMPUTE_X>;
        BPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BPK_LAUNCH(kern, grid, PCA_WARPS * 32, smem, Y, M, N, K, A, b, X, partial, ntiles, stop);
    }
    *partial_out = partial;
    *nparts_out = grid;
    return BPK_OK;
}

template <int NT, int STAGES, bool COMPUTE_X>
static int pca_launch(const double *Y, int64_t M, int64_t N, int K, const double *A, const double *b,
                      double *X, double *stats) {
    double *partial = nullptr;
    int grid = 0;
    int rc = pca_launch_partials<NT, STAGES, COMPUTE_X>(Y, M, N, K, A, b, X, nullptr, &partial, &grid);
    if (rc) return rc;
    int total = (int)(M * K + K * K + K);
    BPK_LAUNCH(pca_stats_final_kernel, (total + 127) / 128, 128, 0, partial, grid, (int)M, K, stats);
    return BPK_OK;
}

// will bpk_pca_xsweep use the TMA kernel for this input (16-byte aligned rows)?
bool pca_ws_available(const double *Y, int64_t N) {
    return (N % 2 == 0) && (((uintptr_t)Y & 15u) == 0) && pca_variant() >= 0;
}

// resident VB loop (pca_vb.cu): sweep only, per-CTA partials left in scratch
int pca_xsweep_partials(const double *Y, int64_t M, int64_t N, int K, const double *A, const double *b,
                        double *X, const int *stop, double **partial_out, int *nparts_out,
                        const PcaVbArgs *tail, int *tail_done) {
    if (tail_done) *tail_done = 0;
    return pca_launch_partials<16, 2, true>(Y, M, N, K, A, b, X, stop, partial_out, nparts_out, tail, tail_done);
}

// ---- generic shapes (M > 64 or K > 16): plain kernels, not roofline-tuned -----------
__global__ void pca_x_generic_kernel(const double *__restrict__ Y, int64_t M, int64_t N, int K,
                                     const double *__restrict__ A, const double *__restrict__ b,
                                     double *__restrict__ X) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; e < N * K; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = e / K;
        int k = (int)(e - n * K);
        double s = b ? b[k] : 0.0;
        for (int64_t m = 0; m < M; ++m) s += A[(int64_t)k * M + m] * Y[m * N + n];
        X[e] = s;
    }
}

static int pca_stats_generic(const double *Y, int64_t M, int64_t N, int K, const double *X, double *stats) {
    // S_yx[m,k] += sum_n Y[m,n] X[n,k] ; S_xx[i,j] += sum_n X[n,i] X[n,j] ; s_x[k] += sum_n X[n,k]
    const void *in[2];
    int dt[2] = {BPK_F64, BPK_F64};
    {
        int64_t shape[3] = {M, N, K};
        int64_t is[6] = {N, 1, 0, 0, K, 1};
        int64_t os[3] = {K, 0, 1};
        in[0] = Y; in[1] = X;
        int rc = bpk_sum_multiply(3, shape, 2, in, dt, is, stats, os, 1.0, 1);
        if (rc) return rc;
    }
    {
        int64_t shape[3] = {N, K, K};
        int64_t is[6] = {K, 1, 0, K, 0, 1};
        int64_t os[3] = {0, K, 1};
        in[0] = X; in[1] = X;
        int rc = bpk_sum_multiply(3, shape, 2, in, dt, is, stats + M * K, os, 1.0, 1);
        if (rc) return rc;
    }
    {
        int64_t shape[2] = {N, K};
        int64_t is[2] = {K, 1};
        int64_t os[2] = {0, 1};
        in[0] = X;
        int rc = bpk_sum_multiply(2, shape, 1, in, dt, is, stats + M * K + (int64_t)K * K, os, 1.0, 1);
        if (rc) return rc;
    }
    return BPK_OK;
}

extern "C" int bpk_pca_xsweep(const double *Y, int64_t M, int64_t N, int K,
                              const double *A, const double *b, double *X, double *stats) {
    BPK_REQUIRE_INIT();
    if (M < 1 || K < 1 || N < 0) return bpk_set_error(BPK_EINVAL, "bpk_pca_xsweep: bad shape");
    if (N == 0) return BPK_OK;
    if (M <= PCA_MP && K <= PCA_KP) return pca_launch<16, 2, true>(Y, M, N, K, A, b, X, stats);
    int64_t blocks = (N * K + 255) / 256;
    int64_t cap = (int64_t)g_bpk.sm_count * 32;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(pca_x_generic_kernel, (unsigned)blocks, 256, 0, Y, M, N, K, A, b, X);
    return pca_stats_generic(Y, M, N, K, X, stats);
}

extern "C" int bpk_pca_stats(const double *Y, int64_t M, int64_t N, int K, const double *X, double *stats) {
    BPK_REQUIRE_INIT();
    if (M < 1 || K < 1 || N < 0) return bpk_set_error(BPK_EINVAL, "bpk_pca_stats: bad shape");
    if (N == 0) return BPK_OK;
    if (M <= PCA_MP && K <= PCA_KP)
        return pca_launch<16, 2, false>(Y, M, N, K, nullptr, nullptr, const_cast<double *>(X), stats);
    return pca_stats_generic(Y, M, N, K, X, stats);
}

// ---- sum y^2 / count of observed entries (constant of the tau update) ---------------
__global__ void __launch_bounds__(256) sumsq_kernel(const double *__restrict__ Y, const uint8_t *__restrict__ mask,
                                                    int64_t count, double *__restrict__ partial) {
    double s = 0.0, c = 0.0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double y = Y[i];
        if (!mask || mask[i]) { s += y * y; c += 1.0; }
    }
    __shared__ double rs[8], rc[8];
    s = warp_sum(s); c = warp_sum(c);
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { rs[w] = s; rc[w] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < 8; ++k) { a += rs[k]; b += rc[k]; }
        partial[2 * blockIdx.x] = a;
        partial[2 * blockIdx.x + 1] = b;
    }
}
__global__ void sumsq_final_kernel(const double *__restrict__ partial, int nblocks, double *__restrict__ out2) {
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += partial[2 * b + threadIdx.x];
        out2[threadIdx.x] = s;
    }
}
extern "C" int bpk_sumsq(const double *Y, const uint8_t *mask, int64_t count, double *out2) {
    BPK_REQUIRE_INIT();
    int64_t blocks = (count + 255) / 256;
    int64_t cap = (int64_t)g_bpk.sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    double *partial = bpk_scratch((size_t)blocks * 2 * sizeof(double));
    if (!partial) return bpk_set_error(BPK_ECUDA, "sumsq: scratch allocation failed");
    BPK_LAUNCH(sumsq_kernel, (unsigned)blocks, 256, 0, Y, mask, count, partial);
    BPK_LAUNCH(sumsq_final_kernel, 1, 32, 0, partial, (int)blocks, out2);
    return BPK_OK;
}

// ---- masked variant (missing values): per-column precision -------------------------------
// v0: one warp per column builds Lam_n = diag(alpha) + tau sum_m mask[m,n] <w_m w_m^T>
// in shared memory, factors/inverts it with the warp-cooperative routines of linalg.cu,
// and writes x_n, Cov_n, g_n.  The mask-weighted statistics are then two plate-sums over
// n (bpk_sum_multiply).  Next: fuse the statistics into the column kernel and move
// the two M x K^2 contractions onto DMMA (they are GEMMs with the mask as one operand).
#define MLD(D) SPD_LD(D)
#define m_warp_chol_upper spd_warp_chol_upper

__global__ void pca_masked_cols_kernel(const double *__restrict__ Y, const uint8_t *__restrict__ mask,
                                       int64_t M, int64_t N, int K,
                                       const double *__restrict__ W, const double *__restrict__ WW,
                                       double tau, const double *__restrict__ alpha, const double *__restrict__ amu,
                                       double *__restrict__ X, double *__restrict__ COV, double *__restrict__ g,
                                       int *flag) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ld = MLD(K);
    double *S = smem + (size_t)w * (2 * (size_t)K * ld + (size_t)K * 32 + K);
    double *B = S + (size_t)K * ld;
    double *C = B + (size_t)K * 32;
    double *p0 = C + (size_t)K * ld;
    for (int64_t n = (int64_t)blockIdx.x * nw + w; n < N; n += (int64_t)gridDim.x * nw) {
        __syncwarp();
        for (int e = lane; e < K * K; e += 32) {
            int i = e / K, j = e - i * K;
            double s = (i == j) ? alpha[i] : 0.0;
            for (int64_t m = 0; m < M; ++m)
                if (mask[m * N + n]) s += tau * WW[m * K * K + e];
            S[i * ld + j] = s;
        }
        for (int k = lane; k < K; k += 32) {
            double s = amu ? amu[k] : 0.0;
            for (int64_t m = 0; m < M; ++m)
                if (mask[m * N + n]) s += tau * Y[m * N + n] * W[m * K + k];
            p0[k] = s;
        }
        __syncwarp();
        int bad = m_warp_chol_upper(S, K, ld, lane);
        if (bad && lane == 0) atomicOr(flag, BPK_FLAG_NOTSPD);
        double ldt = 0.0;
        for (int i = lane; i < K; i += 32) ldt += log(S[i * ld + i]);
        ldt = 2.0 * warp_sum(ldt);
        // inverse: lane-per-column solves against the identity
        for (int c0 = 0; c0 < K; c0 += 32) {
            int nc = K - c0 < 32 ? K - c0 : 32;
            __syncwarp();
            for (int e = lane; e < K * nc; e += 32) {
                int i = e / nc, c = e % nc;
                B[i * 32 + c] = (i == c0 + c) ? 1.0 : 0.0;
            }
            __syncwarp();
            if (lane < nc) {
                for (int i = 0; i < K; ++i) {
                    double s = B[i * 32 + lane];
                    for (int j = 0; j < i; ++j) s -= S[j * ld + i] * B[j * 32 + lane];
                    B[i * 32 + lane] = s / S[i * ld + i];
                }
                for (int i = K - 1; i >= 0; --i) {
                    double s = B[i * 32 + lane];
                    for (int j = i + 1; j < K; ++j) s -= S[i * ld + j] * B[j * 32 + lane];
                    B[i * 32 + lane] = s / S[i * ld + i];
                }
            }
            __syncwarp();
            for (int e = lane; e < K * nc; e += 32) {
                int i = e / nc, c = e % nc;
                C[i * ld + c0 + c] = B[i * 32 + c];
            }
        }
        __syncwarp();
        double dot = 0.0;
        for (int i = lane; i < K; i += 32) {
            double s = 0.0;
            for (int j = 0; j < K; ++j) s += C[i * ld + j] * p0[j];
            X[n * K + i] = s;
            dot += s * p0[i];
        }
        dot = warp_sum(dot);
        if (g && lane == 0) g[n] = -0.5 * dot + 0.5 * ldt;
        if (COV)
            for (int e = lane; e < K * K; e += 32) COV[n * K * K + e] = C[(e / K) * ld + (e % K)];
    }
}

extern "C" int bpk_pca_xsweep_masked(const double *Y, const uint8_t *mask, int64_t M, int64_t N, int K,
                                     const double *W, const double *WW, double tau,
                                     const double *alpha, const double *amu,
                                     double *X, double *COV, double *g, double *stats, int check) {
    BPK_REQUIRE_INIT();
    if (M < 1 || K < 1 || K > BPK_MAXDIM || N < 0) return bpk_set_error(BPK_EINVAL, "bpk_pca_xsweep_masked: bad shape");
    if (N == 0) return BPK_OK;
    const int ld = MLD(K);
    size_t per = (2 * (size_t)K * ld + (size_t)K * 32 + K) * sizeof(double);
    int nw = (int)((160u << 10) / per);
    if (nw > 8) nw = 8;
    if (nw < 1) nw = 1;
    size_t smem = nw * per;
    if (smem > (48u << 10))
        BPK_CUDA(cudaFuncSetAttribute(pca_masked_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double *cov = COV;
    bool own_cov = false;
    if (!cov) {
        BPK_CUDA(cudaMallocAsync((void **)&cov, (size_t)N * K * K * sizeof(double), g_bpk.stream));
        own_cov = true;
    }
    int64_t blocks = (N + nw - 1) / nw;
    int64_t cap = (int64_t)g_bpk.sm_count * 8;
    if (blocks > cap) blocks = cap;
    BPK_LAUNCH(pca_masked_cols_kernel, (unsigned)blocks, nw * 32, smem, Y, mask, M, N, K, W, WW, tau, alpha, amu,
               X, cov, g, g_bpk.d_flag);
    // statistics: S_yx[m,k] += sum_n mask y x ; S_xx[m,i,j] += sum_n mask (cov + x x^T)
    int rc;
    {
        const void *in[3] = {mask, Y, X};
        int dt[3] = {BPK_U8, BPK_F64, BPK_F64};
        int64_t shape[3] = {M, N, K};
        int64_t is[9] = {N, 1, 0, N, 1, 0, 0, K, 1};
        int64_t os[3] = {K, 0, 1};
        rc = bpk_sum_multiply(3, shape, 3, in, dt, is, stats, os, 1.0, 1);
        if (rc) return rc;
    }
    {
        const void *in[2] = {mask, cov};
        int dt[2] = {BPK_U8, BPK_F64};
        int64_t shape[3] = {M, N, (int64_t)K * K};
        int64_t is[6] = {N, 1, 0, 0, (int64_t)K * K, 1};
        int64_t os[3] = {(int64_t)K * K, 0, 1};
        rc = bpk_sum_multiply(3, shape, 2, in, dt, is, stats + M * K, os, 1.0, 1);
        if (rc) return rc;
    }
    {
        const void *in[3] = {mask, X, X};
        int dt[3] = {BPK_U8, BPK_F64, BPK_F64};
        int64_t shape[4] = {M, N, K, K};
        int64_t is[12] = {N, 1, 0, 0, 0, K, 1, 0, 0, K, 0, 1};
        int64_t os[4] = {(int64_t)K * K, 0, K, 1};
        rc = bpk_sum_multiply(4, shape, 3, in, dt, is, stats + M * K, os, 1.0, 1);
        if (rc) return rc;
    }
    if (own_cov) BPK_CUDA(cudaFreeAsync(cov, g_bpk.stream));
    if (check) return bpk_check_flag(BPK_ENOTSPD);
    return BPK_OK;
}
