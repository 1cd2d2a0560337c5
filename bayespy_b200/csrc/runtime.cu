// runtime.cu — context, memory, timers and the NCCL communicator of libbpk.
// One process drives one GPU; every call is ordered on one compute stream.
#include "common.cuh"
#include <time.h>
#include <stdarg.h>
#include <string.h>
#include <dlfcn.h>
#include <vector>

BpkCtx g_bpk;

int bpk_set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_bpk.err, sizeof(g_bpk.err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *bpk_last_error(void) { return g_bpk.err; }

extern "C" int bpk_init(int device) {
    if (g_bpk.ready) {
        if (device == g_bpk.device) return BPK_OK;
        return bpk_set_error(BPK_EINVAL, "bpk already bound to device %d", g_bpk.device);
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return bpk_set_error(BPK_ENOGPU, "no CUDA device available (%s); libbpk has no CPU path",
                             e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n)
        return bpk_set_error(BPK_EINVAL, "device %d out of range (have %d)", device, n);
    BPK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    BPK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return bpk_set_error(BPK_ENOGPU, "device %s is sm_%d%d; libbpk is built for sm_100a only",
                             prop.name, prop.major, prop.minor);
    g_bpk.device = device;
    g_bpk.sm_count = prop.multiProcessorCount;
    BPK_CUDA(cudaStreamCreateWithFlags(&g_bpk.stream, cudaStreamNonBlocking));
    BPK_CUDA(cudaDeviceGetDefaultMemPool(&g_bpk.pool, device));
    uint64_t thresh = UINT64_MAX;   // keep freed blocks cached: the sweep re-allocates the same sizes
    BPK_CUDA(cudaMemPoolSetAttribute(g_bpk.pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    BPK_CUDA(cudaMalloc(&g_bpk.d_flag, sizeof(int)));
    BPK_CUDA(cudaMemset(g_bpk.d_flag, 0, sizeof(int)));
    BPK_CUDA(cudaMallocHost(&g_bpk.h_flag, sizeof(int)));
    g_bpk.scratch_bytes = 8u << 20;
    BPK_CUDA(cudaMalloc(&g_bpk.scratch, g_bpk.scratch_bytes));
    g_bpk.launches = 0;
    g_bpk.ready = true;
    return BPK_OK;
}

extern "C" int bpk_shutdown(void) {
    if (!g_bpk.ready) return BPK_OK;
    cudaStreamSynchronize(g_bpk.stream);
    if (g_bpk.l2buf) cudaFree(g_bpk.l2buf);
    cudaFree(g_bpk.scratch);
    cudaFree(g_bpk.d_flag);
    cudaFreeHost(g_bpk.h_flag);
    cudaStreamDestroy(g_bpk.stream);
    g_bpk = BpkCtx();
    return BPK_OK;
}

double *bpk_scratch(size_t bytes) {
    if (bytes > g_bpk.scratch_bytes) {
        // stream-ordered replacement: earlier kernels may still read the old block
        cudaFreeAsync(g_bpk.scratch, g_bpk.stream);
        size_t nb = bytes + (bytes >> 1);
        if (cudaMallocAsync((void **)&g_bpk.scratch, nb, g_bpk.stream) != cudaSuccess) {
            g_bpk.scratch = nullptr;
            g_bpk.scratch_bytes = 0;
            return nullptr;
        }
        g_bpk.scratch_bytes = nb;
    }
    return g_bpk.scratch;
}

int bpk_check_flag(int code_if_set) {
    BPK_CUDA(cudaMemcpyAsync(g_bpk.h_flag, g_bpk.d_flag, sizeof(int), cudaMemcpyDeviceToHost, g_bpk.stream));
    BPK_CUDA(cudaMemsetAsync(g_bpk.d_flag, 0, sizeof(int), g_bpk.stream));
    BPK_CUDA(cudaStreamSynchronize(g_bpk.stream));
    int f = *g_bpk.h_flag;
    if (f & BPK_FLAG_NOTSPD) return bpk_set_error(BPK_ENOTSPD, "Matrix not positive definite");
    if (f & BPK_FLAG_DOMAIN) return bpk_set_error(BPK_EDOMAIN, "Natural parameters should be positive");
    (void)code_if_set;
    return BPK_OK;
}

extern "C" int bpk_device_info(int *sm_count, int *cc_major, int *cc_minor,
                               uint64_t *hbm_total, uint64_t *hbm_free) {
    BPK_REQUIRE_INIT();
    cudaDeviceProp prop;
    BPK_CUDA(cudaGetDeviceProperties(&prop, g_bpk.device));
    size_t fr = 0, tot = 0;
    BPK_CUDA(cudaMemGetInfo(&fr, &tot));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (hbm_total) *hbm_total = tot;
    if (hbm_free) *hbm_free = fr;
    return BPK_OK;
}

extern "C" int bpk_device_pci_bus_id(char *buf, int len) {
    BPK_REQUIRE_INIT();
    if (!buf || len < 13) return bpk_set_error(BPK_EINVAL, "bpk_device_pci_bus_id: buffer too small");
    BPK_CUDA(cudaDeviceGetPCIBusId(buf, len, g_bpk.device));
    return BPK_OK;
}

extern "C" int bpk_sync(void) {
    BPK_REQUIRE_INIT();
    BPK_CUDA(cudaStreamSynchronize(g_bpk.stream));
    return BPK_OK;
}

extern "C" uint64_t bpk_launch_count(void) { return g_bpk.launches; }

extern "C" int bpk_malloc(void **dev, uint64_t bytes) {
    BPK_REQUIRE_INIT();
    if (!dev) return bpk_set_error(BPK_EINVAL, "bpk_malloc: null out pointer");
    if (bytes == 0) bytes = 8;
    static const bool trace = getenv("BPK_TRACE_SLOW") != nullptr;       // diagnostics: allocations that reach the driver
    if (trace) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        cudaError_t e = cudaMallocAsync(dev, bytes, g_bpk.stream);
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double ms = 1e3 * (t1.tv_sec - t0.tv_sec) + 1e-6 * (t1.tv_nsec - t0.tv_nsec);
        if (ms > 2.0) fprintf(stderr, "[bpk] cudaMallocAsync(%.1f MB) took %.1f ms\n", bytes / 1048576.0, ms);
        if (e != cudaSuccess) return bpk_set_error(BPK_ECUDA, "cudaMallocAsync: %s", cudaGetErrorString(e));
        return BPK_OK;
    }
    BPK_CUDA(cudaMallocAsync(dev, bytes, g_bpk.stream));
    return BPK_OK;
}
extern "C" int bpk_free(void *dev) {
    if (!g_bpk.ready || !dev) return BPK_OK;
    BPK_CUDA(cudaFreeAsync(dev, g_bpk.stream));
    return BPK_OK;
}
extern "C" int bpk_h2d(void *dev, const void *host, uint64_t bytes) {
    BPK_REQUIRE_INIT();
    if (bytes == 0) return BPK_OK;
    BPK_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, g_bpk.stream));
    // pageable sources are staged by the driver before return; pinned ones are
    // truly async, so make the call safe for any caller-owned buffer:
    BPK_CUDA(cudaStreamSynchronize(g_bpk.stream));
    return BPK_OK;
}
extern "C" int bpk_d2h(void *host, const void *dev, uint64_t bytes) {
    BPK_REQUIRE_INIT();
    if (bytes == 0) return BPK_OK;
    BPK_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, g_bpk.stream));
    BPK_CUDA(cudaStreamSynchronize(g_bpk.stream));
    return BPK_OK;
}
extern "C" int bpk_d2d(void *dst, const void *src, uint64_t bytes) {
    BPK_REQUIRE_INIT();
    if (bytes == 0) return BPK_OK;
    BPK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_bpk.stream));
    return BPK_OK;
}
extern "C" int bpk_memset(void *dev, int byte, uint64_t bytes) {
    BPK_REQUIRE_INIT();
    if (bytes == 0) return BPK_OK;
    BPK_CUDA(cudaMemsetAsync(dev, byte, bytes, g_bpk.stream));
    return BPK_OK;
}
extern "C" int bpk_host_alloc(void **host, uint64_t bytes) {
    BPK_REQUIRE_INIT();
    BPK_CUDA(cudaMallocHost(host, bytes ? bytes : 8));
    return BPK_OK;
}
extern "C" int bpk_host_free(void *host) {
    if (host) BPK_CUDA(cudaFreeHost(host));
    return BPK_OK;
}

// ---- timers -----------------------------------------------------------------
static std::vector<cudaEvent_t> g_events;   // [2*id] start, [2*id+1] stop

extern "C" int bpk_timer_create(int *id) {
    BPK_REQUIRE_INIT();
    cudaEvent_t a, b;
    BPK_CUDA(cudaEventCreate(&a));
    BPK_CUDA(cudaEventCreate(&b));
    g_events.push_back(a);
    g_events.push_back(b);
    *id = (int)(g_events.size() / 2) - 1;
    return BPK_OK;
}
extern "C" int bpk_timer_record(int id, int which) {
    BPK_REQUIRE_INIT();
    if (id < 0 || 2 * id + 1 >= (int)g_events.size() || (which != 0 && which != 1))
        return bpk_set_error(BPK_EINVAL, "bad timer id");
    BPK_CUDA(cudaEventRecord(g_events[2 * id + which], g_bpk.stream));
    return BPK_OK;
}
extern "C" int bpk_timer_elapsed_ms(int id, double *ms) {
    BPK_REQUIRE_INIT();
    if (id < 0 || 2 * id + 1 >= (int)g_events.size()) return bpk_set_error(BPK_EINVAL, "bad timer id");
    BPK_CUDA(cudaEventSynchronize(g_events[2 * id + 1]));
    float f = 0.f;
    BPK_CUDA(cudaEventElapsedTime(&f, g_events[2 * id], g_events[2 * id + 1]));
    *ms = (double)f;
    return BPK_OK;
}

__global__ void bpk_l2_flush_kernel(double *p, size_t n, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t s = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += s) p[i] = v;
}
extern "C" int bpk_flush_l2(void) {
    BPK_REQUIRE_INIT();
    if (!g_bpk.l2buf) {
        g_bpk.l2bytes = 256u << 20;   // 2x the 126 MB L2
        BPK_CUDA(cudaMalloc(&g_bpk.l2buf, g_bpk.l2bytes));
    }
    BPK_LAUNCH(bpk_l2_flush_kernel, g_bpk.sm_count * 4, 512, 0,
               (double *)g_bpk.l2buf, g_bpk.l2bytes / 8, 1.0);
    return BPK_OK;
}

// ---- NCCL (loaded lazily so the library loads on boxes without it) --------
typedef struct { char internal[128]; } bpk_nccl_id;
typedef void *ncclComm_t_;
static void *g_nccl = nullptr;
static ncclComm_t_ g_comm = nullptr;
static int g_nranks = 1, g_rank = 0;
static int (*p_ncclGetUniqueId)(bpk_nccl_id *) = nullptr;
static int (*p_ncclCommInitRank)(ncclComm_t_ *, int, bpk_nccl_id, int) = nullptr;
static int (*p_ncclAllReduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
static int (*p_ncclCommDestroy)(ncclComm_t_) = nullptr;
static const char *(*p_ncclGetErrorString)(int) = nullptr;

static int load_nccl() {
    if (g_nccl) return BPK_OK;
    g_nccl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl) return bpk_set_error(BPK_ENCCL, "dlopen(libnccl.so.2): %s", dlerror());
    p_ncclGetUniqueId = (int (*)(bpk_nccl_id *))dlsym(g_nccl, "ncclGetUniqueId");
    p_ncclCommInitRank = (int (*)(ncclComm_t_ *, int, bpk_nccl_id, int))dlsym(g_nccl, "ncclCommInitRank");
    p_ncclAllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t))dlsym(g_nccl, "ncclAllReduce");
    p_ncclCommDestroy = (int (*)(ncclComm_t_))dlsym(g_nccl, "ncclCommDestroy");
    p_ncclGetErrorString = (const char *(*)(int))dlsym(g_nccl, "ncclGetErrorString");
    if (!p_ncclGetUniqueId || !p_ncclCommInitRank || !p_ncclAllReduce || !p_ncclCommDestroy)
        return bpk_set_error(BPK_ENCCL, "libnccl.so.2 lacks required symbols");
    return BPK_OK;
}
#define BPK_NCCL(call)                                                            \
    do {                                                                          \
        int r_ = (call);                                                          \
        if (r_ != 0)                                                              \
            return bpk_set_error(BPK_ENCCL, "%s failed: %s", #call,               \
                                 p_ncclGetErrorString ? p_ncclGetErrorString(r_) : "?"); \
    } while (0)

extern "C" int bpk_comm_unique_id(char id[128]) {
    int rc = load_nccl();
    if (rc) return rc;
    bpk_nccl_id u;
    BPK_NCCL(p_ncclGetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return BPK_OK;
}
extern "C" int bpk_comm_init(const char id[128], int nranks, int rank) {
    BPK_REQUIRE_INIT();
    int rc = load_nccl();
    if (rc) return rc;
    if (g_comm) return bpk_set_error(BPK_EINVAL, "communicator already initialised");
    bpk_nccl_id u;
    memcpy(u.internal, id, 128);
    BPK_NCCL(p_ncclCommInitRank(&g_comm, nranks, u, rank));
    g_nranks = nranks;
    g_rank = rank;
    return BPK_OK;
}
extern "C" int bpk_comm_size(int *nranks, int *rank) {
    if (nranks) *nranks = g_nranks;
    if (rank) *rank = g_rank;
    return BPK_OK;
}
extern "C" int bpk_allreduce_sum_f64(double *dev, uint64_t count) {
    BPK_REQUIRE_INIT();
    if (g_nranks == 1 && !g_comm) return BPK_OK;      // single rank: identity
    if (!g_comm) return bpk_set_error(BPK_ENCCL, "bpk_comm_init has not been called");
    // ncclFloat64 = 8, ncclSum = 0
    BPK_NCCL(p_ncclAllReduce(dev, dev, (size_t)count, 8, 0, g_comm, g_bpk.stream));
    g_bpk.launches++;
    return BPK_OK;
}
extern "C" int bpk_allreduce_sum_f64_oop(const double *src, double *dst, uint64_t count) {
    BPK_REQUIRE_INIT();
    if (g_nranks == 1 && !g_comm) {
        if (src != dst) BPK_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToDevice, g_bpk.stream));
        return BPK_OK;
    }
    if (!g_comm) return bpk_set_error(BPK_ENCCL, "bpk_comm_init has not been called");
    BPK_NCCL(p_ncclAllReduce(src, dst, (size_t)count, 8, 0, g_comm, g_bpk.stream));
    g_bpk.launches++;
    return BPK_OK;
}
// ---- peer-memory exchange window (in-kernel all-reduce over NVLink) ---------------------------
// Every rank cudaMalloc's one window and opens its peers' through CUDA IPC; kernels then
// store their partial statistics straight into every peer's window and spin on sequence
// flags (pca_vb_ops.cuh: STATS), so the sweep's one exchange costs no launch and no NCCL call.
BpkXchg g_xchg;

extern "C" int bpk_xchg_create(char handle[64]) {
    BPK_REQUIRE_INIT();
    if (!g_xchg.own) {
        BPK_CUDA(cudaMalloc((void **)&g_xchg.own, BPK_XCHG_BYTES));
        BPK_CUDA(cudaMemset(g_xchg.own, 0, BPK_XCHG_BYTES));
    }
    cudaIpcMemHandle_t h;
    BPK_CUDA(cudaIpcGetMemHandle(&h, g_xchg.own));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle, &h, 64);
    return BPK_OK;
}
// single-GPU runs: the same window, local only (the fused sweep kernel hands its reduced statistics to the
// tail through it with the same packet protocol the multi-GPU exchange uses)
int bpk_xchg_local(void) {
    BPK_REQUIRE_INIT();
    if (!g_xchg.own) {
        BPK_CUDA(cudaMalloc((void **)&g_xchg.own, BPK_XCHG_BYTES));
        BPK_CUDA(cudaMemset(g_xchg.own, 0, BPK_XCHG_BYTES));
    }
    if (!g_xchg.ready) g_xchg.win[0] = g_xchg.own;
    return BPK_OK;
}
extern "C" int bpk_xchg_open(const char *handles, int nranks, int rank) {
    BPK_REQUIRE_INIT();
    if (nranks < 1 || nranks > BPK_XCHG_MAXRANKS || rank < 0 || rank >= nranks)
        return bpk_set_error(BPK_EINVAL, "bpk_xchg_open: bad rank/nranks (max %d ranks)", BPK_XCHG_MAXRANKS);
    if (!g_xchg.own) return bpk_set_error(BPK_EINVAL, "bpk_xchg_open: bpk_xchg_create has not been called");
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { g_xchg.win[r] = g_xchg.own; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void *p = nullptr;
        BPK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        g_xchg.win[r] = (double *)p;
    }
    g_xchg.nranks = nranks;
    g_xchg.rank = rank;
    g_xchg.ready = true;
    return BPK_OK;
}
extern "C" int bpk_xchg_close(void) {
    if (!g_xchg.own) return BPK_OK;
    cudaStreamSynchronize(g_bpk.stream);
    for (int r = 0; r < g_xchg.nranks; ++r)
        if (r != g_xchg.rank && g_xchg.win[r]) cudaIpcCloseMemHandle(g_xchg.win[r]);
    cudaFree(g_xchg.own);
    g_xchg = BpkXchg();
    return BPK_OK;
}

extern "C" int bpk_comm_destroy(void) {
    if (g_comm) {
        cudaStreamSynchronize(g_bpk.stream);
        p_ncclCommDestroy(g_comm);
        g_comm = nullptr;
        g_nranks = 1;
        g_rank = 0;
    }
    return BPK_OK;
}
