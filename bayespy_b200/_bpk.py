"""ctypes binding of libbpk.so — the only door between the Python host layer and
the sm_100a kernels (include/bpk.h).  There is no CPU implementation behind
this module: if the shared library or a B200 is missing, ``get()`` raises.

Every method of :class:`CudaBackend` maps 1:1 onto the C entry point of the
same name (``bpk_`` prefix dropped); device pointers are plain ints.

Tests may install a checker backend with :func:`_set_backend_for_testing`
(the NumPy restatement in ``oracle/bpk_ref.py``); nothing in the product does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbpk.so")

MAXD = 8
MAXIN = 4
MAXDIM = 64
F64, U8 = 0, 1

OK, ECUDA, EINVAL, ENOTSPD, EDOMAIN, ENCCL, ENOGPU = range(7)

OPS = dict(COPY=0, ADD=1, SUB=2, MUL=3, DIV=4, AXPBY=5, AFFINE=6, FMA=7, WHERE=8,
           LOG=9, EXP=10, RECIP=11, SQUARE=12, SQRT=13, LGAMMA=14, DIGAMMA=15,
           MVLGAMMA=16, MVDIGAMMA=17, NONZERO=18, TRIGAMMA=19)

# name -> (restype, argtypes); the authoritative list of exported symbols
# (tests/test_abi.py checks it against include/bpk.h and the built library)
_i64p = C.POINTER(C.c_int64)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_dp = C.c_void_p     # device double*
PROTOTYPES = {
    "bpk_init": (C.c_int, [C.c_int]),
    "bpk_shutdown": (C.c_int, []),
    "bpk_last_error": (C.c_char_p, []),
    "bpk_device_info": (C.c_int, [_ip, _ip, _ip, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bpk_device_pci_bus_id": (C.c_int, [C.c_char_p, C.c_int]),
    "bpk_sync": (C.c_int, []),
    "bpk_launch_count": (C.c_uint64, []),
    "bpk_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "bpk_free": (C.c_int, [_vp]),
    "bpk_h2d": (C.c_int, [_vp, _vp, C.c_uint64]),
    "bpk_d2h": (C.c_int, [_vp, _vp, C.c_uint64]),
    "bpk_d2d": (C.c_int, [_vp, _vp, C.c_uint64]),
    "bpk_memset": (C.c_int, [_vp, C.c_int, C.c_uint64]),
    "bpk_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "bpk_host_free": (C.c_int, [_vp]),
    "bpk_timer_create": (C.c_int, [_ip]),
    "bpk_timer_record": (C.c_int, [C.c_int, C.c_int]),
    "bpk_timer_elapsed_ms": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "bpk_flush_l2": (C.c_int, []),
    "bpk_comm_unique_id": (C.c_int, [C.c_char_p]),
    "bpk_comm_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int]),
    "bpk_comm_size": (C.c_int, [_ip, _ip]),
    "bpk_allreduce_sum_f64": (C.c_int, [_dp, C.c_uint64]),
    "bpk_allreduce_sum_f64_oop": (C.c_int, [_dp, _dp, C.c_uint64]),
    "bpk_comm_destroy": (C.c_int, []),
    "bpk_xchg_create": (C.c_int, [C.c_char_p]),
    "bpk_xchg_open": (C.c_int, [C.c_char_p, C.c_int, C.c_int]),
    "bpk_xchg_close": (C.c_int, []),
    "bpk_ewise": (C.c_int, [C.c_int, C.c_int, _i64p, _dp, _i64p, C.c_int, C.POINTER(C.c_void_p), _ip, _i64p,
                            C.c_double, C.c_double]),
    "bpk_sum_multiply": (C.c_int, [C.c_int, _i64p, C.c_int, C.POINTER(C.c_void_p), _ip, _i64p, _dp, _i64p,
                                   C.c_double, C.c_int]),
    "bpk_chol": (C.c_int, [_dp, _dp, C.c_int64, C.c_int, C.c_int]),
    "bpk_chol_solve": (C.c_int, [_dp, C.c_int64, _dp, C.c_int64, _dp, C.c_int64, C.c_int, C.c_int]),
    "bpk_chol_inv": (C.c_int, [_dp, _dp, C.c_int64, C.c_int]),
    "bpk_chol_logdet": (C.c_int, [_dp, _dp, C.c_int64, C.c_int]),
    "bpk_block_banded_solve": (C.c_int, [_dp, _dp, _dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, _dp, _dp, C.c_int]),
    "bpk_gaussian_moments": (C.c_int, [_dp, C.c_int64, _dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, _dp, _dp,
                                       C.c_int]),
    "bpk_outer_add": (C.c_int, [_dp, _dp, C.c_int64, C.c_int64, C.c_int, _dp]),
    "bpk_gamma_moments": (C.c_int, [_dp, C.c_int64, _dp, C.c_int64, C.c_int64, _dp, _dp, _dp, C.c_int]),
    "bpk_wishart_moments": (C.c_int, [_dp, _dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, _dp, C.c_int]),
    "bpk_dirichlet_moments": (C.c_int, [_dp, C.c_int64, C.c_int, _dp, _dp, C.c_int]),
    "bpk_softmax_moments": (C.c_int, [_dp, C.c_int64, C.c_int, _dp, _dp]),
    "bpk_one_hot": (C.c_int, [_vp, C.c_int64, C.c_int, _dp, C.c_int]),
    "bpk_take": (C.c_int, [_dp, C.c_int64, C.c_int64, C.c_int64, _vp, C.c_int64, _dp]),
    "bpk_put_add": (C.c_int, [_dp, C.c_int64, C.c_int64, C.c_int64, _vp, _vp, C.c_int64, _dp]),
    "bpk_pca_xsweep": (C.c_int, [_dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, _dp, _dp]),
    "bpk_pca_stats": (C.c_int, [_dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp]),
    "bpk_pca_xsweep_masked": (C.c_int, [_dp, _vp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_double, _dp, _dp,
                                        _dp, _dp, _dp, _dp, C.c_int]),
    "bpk_pca_xsweep_masked_fused": (C.c_int, [_dp, _vp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_double, _dp, _dp,
                                              _dp, _dp, _dp, C.c_int]),
    "bpk_sumsq": (C.c_int, [_dp, _vp, C.c_int64, _dp]),
    "bpk_gmm_sweep": (C.c_int, [_dp, C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "bpk_gmm_stats": (C.c_int, [_dp, C.c_int64, C.c_int, C.c_int, _dp, _dp]),
    "bpk_pca_vb_layout": (C.c_int, [C.c_int, C.c_int, _i64p, _ip]),
    "bpk_pca_vb_field_name": (C.c_char_p, [C.c_int]),
    "bpk_pca_vb_run": (C.c_int, [_dp, C.c_int64, C.c_int64, C.c_int, _dp, _dp, _ip, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_double, _dp, C.c_int, _vp]),
    "bpk_pca_vb_set_mode": (C.c_int, [C.c_int, _ip]),
    "bpk_pca_vb_set_timers": (C.c_int, [_ip, C.c_int]),
    "bpk_pca_vb_timers_used": (C.c_int, []),
    "bpk_debug_stamps": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    "bpk_gmm_vb_layout": (C.c_int, [C.c_int, C.c_int, _i64p, _ip]),
    "bpk_gmm_vb_field_name": (C.c_char_p, [C.c_int]),
    "bpk_gmm_vb_run": (C.c_int, [_dp, C.c_int64, C.c_int, C.c_int, _dp, _dp, _dp, _ip, C.c_int, C.c_int, C.c_double,
                                 _dp, C.c_int, _vp]),
    "bpk_gmm_vb_set_timers": (C.c_int, [_ip, C.c_int]),
}

# opcodes of bpk_pca_vb_run (include/bpk.h)
VBOP = dict(XSWEEP=1, STATS=2, SXXT=3, XPRE=4, ROW=5, ALPHA=6, TAU=7, BOUND=8)
# opcodes of bpk_gmm_vb_run
GMMOP = dict(Z=1, MU=2, LAMBDA=3, ALPHA=4, BOUND=5)


class BpkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class NotPositiveDefinite(Exception):
    """Mirrors the bare Exception("Matrix not positive definite") of linalg.py:58-59."""


def _raise(code, msg):
    if code == ENOTSPD:
        raise NotPositiveDefinite("Matrix not positive definite")
    if code == EDOMAIN:
        raise ValueError(msg or "Natural parameters should be positive")
    if code == EINVAL:
        raise ValueError(msg)
    raise BpkError(code, msg)


def _i64(seq):
    return (C.c_int64 * max(len(seq), 1))(*[int(v) for v in seq])


class CudaBackend:
    """libbpk.so bound to one GPU."""

    name = "cuda"

    def __init__(self, device=0):
        if not os.path.exists(LIB_PATH):
            raise BpkError(ENOGPU, "libbpk.so is not built (%s); run __graft_entry__.build(). "
                                   "bayespy_b200 has no CPU fallback." % LIB_PATH)
        self.lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(self.lib, name)
            fn.restype = res
            fn.argtypes = args
        self._chk(self.lib.bpk_init(int(device)))
        self.device = device

    # -- plumbing
    def _chk(self, rc):
        if rc != 0:
            _raise(rc, (self.lib.bpk_last_error() or b"").decode())

    def sync(self):
        self._chk(self.lib.bpk_sync())

    def launch_count(self):
        return int(self.lib.bpk_launch_count())

    def device_info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        t, f = C.c_uint64(), C.c_uint64()
        self._chk(self.lib.bpk_device_info(C.byref(a), C.byref(b), C.byref(c), C.byref(t), C.byref(f)))
        return dict(sm_count=a.value, cc=(b.value, c.value), hbm_total=t.value, hbm_free=f.value)

    def pci_bus_id(self):
        buf = C.create_string_buffer(32)
        self._chk(self.lib.bpk_device_pci_bus_id(buf, 32))
        return buf.value.decode().lower()

    def malloc(self, nbytes):
        p = C.c_void_p()
        self._chk(self.lib.bpk_malloc(C.byref(p), int(nbytes)))
        return p.value or 0

    def free(self, ptr):
        if ptr:
            self.lib.bpk_free(ptr)

    def h2d(self, dev, host_array):
        a = np.ascontiguousarray(host_array)
        self._chk(self.lib.bpk_h2d(dev, a.ctypes.data, a.nbytes))   # a outlives the (synchronous) copy

    def h2d_ptr(self, dev, host_ptr, nbytes):
        self._chk(self.lib.bpk_h2d(dev, host_ptr, nbytes))

    def d2h(self, host_array, dev):
        assert host_array.flags.c_contiguous
        self._chk(self.lib.bpk_d2h(host_array.ctypes.data, dev, host_array.nbytes))

    def d2d(self, dst, src, nbytes):
        self._chk(self.lib.bpk_d2d(dst, src, int(nbytes)))

    def memset(self, dev, byte, nbytes):
        self._chk(self.lib.bpk_memset(dev, int(byte), int(nbytes)))

    def host_alloc(self, nbytes):
        p = C.c_void_p()
        self._chk(self.lib.bpk_host_alloc(C.byref(p), int(nbytes)))
        return p.value

    def host_free(self, ptr):
        self._chk(self.lib.bpk_host_free(ptr))

    def timer_create(self):
        i = C.c_int()
        self._chk(self.lib.bpk_timer_create(C.byref(i)))
        return i.value

    def timer_record(self, tid, which):
        self._chk(self.lib.bpk_timer_record(tid, which))

    def timer_elapsed_ms(self, tid):
        d = C.c_double()
        self._chk(self.lib.bpk_timer_elapsed_ms(tid, C.byref(d)))
        return d.value

    def flush_l2(self):
        self._chk(self.lib.bpk_flush_l2())

    # -- NCCL
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self._chk(self.lib.bpk_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, uid, nranks, rank):
        self._chk(self.lib.bpk_comm_init(C.create_string_buffer(bytes(uid), 128), nranks, rank))

    def comm_size(self):
        a, b = C.c_int(), C.c_int()
        self._chk(self.lib.bpk_comm_size(C.byref(a), C.byref(b)))
        return a.value, b.value

    def allreduce_sum_f64(self, dev, count):
        self._chk(self.lib.bpk_allreduce_sum_f64(dev, int(count)))

    def allreduce_sum_f64_oop(self, src, dst, count):
        self._chk(self.lib.bpk_allreduce_sum_f64_oop(src, dst, int(count)))

    def comm_destroy(self):
        self._chk(self.lib.bpk_comm_destroy())

    def xchg_create(self):
        buf = C.create_string_buffer(64)
        self._chk(self.lib.bpk_xchg_create(buf))
        return buf.raw

    def xchg_open(self, handles, nranks, rank):
        blob = b"".join(handles)
        assert len(blob) == 64 * nranks
        self._chk(self.lib.bpk_xchg_open(C.create_string_buffer(blob, len(blob)), nranks, rank))

    def xchg_close(self):
        self._chk(self.lib.bpk_xchg_close())

    # -- generic kernels
    def ewise(self, op, shape, out, out_stride, ins, dtypes, in_strides, alpha=0.0, beta=0.0):
        nd, n_in = len(shape), len(ins)
        flat = [s for st in in_strides for s in st]
        ptrs = (C.c_void_p * n_in)(*ins)
        dts = (C.c_int * n_in)(*dtypes)
        self._chk(self.lib.bpk_ewise(op, nd, _i64(shape), out, _i64(out_stride), n_in, ptrs, dts, _i64(flat),
                                     float(alpha), float(beta)))

    def sum_multiply(self, shape, ins, dtypes, in_strides, out, out_stride, scale=1.0, accumulate=False):
        nd, n_in = len(shape), len(ins)
        flat = [s for st in in_strides for s in st]
        ptrs = (C.c_void_p * n_in)(*ins)
        dts = (C.c_int * n_in)(*dtypes)
        self._chk(self.lib.bpk_sum_multiply(nd, _i64(shape), n_in, ptrs, dts, _i64(flat), out, _i64(out_stride),
                                            float(scale), int(bool(accumulate))))

    # -- linalg
    def chol(self, A, U, batch, D, check=True):
        self._chk(self.lib.bpk_chol(A, U, batch, D, int(check)))

    def chol_solve(self, U, batchU, B, batchB, X, batch, D, nrhs):
        self._chk(self.lib.bpk_chol_solve(U, batchU, B, batchB, X, batch, D, nrhs))

    def chol_inv(self, U, Ainv, batch, D):
        self._chk(self.lib.bpk_chol_inv(U, Ainv, batch, D))

    def chol_logdet(self, U, out, batch, D):
        self._chk(self.lib.bpk_chol_logdet(U, out, batch, D))

    def block_banded_solve(self, A, B, y, batch, T, D, V, Cb, x, logdet, check=True):
        self._chk(self.lib.bpk_block_banded_solve(A, B, y, batch, T, D, V, Cb, x, logdet, int(check)))

    # -- node kernels
    def gaussian_moments(self, phi0, n0, phi1, n1, N, K, u0, cov, g, logdet, check=True):
        self._chk(self.lib.bpk_gaussian_moments(phi0, n0, phi1, n1, N, K, u0, cov, g, logdet, int(check)))

    def outer_add(self, u0, cov, ncov, N, K, u1):
        self._chk(self.lib.bpk_outer_add(u0, cov, ncov, N, K, u1))

    def gamma_moments(self, phi0, n0, phi1, n1, n, u0, u1, g, check=True):
        self._chk(self.lib.bpk_gamma_moments(phi0, n0, phi1, n1, n, u0, u1, g, int(check)))

    def wishart_moments(self, phi0, phi1, n1, n, D, u0, u1, g, check=True):
        self._chk(self.lib.bpk_wishart_moments(phi0, phi1, n1, n, D, u0, u1, g, int(check)))

    def dirichlet_moments(self, phi, n, K, u, g, check=True):
        self._chk(self.lib.bpk_dirichlet_moments(phi, n, K, u, g, int(check)))

    def softmax_moments(self, phi, n, K, u, g):
        self._chk(self.lib.bpk_softmax_moments(phi, n, K, u, g))

    def one_hot(self, labels, n, K, u, check=True):
        self._chk(self.lib.bpk_one_hot(labels, n, K, u, int(check)))

    def take(self, src, pre, L, post, idx, J, out):
        self._chk(self.lib.bpk_take(src, pre, L, post, idx, J, out))

    def put_add(self, src, pre, J, post, order, start, L, out):
        self._chk(self.lib.bpk_put_add(src, pre, J, post, order, start, L, out))

    # -- fused sweeps
    def pca_xsweep(self, Y, M, N, K, A, b, X, stats):
        self._chk(self.lib.bpk_pca_xsweep(Y, M, N, K, A, b, X, stats))

    def pca_stats(self, Y, M, N, K, X, stats):
        self._chk(self.lib.bpk_pca_stats(Y, M, N, K, X, stats))

    def pca_xsweep_masked(self, Y, mask, M, N, K, W, WW, tau, alpha, amu, X, COV, g, stats, check=True):
        self._chk(self.lib.bpk_pca_xsweep_masked(Y, mask, M, N, K, W, WW, float(tau), alpha, amu, X, COV, g, stats,
                                                 int(check)))

    def pca_xsweep_masked_fused(self, Y, mask, M, N, K, W, WW, tau, alpha, amu, X, g, stats, check=True):
        self._chk(self.lib.bpk_pca_xsweep_masked_fused(Y, mask, M, N, K, W, WW, float(tau), alpha, amu, X, g, stats,
                                                       int(check)))

    def sumsq(self, Y, mask, count, out2):
        self._chk(self.lib.bpk_sumsq(Y, mask, count, out2))

    def gmm_sweep(self, Y, N, D, K, c, h, Lam, logpi, P, g, stats):
        self._chk(self.lib.bpk_gmm_sweep(Y, N, D, K, c, h, Lam, logpi, P, g, stats))

    def gmm_stats(self, Y, N, D, K, P, stats):
        self._chk(self.lib.bpk_gmm_stats(Y, N, D, K, P, stats))

    # -- device-resident VB loop of the factor model
    def pca_vb_layout(self, M, K):
        """{field: (offset, size)} of the fp64 state vector, and its total length."""
        n = C.c_int()
        self._chk(self.lib.bpk_pca_vb_layout(M, K, None, C.byref(n)))
        off = (C.c_int64 * (n.value + 1))()
        self._chk(self.lib.bpk_pca_vb_layout(M, K, off, C.byref(n)))
        names = [self.lib.bpk_pca_vb_field_name(i).decode() for i in range(n.value)]
        return {nm: (off[i], off[i + 1] - off[i]) for i, nm in enumerate(names)}, off[n.value]

    def pca_vb_run(self, Y, M, N, K, X, state, ops, niter, has_alpha, has_tau, tol, Lhist, cap, ctrl):
        arr = (C.c_int * len(ops))(*ops)
        self._chk(self.lib.bpk_pca_vb_run(Y, M, N, K, X, state, arr, len(ops), int(niter), int(has_alpha),
                                          int(has_tau), float(tol), Lhist, int(cap), ctrl))

    def pca_vb_set_mode(self, no_loop):
        prev = C.c_int()
        self._chk(self.lib.bpk_pca_vb_set_mode(int(bool(no_loop)), C.byref(prev)))
        return bool(prev.value)

    def pca_vb_set_timers(self, ids):
        arr = (C.c_int * max(len(ids), 1))(*ids)
        self._chk(self.lib.bpk_pca_vb_set_timers(arr, len(ids)))

    def pca_vb_timers_used(self):
        return int(self.lib.bpk_pca_vb_timers_used())

    # -- device-resident VB loop of the Gaussian mixture
    def gmm_vb_layout(self, D, K):
        """{field: (offset, padded size)} of the fp64 state vector, and its total length."""
        n = C.c_int()
        self._chk(self.lib.bpk_gmm_vb_layout(D, K, None, C.byref(n)))
        off = (C.c_int64 * (n.value + 1))()
        self._chk(self.lib.bpk_gmm_vb_layout(D, K, off, C.byref(n)))
        names = [self.lib.bpk_gmm_vb_field_name(i).decode() for i in range(n.value)]
        return {nm: (off[i], off[i + 1] - off[i]) for i, nm in enumerate(names)}, off[n.value]

    def gmm_vb_run(self, Y, N, D, K, P, gz, state, ops, niter, tol, Lhist, cap, ctrl):
        arr = (C.c_int * len(ops))(*ops)
        self._chk(self.lib.bpk_gmm_vb_run(Y, N, D, K, P, gz, state, arr, len(ops), int(niter), float(tol), Lhist,
                                          int(cap), ctrl))

    def gmm_vb_set_timers(self, ids):
        arr = (C.c_int * max(len(ids), 1))(*ids)
        self._chk(self.lib.bpk_gmm_vb_set_timers(arr, len(ids)))

    def debug_stamps(self, n=32):
        out = (C.c_uint64 * n)()
        self._chk(self.lib.bpk_debug_stamps(out, n))
        return [int(v) for v in out]


_backend = None


def get():
    """The process-wide backend; created on first use.  Raises without a GPU."""
    global _backend
    if _backend is None:
        dev = int(os.environ.get("BPK_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _backend = CudaBackend(dev)
    return _backend


def is_cuda():
    return _backend is not None and _backend.name == "cuda"


def _set_backend_for_testing(backend):
    """TEST INFRASTRUCTURE ONLY: install a checker backend (oracle/bpk_ref.py) so
    that the host-side graph logic can be exercised on a box without a GPU.
    Never called by the product; results obtained this way are not parity claims."""
    global _backend
    old = _backend
    _backend = backend
    return old
