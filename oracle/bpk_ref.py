"""TEST INFRASTRUCTURE — NumPy/SciPy restatement of every libbpk entry point.

This is the per-kernel oracle: each method restates, on the CPU, the reference
arithmetic that the CUDA kernel of the same name replaces (reference file:line
cited per method, paths relative to /root/reference).  It is pinned against the
reference itself by ``tests/golden/make_golden.py`` (run in the build container,
where the reference imports) and the committed ``tests/golden/*.npz`` vectors.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg
may import this module.  "Device pointers" are raw host addresses of buffers it
allocates, so pointer arithmetic done by the host layer works unchanged.
"""
import ctypes

import numpy as np
import scipy.linalg
import scipy.special as sp

F64, U8 = 0, 1
ENOTSPD, EDOMAIN = 3, 4


class NotPositiveDefinite(Exception):
    pass


def _view(ptr, shape, strides_elems, dtype=np.float64):
    """ndarray view of raw memory at ``ptr`` with element strides."""
    shape = tuple(int(s) for s in shape)
    itemsize = np.dtype(dtype).itemsize
    if any(s == 0 for s in shape):
        return np.empty(shape, dtype)
    extent = 1 + sum((n - 1) * abs(int(st)) for n, st in zip(shape, strides_elems))
    buf = (ctypes.c_byte * (extent * itemsize)).from_address(int(ptr))
    base = np.frombuffer(buf, dtype=dtype)
    return np.lib.stride_tricks.as_strided(base, shape, [int(st) * itemsize for st in strides_elems])


def _dense(ptr, shape, dtype=np.float64):
    shape = tuple(int(s) for s in shape)
    st = []
    acc = 1
    for n in reversed(shape):
        st.append(acc)
        acc *= n
    return _view(ptr, shape, list(reversed(st)), dtype)


class RefBackend:
    """Same method surface as bayespy_b200._bpk.CudaBackend."""

    name = "oracle"

    def __init__(self):
        self._bufs = {}
        self._launches = 0
        self._timers = []
        self.nranks, self.rank = 1, 0
        self._allreduce_hook = None   # tests install a gloo all-reduce here

    # ---- plumbing -------------------------------------------------------------
    def sync(self):
        pass

    def launch_count(self):
        return self._launches

    def device_info(self):
        return dict(sm_count=0, cc=(0, 0), hbm_total=0, hbm_free=0)

    def malloc(self, nbytes):
        a = np.zeros(max(int(nbytes), 8) + 64, dtype=np.uint8)
        addr = a.ctypes.data
        off = (-addr) % 64
        p = addr + off
        self._bufs[p] = a
        a[:] = 0xFF   # poison: unwritten doubles read as NaN
        return p

    def free(self, ptr):
        self._bufs.pop(ptr, None)

    def h2d(self, dev, host_array):
        a = np.ascontiguousarray(host_array)
        ctypes.memmove(dev, a.ctypes.data, a.nbytes)

    def d2h(self, host_array, dev):
        ctypes.memmove(host_array.ctypes.data, dev, host_array.nbytes)

    def d2d(self, dst, src, nbytes):
        ctypes.memmove(dst, src, int(nbytes))

    def memset(self, dev, byte, nbytes):
        ctypes.memset(dev, int(byte), int(nbytes))

    def host_alloc(self, nbytes):
        return self.malloc(nbytes)

    def host_free(self, ptr):
        self.free(ptr)

    def timer_create(self):
        self._timers.append([0.0, 0.0])
        return len(self._timers) - 1

    def timer_record(self, tid, which):
        import time
        self._timers[tid][which] = time.perf_counter()

    def timer_elapsed_ms(self, tid):
        return 1e3 * (self._timers[tid][1] - self._timers[tid][0])

    def flush_l2(self):
        pass

    def comm_unique_id(self):
        return b"\0" * 128

    def comm_init(self, uid, nranks, rank):
        self.nranks, self.rank = nranks, rank

    def comm_size(self):
        return self.nranks, self.rank

    def allreduce_sum_f64(self, dev, count):
        if self._allreduce_hook is not None:
            v = _dense(dev, (count,))
            v[...] = self._allreduce_hook(v.copy())

    def allreduce_sum_f64_oop(self, src, dst, count):
        v = _dense(src, (count,)).copy()
        if self._allreduce_hook is not None:
            v = self._allreduce_hook(v)
        _dense(dst, (count,))[...] = v

    def comm_destroy(self):
        self.nranks, self.rank = 1, 0

    # ---- generic kernels --------------------------------------------------------
    def ewise(self, op, shape, out, out_stride, ins, dtypes, in_strides, alpha=0.0, beta=0.0):
        self._launches += 1
        arrs = [_view(p, shape, st, np.uint8 if dt == U8 else np.float64).astype(np.float64)
                for p, dt, st in zip(ins, dtypes, in_strides)]
        a = arrs[0]
        b = arrs[1] if len(arrs) > 1 else None
        c = arrs[2] if len(arrs) > 2 else None
        with np.errstate(all="ignore"):
            if op == 0: r = a
            elif op == 1: r = a + b
            elif op == 2: r = a - b
            elif op == 3: r = a * b
            elif op == 4: r = a / b
            elif op == 5: r = alpha * a + beta * b
            elif op == 6: r = alpha * a + beta
            elif op == 7: r = alpha * a * b + beta * c
            elif op == 8: r = np.where(a != 0, b, c)
            elif op == 9: r = np.log(a)
            elif op == 10: r = np.exp(a)
            elif op == 11: r = alpha / a
            elif op == 12: r = a * a
            elif op == 13: r = np.sqrt(a)
            elif op == 14: r = sp.gammaln(a)                       # gamma.py:147
            elif op == 15: r = sp.digamma(a)                       # gamma.py:145
            elif op == 16: r = sp.multigammaln(a, int(alpha)) if np.ndim(a) == 0 else \
                np.vectorize(lambda v: sp.multigammaln(v, int(alpha)))(a)   # wishart.py:187
            elif op == 17: r = np.sum(sp.digamma(a[..., None] - 0.5 * np.arange(int(alpha))), axis=-1)  # misc.py:1146
            elif op == 18: r = np.where(a != 0, b, 0.0)            # expfamily.py:463
            elif op == 19: r = sp.polygamma(1, a)                  # gamma.py:210
            else:
                raise ValueError("bad op")
        o = _view(out, shape, out_stride)
        o[...] = r

    def sum_multiply(self, shape, ins, dtypes, in_strides, out, out_stride, scale=1.0, accumulate=False):
        """misc.py:851-933 (np.einsum over broadcast operands) / node.py:650."""
        self._launches += 1
        nd = len(shape)
        arrs = [_view(p, shape, st, np.uint8 if dt == U8 else np.float64).astype(np.float64)
                for p, dt, st in zip(ins, dtypes, in_strides)]
        keys = list(range(nd))
        kept = [d for d in keys if out_stride[d] != 0 or shape[d] == 1]
        args = []
        for a in arrs:
            args += [a, keys]
        r = np.einsum(*args, kept) if nd > 0 else np.prod([a for a in arrs])
        kshape = [shape[d] for d in kept]
        kstr = [out_stride[d] for d in kept]
        o = _view(out, kshape, kstr)
        if any(s == 0 for s in shape):
            return
        if accumulate:
            o[...] = o + scale * r
        else:
            o[...] = scale * r

    # ---- linalg (bayespy/utils/linalg.py) ----------------------------------------
    def chol(self, A, U, batch, D, check=True):
        """linalg.py:31-63: per-matrix scipy cho_factor (upper)."""
        self._launches += 1
        a = _dense(A, (batch, D, D))
        u = _dense(U, (batch, D, D))
        bad = False
        for i in range(batch):
            try:
                f = scipy.linalg.cho_factor(a[i])[0]
                u[i] = np.triu(f)
            except (np.linalg.LinAlgError, ValueError):
                u[i] = np.nan
                bad = True
        if bad and check:
            raise NotPositiveDefinite("Matrix not positive definite")   # linalg.py:58-59

    def chol_solve(self, U, batchU, B, batchB, X, batch, D, nrhs):
        """linalg.py:66-171: per-matrix cho_solve."""
        self._launches += 1
        u = _dense(U, (batchU, D, D))
        b = _dense(B, (batchB, D, nrhs))
        x = _dense(X, (batch, D, nrhs))
        for i in range(batch):
            x[i] = scipy.linalg.cho_solve((u[0 if batchU == 1 else i], False), b[0 if batchB == 1 else i])

    def chol_inv(self, U, Ainv, batch, D):
        """linalg.py:174-207."""
        self._launches += 1
        u = _dense(U, (batch, D, D))
        o = _dense(Ainv, (batch, D, D))
        I = np.identity(D)
        for i in range(batch):
            o[i] = scipy.linalg.cho_solve((u[i], False), I)

    def chol_logdet(self, U, out, batch, D):
        """linalg.py:209-223."""
        self._launches += 1
        u = _dense(U, (batch, D, D))
        _dense(out, (batch,))[...] = 2 * np.sum(np.log(np.einsum("...ii->...i", u)), axis=-1)

    # ---- node kernels ----------------------------------------------------------------
    def gaussian_moments(self, phi0, n0, phi1, n1, N, K, u0, cov, g, logdet, check=True):
        """gaussian.py:672-706 / :397-446."""
        self._launches += 1
        p0 = _dense(phi0, (n0, K))
        p1 = _dense(phi1, (n1, K, K))
        covs = np.empty((n1, K, K))
        lds = np.empty(n1)
        I = np.identity(K)
        bad = False
        for i in range(n1):
            try:
                f = scipy.linalg.cho_factor(-2 * p1[i])
                covs[i] = scipy.linalg.cho_solve(f, I)
                lds[i] = 2 * np.sum(np.log(np.diag(f[0])))
            except (np.linalg.LinAlgError, ValueError):
                covs[i] = np.nan
                lds[i] = np.nan
                bad = True
        if cov:
            _dense(cov, (n1, K, K))[...] = covs
        if logdet:
            _dense(logdet, (n1,))[...] = lds
        P0 = np.broadcast_to(p0, (N, K))
        C = np.broadcast_to(covs, (N, K, K))
        m = np.einsum("nij,nj->ni", C, P0)
        if u0:
            _dense(u0, (N, K))[...] = m
        if g:
            _dense(g, (N,))[...] = -0.5 * np.einsum("ni,ni->n", m, P0) + 0.5 * np.broadcast_to(lds, (N,))
        if bad and check:
            raise NotPositiveDefinite("Matrix not positive definite")

    def outer_add(self, u0, cov, ncov, N, K, u1):
        """gaussian.py:695: u1 = outer(u0,u0) + Cov."""
        self._launches += 1
        m = _dense(u0, (N, K))
        r = m[:, :, None] * m[:, None, :]
        if cov:
            r = r + np.broadcast_to(_dense(cov, (ncov, K, K)), (N, K, K))
        _dense(u1, (N, K, K))[...] = r

    def gamma_moments(self, phi0, n0, phi1, n1, n, u0, u1, g, check=True):
        """gamma.py:124-148."""
        self._launches += 1
        p0 = np.broadcast_to(_dense(phi0, (n0,)), (n,))
        p1 = np.broadcast_to(_dense(phi1, (n1,)), (n,))
        bad = bool(np.any(~(-p0 > 0)) or np.any(~(p1 > 0)))
        with np.errstate(all="ignore"):
            logb = np.log(-p0)
            if u0: _dense(u0, (n,))[...] = p1 / (-p0)
            if u1: _dense(u1, (n,))[...] = sp.digamma(p1) - logb
            if g: _dense(g, (n,))[...] = p1 * logb - sp.gammaln(p1)
        if bad and check:
            raise ValueError("Natural parameters should be positive")

    def wishart_moments(self, phi0, phi1, n1, n, D, u0, u1, g, check=True):
        """wishart.py:165-188."""
        self._launches += 1
        p0 = _dense(phi0, (n, D, D))
        p1 = np.broadcast_to(_dense(phi1, (n1,)), (n,))
        I = np.identity(D)
        bad = False
        for i in range(n):
            try:
                f = scipy.linalg.cho_factor(-p0[i])
                ld = 2 * np.sum(np.log(np.diag(f[0])))
                inv = scipy.linalg.cho_solve(f, I)
            except (np.linalg.LinAlgError, ValueError):
                ld, inv, bad = np.nan, np.full((D, D), np.nan), True
            if u0: _dense(u0, (n, D, D))[i] = p1[i] * inv
            if u1: _dense(u1, (n,))[i] = -ld + np.sum(sp.digamma(p1[i] - 0.5 * np.arange(D)))
            if g: _dense(g, (n,))[i] = p1[i] * ld - sp.multigammaln(p1[i], D)
        if bad and check:
            raise NotPositiveDefinite("Matrix not positive definite")

    def dirichlet_moments(self, phi, n, K, u, g, check=True):
        """dirichlet.py:130-160."""
        self._launches += 1
        p = _dense(phi, (n, K))
        bad = bool(np.any(~(p > 0)))
        with np.errstate(all="ignore"):
            s = np.sum(p, axis=-1)
            if u: _dense(u, (n, K))[...] = sp.psi(p) - sp.psi(s)[:, None]
            if g: _dense(g, (n,))[...] = sp.gammaln(s) - np.sum(sp.gammaln(p), axis=-1)
        if bad and check:
            raise ValueError("Natural parameters should be positive")

    def softmax_moments(self, phi, n, K, u, g):
        """multinomial.py:101-121 with misc.py:1366-1401 (max-shift, second renormalisation)."""
        self._launches += 1
        p = _dense(phi, (n, K))
        with np.errstate(all="ignore"):
            m = np.amax(p, axis=-1, keepdims=True)
            m = np.where(np.isfinite(m), m, 0.0)
            lse = np.log(np.sum(np.exp(p - m), axis=-1, keepdims=True)) + m
            q = np.exp(p - lse)
            if u: _dense(u, (n, K))[...] = q / np.sum(q, axis=-1, keepdims=True)
            if g: _dense(g, (n,))[...] = -lse[:, 0]

    def one_hot(self, labels, n, K, u, check=True):
        """categorical.py:30-47."""
        self._launches += 1
        l = _dense(labels, (n,), np.int64)
        if check and (np.any(l < 0) or np.any(l >= K)):
            raise ValueError("Invalid category index")
        o = np.zeros((n, K))
        ok = (l >= 0) & (l < K)
        o[np.nonzero(ok)[0], l[ok]] = 1.0
        _dense(u, (n, K))[...] = o

    def take(self, src, pre, L, post, idx, J, out):
        """take.py:63-73 (np.take along one plate axis)."""
        self._launches += 1
        x = _dense(src, (pre, L, post))
        ix = _dense(idx, (J,), np.int64)
        _dense(out, (pre, J, post))[...] = x[:, ix, :]

    def put_add(self, src, pre, J, post, order, start, L, out):
        """misc.py:549-585 put_simple / take.py:76-88: inverse of take with accumulation."""
        self._launches += 1
        x = _dense(src, (pre, J, post))
        od = _dense(order, (J,), np.int64)
        stt = _dense(start, (L + 1,), np.int64)
        o = _dense(out, (pre, L, post))
        for i in range(L):
            o[:, i, :] = x[:, od[stt[i]:stt[i + 1]], :].sum(axis=1) if stt[i + 1] > stt[i] else 0.0

    # ---- fused sweeps -----------------------------------------------------------------
    def pca_xsweep(self, Y, M, N, K, A, b, X, stats):
        """x_n = A y_n + b and the plate sums of dot.py:581 / :355,403 (see csrc/pca.cu)."""
        self._launches += 1
        y = _dense(Y, (M, N))
        a = _dense(A, (K, M))
        x = y.T @ a.T
        if b:
            x = x + _dense(b, (K,))
        _dense(X, (N, K))[...] = x
        self._pca_stats(y, x, M, N, K, stats)

    def pca_stats(self, Y, M, N, K, X, stats):
        self._launches += 1
        self._pca_stats(_dense(Y, (M, N)), _dense(X, (N, K)), M, N, K, stats)

    @staticmethod
    def _pca_stats(y, x, M, N, K, stats):
        s = _dense(stats, (M * K + K * K + K,))
        s[:M * K] += (y @ x).ravel()
        s[M * K:M * K + K * K] += (x.T @ x).ravel()
        s[M * K + K * K:] += x.sum(axis=0)

    def pca_xsweep_masked(self, Y, mask, M, N, K, W, WW, tau, alpha, amu, X, COV, g, stats, check=True):
        """Per-column precision path (gaussian.py:672-706 with per-plate phi1; dot.py:581 with a mask)."""
        self._launches += 1
        y = _dense(Y, (M, N))
        mk = _dense(mask, (M, N), np.uint8).astype(np.float64)
        w = _dense(W, (M, K))
        ww = _dense(WW, (M, K, K))
        al = _dense(alpha, (K,))
        am = _dense(amu, (K,)) if amu else np.zeros(K)
        Lam = np.diag(al)[None] + tau * np.einsum("mn,mij->nij", mk, ww)
        phi0 = tau * np.einsum("mn,mn,mk->nk", mk, y, w) + am
        cov = np.linalg.inv(Lam)
        x = np.einsum("nij,nj->ni", cov, phi0)
        _dense(X, (N, K))[...] = x
        if COV:
            _dense(COV, (N, K, K))[...] = cov
        if g:
            _dense(g, (N,))[...] = -0.5 * np.einsum("ni,ni->n", x, phi0) + 0.5 * np.linalg.slogdet(Lam)[1]
        xx = cov + x[:, :, None] * x[:, None, :]
        s = _dense(stats, (M * K + M * K * K,))
        s[:M * K] += np.einsum("mn,mn,nk->mk", mk, y, x).ravel()
        s[M * K:] += np.einsum("mn,nij->mij", mk, xx).ravel()

    def pca_xsweep_masked_fused(self, Y, mask, M, N, K, W, WW, tau, alpha, amu, X, g, stats, check=True):
        """csrc/pca_masked.cu restated: per-column precision (gaussian.py:672-706 with per-plate phi1, the per-plate
        loops of linalg.py:50-59,111-146,185-195) and the masked plate sums of dot.py:581 / node.py:650."""
        self._launches += 1
        y = _dense(Y, (M, N))
        mk = _dense(mask, (M, N), np.uint8).astype(np.float64)
        w = _dense(W, (M, K))
        ww = _dense(WW, (M, K, K))
        al = _dense(alpha, (K,))
        am = _dense(amu, (K,)) if amu else np.zeros(K)
        Lam = np.diag(al)[None] + tau * np.einsum("mn,mij->nij", mk, ww)
        phi0 = tau * np.einsum("mn,mn,mk->nk", mk, y, w) + am
        if check and np.any(np.linalg.eigvalsh(Lam)[:, 0] <= 0):
            raise NotPositiveDefinite("Matrix not positive definite")
        cov = np.linalg.inv(Lam)
        x = np.einsum("nij,nj->ni", cov, phi0)
        _dense(X, (N, K))[...] = x
        q = np.einsum("ni,ni->n", x, phi0)
        ld = np.linalg.slogdet(Lam)[1]
        if g:
            _dense(g, (N,))[...] = -0.5 * q + 0.5 * ld
        xx = cov + x[:, :, None] * x[:, None, :]
        s = _dense(stats, (M * K + M * K * K + K * K + K + 2,))
        o = 0
        s[o:o + M * K] += np.einsum("mn,mn,nk->mk", mk, y, x).ravel(); o += M * K
        s[o:o + M * K * K] += np.einsum("mn,nij->mij", mk, xx).ravel(); o += M * K * K
        s[o:o + K * K] += xx.sum(axis=0).ravel(); o += K * K
        s[o:o + K] += x.sum(axis=0); o += K
        s[o] += q.sum()
        s[o + 1] += ld.sum()

    def sumsq(self, Y, mask, count, out2):
        self._launches += 1
        y = _dense(Y, (count,))
        o = _dense(out2, (2,))
        if mask:
            mk = _dense(mask, (count,), np.uint8) != 0
            o[0] = np.sum(y[mk] ** 2)
            o[1] = float(np.sum(mk))
        else:
            o[0] = np.sum(y ** 2)
            o[1] = float(count)

    def gmm_sweep(self, Y, N, D, K, c, h, Lam, logpi, P, g, stats):
        """mixture.py:53-160 (index 0 and index>=1 messages) + multinomial.py:101-121, fused."""
        self._launches += 1
        y = _dense(Y, (N, D))
        cc = _dense(c, (K,))
        hh = _dense(h, (K, D))
        LL = _dense(Lam, (K, D, D))
        lp = _dense(logpi, (K,))
        L = cc[None] + y @ hh.T - 0.5 * np.einsum("ni,kij,nj->nk", y, LL, y) + lp[None]
        m = np.amax(L, axis=-1, keepdims=True)
        lse = np.log(np.sum(np.exp(L - m), axis=-1, keepdims=True)) + m
        q = np.exp(L - lse)
        p = q / np.sum(q, axis=-1, keepdims=True)
        if P: _dense(P, (N, K))[...] = p
        if g: _dense(g, (N,))[...] = -lse[:, 0]
        s = _dense(stats, (K + K * D + K * D * D + 1,))
        s[:K] += p.sum(axis=0)
        s[K:K + K * D] += (p.T @ y).ravel()
        s[K + K * D:K + K * D + K * D * D] += np.einsum("nk,ni,nj->kij", p, y, y).ravel()
        s[-1] += lse.sum()

    # ---- device-resident mixture loop (CPU restatement of csrc/gmm_vb.cu) ------------------------------
    GMM_FIELDS = ["pm0", "pm1", "gpm", "pl0", "pl1", "gpl", "pa", "gpa", "ng", "lprev",
                  "mu_phi0", "mu_phi1", "mu_u0", "mu_cov", "mu_u1", "mu_g",
                  "lam_phi0", "lam_phi1", "lam_u0", "lam_u1", "lam_g", "al_phi", "al_u", "al_g",
                  "z_g", "z_h", "z_logpi", "z_t", "stats", "xstats"]

    def gmm_vb_layout(self, D, K):
        KD, KDD = K * D, K * D * D
        NS = K + KD + KDD + 1
        size = dict(pm0=KD, pm1=KDD, gpm=K, pl0=KDD, pl1=K, gpl=K, pa=K, mu_phi0=KD, mu_phi1=KDD, mu_u0=KD,
                    mu_cov=KDD, mu_u1=KDD, mu_g=K, lam_phi0=KDD, lam_phi1=K, lam_u0=KDD, lam_u1=K, lam_g=K,
                    al_phi=K, al_u=K, z_g=K, z_h=KD, z_logpi=K, stats=NS, xstats=NS)
        out, o = {}, 0
        for f in self.GMM_FIELDS:
            n = size.get(f, 1)
            out[f] = (o, n)
            o += (n + 1) & ~1        # every field starts on a 16-byte boundary, as in csrc/gmm_vb.cu
        return out, o

    def gmm_vb_set_timers(self, ids):
        pass

    def gmm_vb_run(self, Y, N, D, K, P, gz, state, ops, niter, tol, Lhist, cap, ctrl):
        """The VB.update loop of vmp.py:132-172 for the mixture model of gmm.rst:71-98, node by node:
        Z: mixture.py:53-160 + multinomial.py:101-121; mu: gaussian.py:341-446; Lambda: gaussian.py:2496-2522 +
        wishart.py:165-188; alpha: multinomial.py:83-90 + dirichlet.py:130-160; bound: expfamily.py:400-480;
        stop: vmp.py:738-747."""
        lay, total = self.gmm_vb_layout(D, K)
        st = _dense(state, (total,))
        f = {k: st[o:o + n] for k, (o, n) in lay.items()}
        c = _dense(ctrl, (4,), np.int32)
        Lh = _dense(Lhist, (max(cap, 1), 6)) if Lhist else None
        KD, KDD = K * D, K * D * D
        NS = K + KD + KDD + 1
        I = np.identity(D)
        R, S1, S2 = f["stats"][:K], f["stats"][K:K + KD].reshape(K, D), f["stats"][K + KD:K + KD + KDD].reshape(K, D, D)
        mu, mumu = f["mu_u0"].reshape(K, D), f["mu_u1"].reshape(K, D, D)
        Lam = f["lam_u0"].reshape(K, D, D)

        def spd(A):
            try:
                fac = scipy.linalg.cho_factor(A)
            except (np.linalg.LinAlgError, ValueError):
                c[2] |= 1
                return np.full((D, D), np.nan), np.nan
            return scipy.linalg.cho_solve(fac, I), 2 * np.sum(np.log(np.diag(fac[0])))

        for _ in range(niter):
            for op in ops:
                if c[1]:
                    return
                if op == 1:      # Z
                    f["z_g"][:] = -0.5 * np.einsum("kij,kij->k", mumu, Lam) + 0.5 * f["lam_u1"]
                    f["z_h"][:] = np.einsum("kij,kj->ki", Lam, mu).ravel()
                    f["z_logpi"][:] = f["al_u"]
                    f["xstats"][:] = 0.0
                    if N > 0:
                        self.gmm_sweep(Y, N, D, K, f["z_g"].ctypes.data, f["z_h"].ctypes.data, f["lam_u0"].ctypes.data,
                                       f["z_logpi"].ctypes.data, P, gz, f["xstats"].ctypes.data)
                    self.allreduce_sum_f64(f["xstats"].ctypes.data, NS)       # identity without a communicator
                    f["stats"][:] = f["xstats"]
                    f["z_t"][0] = np.sum(f["z_g"] * R) + np.sum(f["z_h"].reshape(K, D) * S1) - 0.5 * np.sum(Lam * S2)
                elif op == 2:    # mu
                    phi0 = f["pm0"].reshape(K, D) + np.einsum("kij,kj->ki", Lam, S1)
                    phi1 = f["pm1"].reshape(K, D, D) - 0.5 * R[:, None, None] * Lam
                    f["mu_phi0"][:] = phi0.ravel()
                    f["mu_phi1"][:] = phi1.ravel()
                    for k in range(K):
                        cov, ld = spd(-2 * phi1[k])
                        u0 = cov @ phi0[k]
                        f["mu_cov"].reshape(K, D, D)[k] = cov
                        mu[k] = u0
                        mumu[k] = cov + np.outer(u0, u0)
                        f["mu_g"][k] = -0.5 * u0 @ phi0[k] + 0.5 * ld
                elif op == 3:    # Lambda
                    t = S2 - S1[:, :, None] * mu[:, None, :] - mu[:, :, None] * S1[:, None, :] + mumu * R[:, None, None]
                    phi0 = f["pl0"].reshape(K, D, D) - 0.5 * t
                    nu2 = f["pl1"] + 0.5 * R
                    f["lam_phi0"][:] = phi0.ravel()
                    f["lam_phi1"][:] = nu2
                    for k in range(K):
                        inv, ld = spd(-phi0[k])
                        Lam[k] = nu2[k] * inv
                        f["lam_u1"][k] = -ld + np.sum(sp.digamma(nu2[k] - 0.5 * np.arange(D)))
                        f["lam_g"][k] = nu2[k] * ld - sp.multigammaln(nu2[k], D)
                elif op == 4:    # alpha
                    a = f["pa"] + R
                    if np.any(~(a > 0)):
                        c[2] |= 2
                    f["al_phi"][:] = a
                    with np.errstate(all="ignore"):
                        f["al_u"][:] = sp.psi(a) - sp.psi(np.sum(a))
                        f["al_g"][0] = sp.gammaln(np.sum(a)) - np.sum(sp.gammaln(a))
                elif op == 5:    # bound
                    g = -0.5 * np.einsum("kij,kij->k", mumu, Lam) + 0.5 * f["lam_u1"]
                    h = np.einsum("kij,kj->ki", Lam, mu)
                    LY = np.sum(g * R) + np.sum(h * S1) - 0.5 * np.sum(Lam * S2) - 0.5 * D * np.log(2 * np.pi) * f["ng"][0]
                    LZ = f["stats"][NS - 1] - f["z_t"][0] + np.sum((f["al_u"] - f["z_logpi"]) * R)
                    LM = np.sum(f["gpm"] - f["mu_g"]) + np.sum((f["pm0"] - f["mu_phi0"]) * f["mu_u0"]) \
                        + np.sum((f["pm1"] - f["mu_phi1"]) * f["mu_u1"])
                    LL = np.sum(f["gpl"] - f["lam_g"]) + np.sum((f["pl0"] - f["lam_phi0"]) * f["lam_u0"]) \
                        + np.sum((f["pl1"] - f["lam_phi1"]) * f["lam_u1"])
                    LA = f["gpa"][0] - f["al_g"][0] + np.sum((f["pa"] - f["al_phi"]) * f["al_u"])
                    L = LY + LZ + LM + LL + LA
                    it = int(c[0])
                    if Lh is not None and it < cap:
                        Lh[it] = [LY, LZ, LM, LL, LA, L]
                    L0 = f["lprev"][0]
                    f["lprev"][0] = L
                    c[0] = it + 1
                    if tol >= 0 and L0 == L0:
                        if (L - L0) / (0.5 * (abs(L0) + abs(L))) < tol:
                            c[1] = 1
                    if c[2]:
                        c[1] = 1
                else:
                    raise ValueError("gmm_vb_run: unknown opcode %d" % op)

    def gmm_stats(self, Y, N, D, K, P, stats):
        """p-weighted plate sums of mixture.py:108-160 + node.py:650 for given responsibilities."""
        self._launches += 1
        y = _dense(Y, (N, D))
        p = _dense(P, (N, K))
        s = _dense(stats, (K + K * D + K * D * D + 1,))
        s[:K] += p.sum(axis=0)
        s[K:K + K * D] += (p.T @ y).ravel()
        s[K + K * D:K + K * D + K * D * D] += np.einsum("nk,ni,nj->kij", p, y, y).ravel()

    # ---- block-tridiagonal SPD solver (csrc/gmc.cu) -------------------------------------------------
    def block_banded_solve(self, A, B, y, batch, T, D, V, C, x, logdet, check=True):
        """utils/linalg.py:468-575: forward block elimination storing the factor of every pivot block,
        then the backward recursion for the solution and the diagonal / super-diagonal blocks of the
        inverse (= the RTS smoother's covariances), restated with dense NumPy solves."""
        self._launches += 1
        Aa = _dense(A, (batch, T, D, D))
        Ba = _dense(B, (batch, max(T - 1, 0), D, D)) if T > 1 else np.zeros((batch, 0, D, D))
        ya = _dense(y, (batch, T, D))
        Vo = _dense(V, (batch, T, D, D))
        Co = _dense(C, (batch, max(T - 1, 0), D, D)) if T > 1 else None
        xo = _dense(x, (batch, T, D))
        ld = _dense(logdet, (batch,))
        for b in range(batch):
            Vt = np.empty((T, D, D))
            Ct = np.empty((max(T - 1, 0), D, D))
            xt = np.empty((T, D))
            piv = Aa[b, 0].copy()
            xt[0] = ya[b, 0]
            ldet = 0.0
            for n in range(T):
                try:
                    U = scipy.linalg.cho_factor(piv)[0]
                except np.linalg.LinAlgError:
                    raise NotPositiveDefinite("Matrix not positive definite")
                ldet += 2 * np.sum(np.log(np.diag(U)))
                Vt[n] = scipy.linalg.cho_solve((U, False), np.identity(D))       # inverse of the pivot block
                if n < T - 1:
                    Ct[n] = Vt[n] @ Ba[b, n]
                    xt[n + 1] = ya[b, n + 1] - Ba[b, n].T @ (Vt[n] @ xt[n])
                    piv = Aa[b, n + 1] - Ba[b, n].T @ Ct[n]
                    piv = 0.5 * (piv + piv.T)
            xt[T - 1] = Vt[T - 1] @ xt[T - 1]
            for n in range(T - 2, -1, -1):
                xt[n] = Vt[n] @ (xt[n] - Ba[b, n] @ xt[n + 1])
                Vn = Vt[n] + Ct[n] @ Vt[n + 1] @ Ct[n].T
                Ct[n] = -Ct[n] @ Vt[n + 1]
                Vt[n] = 0.5 * (Vn + Vn.T)
            Vo[b], xo[b], ld[b] = Vt, xt, ldet
            if T > 1:
                Co[b] = Ct

    # ---- device-resident VB loop of the factor model (csrc/pca_vb.cu) ------------------------------
    VB_FIELDS = ["mux", "ax", "muc", "a0", "b0", "ta0", "tb0", "sumsq", "ng", "reserved",
                 "w", "sww", "covc", "lamc", "logdetc", "phi0c", "gc",
                 "al_phi0", "al_phi1", "al_u0", "al_u1", "al_g",
                 "tau_phi0", "tau_phi1", "tau_u0", "tau_u1", "tau_g",
                 "covx", "lamx", "logdetx", "A", "bx", "stats", "sxxt", "lprev", "stats_local", "phi1x", "phi1c"]

    def pca_vb_layout(self, M, K):
        KK, MK = K * K, M * K
        NS = MK + KK + K
        size = dict(mux=K, ax=K, muc=K, a0=K, b0=K, w=MK, sww=KK, covc=KK, lamc=KK, phi0c=MK, gc=M,
                    al_phi0=K, al_phi1=K, al_u0=K, al_u1=K, al_g=K, covx=KK, lamx=KK, A=MK, bx=K,
                    stats=NS, sxxt=KK, stats_local=NS, phi1x=KK, phi1c=KK)
        out, o = {}, 0
        for f in self.VB_FIELDS:
            n = size.get(f, 1)
            out[f] = (o, n)
            o += n
        return out, o

    def pca_vb_set_timers(self, ids):
        self._vb_timers = list(ids)
        self._vb_timer_pos = 0

    def pca_vb_timers_used(self):
        return getattr(self, "_vb_timer_pos", 0)

    @staticmethod
    def _gamma(phi0, phi1):
        """gamma.py:124-148."""
        a, b = phi1, -phi0
        if np.any(a <= 0) or np.any(b <= 0):
            raise ValueError("Natural parameters should be positive")
        return a / b, sp.psi(a) - np.log(b), a * np.log(b) - sp.gammaln(a)

    def pca_vb_run(self, Y, M, N, K, X, state, ops, niter, has_alpha, has_tau, tol, Lhist, cap, ctrl):
        """The VB.update loop of vmp.py:132-172 for the PCA model of pca.rst:40-66, node by node:
        X: gaussian.py:649-706 + dot.py:581; C: same with roles swapped; alpha, tau: gamma.py:96-148 with
        the messages of gaussian.py:609-637 / :2351-2371; bound: expfamily.py:400-480; stop: vmp.py:738-747."""
        lay, total = self.pca_vb_layout(M, K)
        st = _dense(state, (total,))
        f = {k: st[o:o + n] for k, (o, n) in lay.items()}
        c = _dense(ctrl, (4,), np.int32)
        Lh = _dense(Lhist, (cap, 6))
        y = _dense(Y, (M, N))
        x = _dense(X, (N, K))
        MK, KK = M * K, K * K
        W = f["w"].reshape(M, K)
        LOG2PI = np.log(2 * np.pi)
        Ng = float(f["ng"][0])

        def E2():
            return f["sumsq"][0] - 2 * np.sum(W * f["stats"][:MK].reshape(M, K)) + np.sum(f["sww"] * f["sxxt"])

        for _ in range(niter):
            for op in ops:
                if c[1]:
                    return
                tau = f["tau_u0"][0]
                if op == 4:      # XPRE
                    lam = np.diag(f["ax"]) + tau * f["sww"].reshape(K, K)
                    cov = np.linalg.inv(lam)
                    f["lamx"][:] = lam.ravel()
                    f["phi1x"][:] = -0.5 * lam.ravel()
                    f["covx"][:] = cov.ravel()
                    f["logdetx"][0] = np.linalg.slogdet(lam)[1]
                    f["bx"][:] = cov @ (f["ax"] * f["mux"])
                    f["A"][:] = (tau * cov @ W.T).ravel()
                elif op == 1:    # XSWEEP
                    self._launches += 1
                    tp = getattr(self, "_vb_timer_pos", 0)
                    if tp < len(getattr(self, "_vb_timers", [])):
                        self._vb_timer_pos = tp + 1
                    x[...] = y.T @ f["A"].reshape(K, M).T + f["bx"]
                    self._vb_local = np.concatenate([(y @ x).ravel(), (x.T @ x).ravel(), x.sum(0)])
                elif op == 2:    # STATS (+ the sweep's one all-reduce)
                    self._launches += 1
                    v = self._vb_local
                    if self._allreduce_hook is not None:
                        f["stats_local"][:] = v
                        v = self._allreduce_hook(v.copy())
                    f["stats"][:] = v
                elif op == 3:    # SXXT
                    f["sxxt"][:] = Ng * f["covx"] + f["stats"][MK:MK + KK]
                elif op == 5:    # ROW
                    lam = np.diag(f["al_u0"]) + tau * f["sxxt"].reshape(K, K)
                    cov = np.linalg.inv(lam)
                    ld = np.linalg.slogdet(lam)[1]
                    phi0 = f["al_u0"] * f["muc"] + tau * f["stats"][:MK].reshape(M, K)
                    W[...] = phi0 @ cov.T
                    f["lamc"][:] = lam.ravel()
                    f["phi1c"][:] = -0.5 * lam.ravel()
                    f["covc"][:] = cov.ravel()
                    f["logdetc"][0] = ld
                    f["phi0c"][:] = phi0.ravel()
                    f["gc"][:] = -0.5 * np.sum(W * phi0, axis=1) + 0.5 * ld
                    f["sww"][:] = (M * cov + W.T @ W).ravel()
                elif op == 6:    # ALPHA
                    mu = f["muc"]
                    d = np.diag(f["sww"].reshape(K, K)) - 2 * mu * W.sum(0) + M * mu * mu
                    f["al_phi0"][:] = -f["b0"] - 0.5 * d
                    f["al_phi1"][:] = f["a0"] + 0.5 * M
                    f["al_u0"][:], f["al_u1"][:], f["al_g"][:] = self._gamma(f["al_phi0"], f["al_phi1"])
                elif op == 7:    # TAU
                    f["tau_phi0"][0] = -f["tb0"][0] - 0.5 * E2()
                    f["tau_phi1"][0] = f["ta0"][0] + 0.5 * M * Ng
                    f["tau_u0"][0], f["tau_u1"][0], f["tau_g"][0] = self._gamma(f["tau_phi0"][0], f["tau_phi1"][0])
                elif op == 8:    # BOUND
                    self._launches += 1
                    logtau = f["tau_u1"][0]
                    LY = -0.5 * tau * E2() + 0.5 * M * Ng * (logtau - LOG2PI)
                    Sxx = f["stats"][MK:MK + KK]
                    sx = f["stats"][MK + KK:]
                    ax, mux = f["ax"], f["mux"]
                    trLS = np.sum(f["lamx"] * Sxx)
                    dphi1 = 0.5 * f["lamx"].reshape(K, K) - 0.5 * np.diag(ax)
                    LX = (np.sum(ax * mux * sx) - trLS + np.sum(dphi1.ravel() * f["sxxt"]) + 0.5 * trLS
                          - 0.5 * Ng * f["logdetx"][0] + Ng * np.sum(-0.5 * ax * mux ** 2 + 0.5 * np.log(ax)))
                    al, lal, muc = f["al_u0"], f["al_u1"], f["muc"]
                    phi0c = f["phi0c"].reshape(M, K)
                    LC = (np.sum((al * muc - phi0c) * W)
                          + np.sum((0.5 * f["lamc"].reshape(K, K) - 0.5 * np.diag(al)).ravel() * f["sww"])
                          - np.sum(f["gc"]) + M * np.sum(-0.5 * al * muc ** 2 + 0.5 * lal))
                    LA = 0.0
                    if has_alpha:
                        a0, b0 = f["a0"], f["b0"]
                        LA = np.sum((-b0 - f["al_phi0"]) * al + (a0 - f["al_phi1"]) * lal
                                    + a0 * np.log(b0) - sp.gammaln(a0) - f["al_g"])
                    LT = 0.0
                    if has_tau:
                        a0, b0 = f["ta0"][0], f["tb0"][0]
                        LT = ((-b0 - f["tau_phi0"][0]) * tau + (a0 - f["tau_phi1"][0]) * logtau
                              + a0 * np.log(b0) - sp.gammaln(a0) - f["tau_g"][0])
                    L = LY + LX + LC + LA + LT
                    it = int(c[0])
                    if it < cap:
                        Lh[it] = [LY, LX, LC, LA, LT, L]
                    L0 = f["lprev"][0]
                    f["lprev"][0] = L
                    c[0] = it + 1
                    if tol >= 0 and not np.isnan(L0):
                        if (L - L0) / (0.5 * (abs(L0) + abs(L))) < tol:
                            c[1] = 1
                else:
                    raise ValueError("unknown opcode %d" % op)
