"""Gaussian-family nodes on device: GaussianARD and Gaussian.

Replaces the array math of ``bayespy/inference/vmp/nodes/gaussian.py``:
GaussianARDDistribution :576-889, GaussianDistribution :293-573, and the glue
nodes GaussianToGaussianGamma :2226, WrapToGaussianGamma :2299,
WrapToGaussianWishart :2374, whose only job in the reference is to multiply
Gaussian messages/moments by <alpha> or <Lambda> and to split a joint message
between the mean and precision parents.  Here a Gaussian node simply has two
parents (mean, precision) and computes those products itself, so the messages
that arrive at the user-visible parent nodes are identical to the reference's.

The per-plate K x K work (Cholesky, inverse, solve, log-det; linalg.py:31-223
called from gaussian.py:672-706 / :397-446) is one fused kernel,
``bpk_gaussian_moments``.  The second moment <x x^T> = Cov + <x><x>^T is held
in factored form (Cov may be shared by all plates) and only materialised when
somebody asks for ``u[1]``.
"""
import numpy as np

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .expfam import Distribution, ExponentialFamily
from .node import Constant, Deterministic, Node, broadcast_plates, is_subshape

LOG2PI = float(np.log(2 * np.pi))


# --------------------------------------------------------------------------------------------
# lazy second moment
# --------------------------------------------------------------------------------------------
class FactoredSecondMoment:
    """<x x^T> = cov + u0 u0^T, with cov possibly broadcast over plates."""

    def __init__(self, u0, cov, shape):
        self.u0 = u0                  # plates_b + (K,)
        self.cov = cov                # plates_c + (K, K)
        self.var_shape = tuple(shape)
        self._dense = None

    @property
    def shape(self):
        pl = np.broadcast_shapes(self.u0.shape[:-1], self.cov.shape[:-2])
        return tuple(pl) + self.var_shape + self.var_shape

    def materialize(self):
        if self._dense is None:
            K = self.u0.shape[-1]
            pl = tuple(np.broadcast_shapes(self.u0.shape[:-1], self.cov.shape[:-2]))
            N = int(np.prod(pl, dtype=np.int64)) if pl else 1
            u0 = self.u0 if self.u0.shape[:-1] == pl else self.u0.broadcast_to(pl + (K,))
            u0 = u0.contiguous()
            ncov = int(np.prod(self.cov.shape[:-2], dtype=np.int64)) if self.cov.ndim > 2 else 1
            cov = dense(self.cov)             # the covariance itself may be produced on demand
            if ncov != 1 and tuple(cov.shape[:-2]) != pl:
                cov = cov.broadcast_to(pl + (K, K))
                ncov = N
            cov = cov.contiguous()
            out = DArray.empty(pl + (K, K))
            _bpk.get().outer_add(u0.ptr, cov.ptr, ncov, N, K, out.ptr)
            self._dense = out.reshape(pl + self.var_shape + self.var_shape)
        return self._dense

    def numpy(self):
        return self.materialize().numpy()

    def __array__(self, dtype=None, copy=None):
        return self.numpy()

    # array protocol of the dense form, for callers that treat the moment as an array
    @property
    def T(self): return self.materialize().T
    @property
    def ndim(self): return len(self.shape)
    def __getitem__(self, index): return self.materialize()[index]
    def __add__(self, o): return self.materialize() + dense(o)
    def __radd__(self, o): return dense(o) + self.materialize()
    def __sub__(self, o): return self.materialize() - dense(o)
    def __rsub__(self, o): return dense(o) - self.materialize()
    def __mul__(self, o): return self.materialize() * dense(o)
    def __rmul__(self, o): return dense(o) * self.materialize()
    def __neg__(self): return -self.materialize()


def dense(x):
    return x.materialize() if hasattr(x, "materialize") else x


# --------------------------------------------------------------------------------------------
# parent coercion
# --------------------------------------------------------------------------------------------
def gaussian_constant(x, ndim):
    """Constant Gaussian moments [x, x x^T] of a fixed array (GaussianMoments.compute_fixed_moments,
    gaussian.py:42-100)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim < ndim:
        raise ValueError("Array has fewer axes than ndim=%d" % ndim)
    shape = x.shape[x.ndim - ndim:] if ndim > 0 else ()
    plates = x.shape[:x.ndim - ndim]
    xd = D.asarray(x)
    if ndim == 0:
        xx = D.square(xd)
    else:
        xx = D.mul(xd.add_trailing(ndim), xd.reshape(plates + (1,) * ndim + shape))
    return Constant("gaussian", [xd, xx], dims=(shape, shape + shape), plates=plates, value=x)


def gamma_constant(a):
    """[a, log a] (GammaMoments.compute_fixed_moments, gamma.py:62-87)."""
    a = np.asarray(a, dtype=np.float64)
    if np.any(a < 0):
        raise ValueError("Values must be positive")
    ad = D.asarray(a)
    with np.errstate(divide="ignore"):
        return Constant("gamma", [ad, D.log(ad)], dims=((), ()), plates=a.shape, value=a)


def ensure_gaussian(x, ndim):
    if isinstance(x, Node) and hasattr(x, "_to_gaussian"):
        x = x._to_gaussian()          # e.g. a Gaussian Markov chain seen as Gaussians plated over time
    if isinstance(x, Node):
        if x.moment_kind != "gaussian":
            raise ValueError("Expected a Gaussian-like node, got %s" % type(x).__name__)
        return x
    return gaussian_constant(x, ndim)


def ensure_gamma(a):
    if isinstance(a, Node):
        if a.moment_kind != "gamma":
            raise ValueError("Expected a gamma-like node, got %s" % type(a).__name__)
        return a
    return gamma_constant(a)


class GaussianDimsToPlates(Deterministic):
    """View a Gaussian node with variable shape S as scalar Gaussians plated over S
    (the reference does this with a moments converter when a vector Gaussian is the
    mean of a GaussianARD, gaussian.py:1621).  Only the diagonal of <x x^T> is used."""
    moment_kind = "gaussian"

    def __init__(self, X, name=""):
        self.shape_in = tuple(X.dims[0])
        self.nd = len(self.shape_in)
        super().__init__(X, dims=((), ()), plates=tuple(X.plates) + self.shape_in, name=name or X.name)

    def _plates_from_parent(self, index):
        return tuple(self.parents[0].plates) + tuple(self.parents[0].dims[0])

    def _plates_to_parent(self, index):
        return tuple(self.plates[:len(self.plates) - self.nd])

    def _weights_to_parent(self, index, mask):
        mask = np.asarray(mask)
        if mask.ndim <= self.nd:
            return np.any(mask)
        return np.any(mask, axis=tuple(range(mask.ndim - self.nd, mask.ndim)))

    def _compute_moments(self, u):
        return [u[0], dense(u[1]).diag_view(self.nd)]

    def _compute_message_to_parent(self, index, m, u):
        m0, m1 = m
        out1 = None
        if m1 is not None:
            pl = m1.shape[:m1.ndim - self.nd] if m1.ndim >= self.nd else ()
            m1b = m1.broadcast_to(tuple(pl) + self.shape_in) if m1.ndim >= self.nd else \
                m1.broadcast_to(self.shape_in)
            out1 = DArray.zeros(tuple(m1b.shape) + self.shape_in)
            D.copy_into(out1.diag_view(self.nd), m1b)
        if m0 is not None and m0.ndim >= self.nd:
            m0 = m0.broadcast_to(tuple(m0.shape[:m0.ndim - self.nd]) + self.shape_in)
        elif m0 is not None:
            m0 = m0.broadcast_to(self.shape_in)
        return [m0, out1]

    def message_to_parent(self, index):
        # the dims of this node are plates here but variable axes in the parent:
        # reduce over true plates only.
        m, mask = self._message_and_mask_to_parent(index)
        parent = self.parents[0]
        mdev = self.mask_device(mask)
        out = []
        for i, mi in enumerate(m):
            if mi is None:
                out.append(None)
                continue
            nd = len(parent.dims[i])
            frm = tuple(self._plates_to_parent(0)) + tuple(parent.dims[i])
            to = tuple(parent.plates) + tuple(parent.dims[i])
            mk = mdev.add_trailing(nd) if mdev is not None else None
            out.append(D.reduce_to_shape(mi, to, mask=mk, from_shape=frm))
        return out


# --------------------------------------------------------------------------------------------
# shared device routine: moments of N(phi) for flattened K
# --------------------------------------------------------------------------------------------
def _flat_count(shape):
    return int(np.prod(shape, dtype=np.int64)) if len(shape) else 1


def _align_batch(a, P, tail):
    """Return (contiguous array, count) with count in {1, prod(P)} for an operand whose
    plate part broadcasts to P."""
    Pa = tuple(a.shape[:a.ndim - len(tail)])
    N = _flat_count(P)
    if _flat_count(Pa) == 1:
        return a.contiguous(), 1
    if (1,) * (len(P) - len(Pa)) + Pa != tuple(P):
        a = a.broadcast_to(tuple(P) + tuple(tail))
    return a.contiguous(), N


def gaussian_moments_device(phi0, phi1, K):
    """phi0: P0+(K,), phi1: P1+(K,K) with P0, P1 mutually broadcastable.
    Returns u0 (P+(K,)), cov (P or ones+(K,K)), g (P), logdet."""
    be = _bpk.get()
    P = tuple(np.broadcast_shapes(tuple(phi0.shape[:-1]), tuple(phi1.shape[:-2])))
    N = _flat_count(P)
    phi0c, n0 = _align_batch(phi0, P, (K,))
    phi1c, n1 = _align_batch(phi1, P, (K, K))
    if N == 1:
        n0 = n1 = 1
    covP = P if n1 == N else (1,) * len(P)
    u0 = DArray.empty(P + (K,))
    cov = DArray.empty(tuple(covP) + (K, K))
    g = DArray.empty(P)
    logdet = DArray.empty(covP)
    be.gaussian_moments(phi0c.ptr, n0, phi1c.ptr, n1, N, K, u0.ptr, cov.ptr, g.ptr, logdet.ptr, True)
    return u0, cov, g, logdet


class _GaussianNode(ExponentialFamily):
    moment_kind = "gaussian"

    def get_moments(self):
        if isinstance(self.u[1], FactoredSecondMoment):
            self.u[1] = self.u[1].materialize()
        return list(self.u)

    def cov_form(self):
        """(u0, cov, var_shape) with cov possibly shared across plates, or None if unknown."""
        if isinstance(self.u[1], FactoredSecondMoment):
            return self.u[1].u0, self.u[1].cov
        return None

    def _store(self, u, g, update_mask):
        from .node import mask_is_full
        if not (mask_is_full(update_mask) or self.u[0] is None):
            u = [u[0], dense(u[1])]
            if isinstance(self.u[1], FactoredSecondMoment):
                self.u[1] = self.u[1].materialize()
        if mask_is_full(update_mask) or self.u[0] is None:
            self.u = [u[0], u[1]]
            if g is not None:
                self.g = g
            self._version += 1
        else:
            super()._store(u, g, update_mask)


# --------------------------------------------------------------------------------------------
# GaussianARD
# --------------------------------------------------------------------------------------------

def gaussian_fisher_apply(g, u, K):
    """Euclidean gradient of a Gaussian's bound with respect to its natural parameters from the Riemannian gradient
    ``g`` = [g0 (.., K), g1 (.., K, K)] (gaussian.py:489-555, :824-890): the Fisher information d<u>/dphi applied to g.
    With m = <x>, S = Cov(x), M2 = <x x^T>:
        d0 = S g0 + 2 S g1 m
        d1 = (S g0) m^T + m (S g0)^T + 2 M2 g1 M2 - 2 (m^T g1 m) m m^T
    Device contractions over the flattened plates; K is the flattened variable size."""
    x = D.asarray(u[0])
    lead = tuple(x.shape[:x.ndim - 1]) if K > 1 or x.ndim >= 1 else ()
    xf = x.reshape((-1, K))
    n = xf.shape[0]
    xx = D.asarray(dense(u[1])).reshape((-1, K, K)).broadcast_to((n, K, K))
    g0 = D.asarray(g[0]).reshape((-1, K)).broadcast_to((n, K))
    g1 = D.asarray(g[1]).reshape((-1, K, K)).broadcast_to((n, K, K))
    xxT = D.mul(xf.reshape((n, K, 1)), xf.reshape((n, 1, K)))
    cov = D.sub(xx, xxT)
    cg0 = D.sum_product([cov, g0], [["n", "i", "j"], ["n", "j"]], ["n", "i"])
    g1x = D.sum_product([g1, xf], [["n", "i", "j"], ["n", "j"]], ["n", "i"])
    d0 = D.axpby(1.0, cg0, 2.0, D.sum_product([cov, g1x], [["n", "i", "j"], ["n", "j"]], ["n", "i"]))
    cg0x = D.mul(cg0.reshape((n, K, 1)), xf.reshape((n, 1, K)))
    xcg0 = D.mul(xf.reshape((n, K, 1)), cg0.reshape((n, 1, K)))
    t = D.sum_product([xx, g1], [["n", "i", "k"], ["n", "k", "l"]], ["n", "i", "l"])
    mid = D.sum_product([t, xx], [["n", "i", "l"], ["n", "l", "j"]], ["n", "i", "j"])
    q = D.sum_product([g1x, xf], [["n", "i"], ["n", "i"]], ["n"])
    d1 = D.add(D.add(cg0x, xcg0), D.axpby(2.0, mid, -2.0, D.mul(xxT, q.reshape((n, 1, 1)))))
    return d0, d1


class GaussianARDDistribution(Distribution):
    """x ~ N(mu, diag(alpha)^-1) over a variable block of shape ``shape``
    (gaussian.py:576-889).  Parents: mu (scalar-Gaussian moments plated over
    plates+shape) and alpha (gamma moments plated over plates+shape)."""

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.ndim = len(self.shape)
        self.K = _flat_count(self.shape)

    def compute_gradient(self, g, u, phi):
        """gaussian.py:824-890."""
        d0, d1 = gaussian_fisher_apply(g, u, self.K)
        return [d0.reshape(D.asarray(g[0]).shape), d1.reshape(D.asarray(g[1]).shape)]

    # -- plates: the variable axes are plate axes of both parents (gaussian.py:744-769)
    def plates_to_parent(self, index, plates):
        return tuple(plates) + self.shape

    def plates_from_parent(self, index, plates):
        return tuple(plates[:max(len(plates) - self.ndim, 0)]) if self.ndim else tuple(plates)

    def compute_weights_to_parent(self, index, weights):
        w = np.asarray(weights)
        return w.reshape(w.shape + (1,) * self.ndim)

    def _expand_var_axes(self, a):
        nd = self.ndim
        if a.ndim < nd:
            a = a.add_leading(nd - a.ndim)
        return a.broadcast_to(tuple(a.shape[:a.ndim - nd]) + self.shape)

    # -- natural parameters (gaussian.py:649-670)
    def compute_phi_from_parents(self, u_mu, u_alpha, mask=True):
        alpha = u_alpha[0]
        phi0 = D.mul(alpha, u_mu[0])
        if self.ndim == 0:
            return [phi0, D.mul(alpha, -0.5)]
        # no broadcasting over the variable axes (gaussian.py:661-666)
        phi0 = self._expand_var_axes(phi0)
        a = self._expand_var_axes(alpha)
        phi1 = DArray.zeros(tuple(a.shape) + self.shape)
        D._ew("AFFINE", a.shape, phi1.diag_view(self.ndim), [a], alpha=-0.5, beta=0.0)
        return [phi0, phi1]

    # -- E[g(parents)] (gaussian.py:709-732)
    def compute_cgf_from_parents(self, u_mu, u_alpha):
        alpha, logalpha = u_alpha
        mu2 = u_mu[1]
        t = D.mul(alpha, mu2)                       # alpha * <mu^2>, plates+shape (broadcastable)
        if self.ndim == 0:
            return D.axpby(-0.5, t, 0.5, logalpha)
        full = self.shape
        nd = self.ndim

        def sum_dims(a):
            # sum over the variable axes, counting broadcast axes with their full extent
            a = a if a.ndim >= nd else a.add_leading(nd - a.ndim)
            npl = a.ndim - nd
            return D.reduce_to_shape(a, tuple(a.shape[:npl]) + (1,) * nd,
                                     from_shape=tuple(a.shape[:npl]) + full).reshape(a.shape[:npl])
        return D.axpby(-0.5, sum_dims(t), 0.5, sum_dims(logalpha))

    # -- moments (gaussian.py:672-706)
    def compute_moments_and_cgf(self, phi, mask=True):
        phi = [D.asarray(dense(v)) for v in phi]
        if self.ndim == 0:
            # scalar closed form :673-678
            u0 = D.div(D.mul(phi[0], -0.5), phi[1])
            var = D._unary("RECIP", phi[1], -0.5)
            u1 = D.add(D.square(u0), var)
            g = D.axpby(-0.5, D.mul(u0, phi[0]), 0.5, D.log(D.mul(phi[1], -2.0)))
            return [u0, u1], g
        K = self.K
        p0 = phi[0].reshape(tuple(phi[0].shape[:phi[0].ndim - self.ndim]) + (K,))
        p1 = phi[1].reshape(tuple(phi[1].shape[:phi[1].ndim - 2 * self.ndim]) + (K, K))
        u0, cov, g, _ = gaussian_moments_device(p0, p1, K)
        u0s = u0.reshape(tuple(u0.shape[:-1]) + self.shape)
        return [u0s, FactoredSecondMoment(u0, cov, self.shape)], g

    # -- messages (gaussian.py:609-637 combined with :2351-2371)
    def compute_message_to_parent(self, parent, index, u, u_mu, u_alpha):
        x = u[0]
        if index == 0:
            alpha = u_alpha[0]
            return [D.mul(alpha, x), D.mul(alpha, -0.5)]
        elif index == 1:
            x2 = dense(u[1]).diag_view(self.ndim)
            mu, mu2 = u_mu
            # -1/2 (<x^2> - 2 <x><mu> + <mu^2>)
            t = D.fma(-2.0, x, mu, 1.0, x2)
            t = D.add(t, mu2)
            # 0.5 * ones(shape): explicit over the variable axes like gaussian.py:633
            half = DArray.full(self.shape, 0.5) if self.ndim else D.asarray(0.5)
            return [D.mul(t, -0.5), half]
        raise ValueError("Invalid parent index")

    def compute_fixed_moments_and_f(self, x, mask=True):
        x = np.asarray(x, dtype=np.float64)
        if self.ndim > 0 and x.shape[-self.ndim:] != self.shape:
            raise ValueError("Invalid shape")
        xd = D.asarray(x)
        # the second moment of observed data is produced on first use only: the fused
        # sweeps never read it, and at N=1e7 it would be another 5 GB of HBM writes
        from .plans import LazyArray
        if self.ndim == 0:
            xx = LazyArray(xd.shape, lambda: D.square(xd))
        else:
            pl = x.shape[:x.ndim - self.ndim]
            nd, shp = self.ndim, self.shape
            xx = LazyArray(tuple(pl) + shp + shp,
                           lambda: D.mul(xd.add_trailing(nd), xd.reshape(pl + (1,) * nd + shp)))
        return [xd, xx], -0.5 * self.K * LOG2PI

    def random(self, *phi, plates=None):
        """Host draw with NumPy's global RNG, same call pattern as gaussian.py:772-821."""
        import scipy.linalg
        D_ = self.ndim
        dims = self.shape
        plates = tuple(plates)
        if self.K == 1:
            phi1 = np.reshape(phi[1], np.shape(phi[1])[:np.ndim(phi[1]) - 2 * D_] + D_ * (1,)) if D_ > 0 else phi[1]
            var = -0.5 / phi1
            z = np.random.randn(*(plates + dims))
            return var * phi[0] + np.sqrt(var) * z
        N = self.K
        plates_cov = np.shape(phi[1])[:np.ndim(phi[1]) - 2 * D_]
        V = -2 * np.reshape(phi[1], plates_cov + (N, N))
        phi0 = np.reshape(phi[0], np.shape(phi[0])[:np.ndim(phi[0]) - D_] + (N,))
        z = np.random.randn(*(plates + (N,)))
        Vb = np.broadcast_to(V, plates + (N, N))
        pb = np.broadcast_to(phi0, plates + (N,))
        x = np.empty(plates + (N,))
        for idx in np.ndindex(*plates):
            U = scipy.linalg.cho_factor(Vb[idx])[0]
            mu = scipy.linalg.cho_solve((U, False), pb[idx])
            x[idx] = mu + scipy.linalg.solve_triangular(U, z[idx], trans="N", lower=False)
        return np.reshape(x, plates + dims)


class GaussianARD(_GaussianNode):
    """``GaussianARD(mu, alpha, ndim=None, shape=None, plates=None, name="")`` —
    same call signature and semantics as the reference node (gaussian.py:1559-1660)."""

    def __init__(self, mu, alpha, ndim=None, shape=None, plates=None, name="", initialize=True, plates_multiplier=None):
        alpha = ensure_gamma(alpha)
        scaled_mean = isinstance(mu, Node) and mu.moment_kind == "gaussian_gamma"
        if scaled_mean and len(mu.dims[0]) != 0:
            # the reference's converter has no ndim change for Gaussian-gamma moments either (gaussian.py:218-225)
            raise NotImplementedError("A Gaussian-gamma mean of GaussianARD must be scalar (ndim=0) over its plates")
        if isinstance(mu, Node):
            if mu.moment_kind not in ("gaussian", "gaussian_gamma"):
                raise ValueError("mu must be a Gaussian-like node")
            mu_nd = len(mu.dims[0])
            mu_full = tuple(mu.plates) + tuple(mu.dims[0])
        else:
            mu_nd = 0
            mu_full = np.shape(mu)
        full = broadcast_plates(mu_full, alpha.plates)
        if ndim is None:
            ndim = len(shape) if shape is not None else 0
            shape = tuple(shape) if shape is not None else ()
        elif shape is not None:
            if ndim != len(shape):
                raise ValueError("Given shape and ndim inconsistent")
            shape = tuple(shape)
        else:
            if ndim > len(full):
                raise ValueError("Cannot determine shape for ndim={0} because parent full shape has "
                                 "ndim={1}.".format(ndim, len(full)))
            shape = tuple(full[len(full) - ndim:]) if ndim > 0 else ()
        # the mean parent is consumed as scalars plated over plates+shape
        if isinstance(mu, Node):
            mu = GaussianDimsToPlates(mu) if mu_nd > 0 else mu
        else:
            mu = gaussian_constant(mu, 0)
        if scaled_mean:
            from .gaussian_gamma import GaussianARDScaledMeanDistribution
            dist = GaussianARDScaledMeanDistribution(shape)
        else:
            dist = GaussianARDDistribution(shape)
        super().__init__(mu, alpha, dims=(shape, shape + shape), distribution=dist, plates=plates,
                         name=name, initialize=initialize, plates_multiplier=plates_multiplier)

    def rotate(self, R, inv=None, logdet=None, axis=-1, Q=None, subset=None):
        """Transform q(x) -> q(R x) along the variable axis (gaussian.py:1693-1745): natural parameters by
        R^-T, moments by R, log-normaliser by -log|det R|.  R is a host K x K matrix; the plated arrays are
        transformed on the device (one pass over u0; the covariance stays factored when it is shared)."""
        if Q is not None or subset is not None:
            raise NotImplementedError()
        if len(self.dims[0]) != 1:
            return self._rotate_axis(R, inv, logdet, axis)
        if axis not in (-1, 0):
            raise ValueError("Axis out of bounds")
        from .plans import LazyArray
        R = np.asarray(R, dtype=np.float64)
        invR = np.linalg.inv(R) if inv is None else np.asarray(inv, dtype=np.float64)
        logdetR = np.linalg.slogdet(R)[1] if logdet is None else float(logdet)
        K = self.dims[0][0]
        Rd, iRT = D.asarray(R), D.asarray(np.ascontiguousarray(invR.T))

        def rot_vec(a, Mx):         # a[..., i] <- sum_k Mx[i, k] a[..., k]
            a = D.asarray(a)
            flat = a.reshape((-1, K))
            return D.sum_product([Mx, flat], [["i", "k"], ["n", "k"]], ["n", "i"]).reshape(a.shape)

        def rot_mat(a, Mx):         # a[..., i, j] <- sum_kl Mx[i, k] a[..., k, l] Mx[j, l]
            a = D.asarray(a)
            flat = a.reshape((-1, K, K))
            return D.sum_product([Mx, flat, Mx], [["i", "k"], ["n", "k", "l"], ["j", "l"]], ["n", "i", "j"]).reshape(a.shape)

        u0 = rot_vec(self.u[0], Rd)
        if isinstance(self.u[1], FactoredSecondMoment):
            u1 = FactoredSecondMoment(u0.reshape(self.u[1].u0.shape), rot_mat(self.u[1].cov, Rd), self.u[1].var_shape)
        else:
            u1 = rot_mat(self.u[1], Rd)
        phi1 = rot_mat(self.phi[1], iRT)
        if isinstance(self.phi[0], LazyArray):
            # fused sweeps keep phi0 = Lam x and g virtual: rebuild them from the rotated pieces
            fz = getattr(self, "_fused", None)
            Lam = D.mul(phi1.reshape((K, K)), -2.0)
            X = u0
            old_g = self.g
            logdet_q = fz["logdet"] if fz is not None else None
            shape0, shapeg = self.phi[0].shape, self.g.shape if isinstance(self.g, LazyArray) else None

            def phi0_fn():
                return D.sum_product([Lam, X.reshape((-1, K))], [["i", "j"], ["n", "j"]], ["n", "i"]).reshape(shape0)
            phi0 = LazyArray(shape0, phi0_fn)
            if isinstance(old_g, LazyArray) and logdet_q is not None:
                new_logdet = D.affine(logdet_q, 1.0, -2.0 * logdetR)      # log|Lam'| = log|Lam| - 2 log|det R|

                def g_fn():
                    xf = X.reshape((-1, K))
                    q = D.sum_product([xf, Lam, xf], [["n", "i"], ["i", "j"], ["n", "j"]], ["n"]).reshape(shapeg)
                    return D.axpby(-0.5, q, 0.5, new_logdet)
                g = LazyArray(shapeg, g_fn)
                self._fused = dict(fz, Lam=Lam, logdet=new_logdet, cov=u1.cov if isinstance(u1, FactoredSecondMoment) else None)
            else:
                g = D.affine(D.asarray(old_g.materialize() if isinstance(old_g, LazyArray) else old_g), 1.0, -logdetR)
        else:
            phi0 = rot_vec(self.phi[0], iRT)
            g = D.affine(D.asarray(self.g), 1.0, -logdetR)
        self.phi = [phi0, phi1]
        self.u = [u0, u1]
        self.g = g
        old_version = self._version
        self._version += 1
        for hook in getattr(self, "_rotate_hooks", ()):
            hook(Rd, old_version)          # e.g. a plan's cached plate sums rotate with the node

    def _rotate_axis(self, R, inv, logdet, axis):
        """Rotation along one axis of a variable block with several axes (gaussian.py:1693-1730 with rotate_mean /
        rotate_covariance :2612-2666): every array is viewed as (rest, K, rest[, K, rest]) and contracted with R on the
        device; the log-normaliser moves by -log|det R| times the number of rotated fibres."""
        nd = len(self.dims[0])
        if nd == 0 or not -nd <= axis < nd:
            raise ValueError("Axis out of bounds")
        axis = axis % nd - nd                                         # counted from the end
        R = np.asarray(R, dtype=np.float64)
        invR = np.linalg.inv(R) if inv is None else np.asarray(inv, dtype=np.float64)
        logdetR = np.linalg.slogdet(R)[1] if logdet is None else float(logdet)
        K = self.dims[0][axis]
        if R.shape != (K, K):
            raise ValueError("The rotation matrix must be %d x %d" % (K, K))

        def n_(shape):
            return int(np.prod(shape, dtype=np.int64)) if len(shape) else 1

        def rot_mean(a, Mx):
            a = D.asarray(dense(a)).contiguous()
            p = a.ndim + axis
            v = a.reshape((n_(a.shape[:p]), K, n_(a.shape[p + 1:])))
            return D.sum_product([Mx, v], [["i", "k"], ["n", "k", "m"]], ["n", "i", "m"]).reshape(a.shape)

        def rot_cov(a, Mx):
            a = D.asarray(dense(a)).contiguous()
            p2 = a.ndim + axis
            p1 = p2 - nd
            v = a.reshape((n_(a.shape[:p1]), K, n_(a.shape[p1 + 1:p2]), K, n_(a.shape[p2 + 1:])))
            return D.sum_product([Mx, v, Mx], [["i", "k"], ["n", "k", "m", "l", "q"], ["j", "l"]],
                                 ["n", "i", "m", "j", "q"]).reshape(a.shape)
        Rd, iRT = D.asarray(R), D.asarray(np.ascontiguousarray(invR.T))
        g_old = self.g.materialize() if hasattr(self.g, "materialize") else self.g
        self._fused = None
        self.phi = [rot_mean(self.phi[0], iRT), rot_cov(self.phi[1], iRT)]
        self.u = [rot_mean(self.u[0], Rd), rot_cov(self.u[1], Rd)]
        fibres = n_(self.dims[0]) // K
        self.g = D.affine(D.asarray(g_old), 1.0, -logdetR * fibres)

    def rotate_plates(self, Q, plate_axis=-1):
        """Approximate rotation of a plate axis (gaussian.py:1743-1774): the means mix exactly, <x_i> <- sum_k Q_ik <x_k>;
        the precision of plate p is scaled by (sum_i Q_ip)^-2 instead of being mixed; then the moments and the
        log-normaliser are recomputed from the natural parameters."""
        if not isinstance(plate_axis, int):
            raise ValueError("Plate axis must be integer")
        if plate_axis >= 0:
            plate_axis -= len(self.plates)
        if plate_axis < -len(self.plates) or plate_axis >= 0:
            raise ValueError("Axis out of bounds")
        Q = np.asarray(Q, dtype=np.float64)
        shape = tuple(self.dims[0])
        P, K = self.plates[plate_axis], _flat_count(shape)        # the variable block counts as one flattened axis
        if Q.shape != (P, P):
            raise ValueError("Q must be a square matrix over the rotated plate axis")
        npl = len(self.plates)
        ax = npl + plate_axis                                   # position of the rotated axis among the plates
        pk = [("p", j) for j in range(npl)]
        u0 = D.asarray(self.u[0]).broadcast_to(tuple(self.plates) + shape).contiguous().reshape(tuple(self.plates) + (K,))
        keys_in = pk[:ax] + ["k"] + pk[ax + 1:] + ["d"]
        keys_out = pk[:ax] + ["i"] + pk[ax + 1:] + ["d"]
        u0 = D.sum_product([D.asarray(Q), u0], [["i", "k"], keys_in], keys_out)
        s = np.sum(Q, axis=0)
        scale = D.asarray((s ** -2.0).reshape((P,) + (1,) * (npl - ax - 1) + (1, 1)))
        phi1 = D.asarray(dense(self.phi[1])).broadcast_to(tuple(self.plates) + shape + shape).contiguous()
        phi1 = D.mul(phi1.reshape(tuple(self.plates) + (K, K)), scale)
        phi0 = D.sum_product([phi1, u0], [pk + ["a", "b"], pk + ["b"]], pk + ["a"], scale=-2.0)
        self.phi = [phi0.reshape(tuple(self.plates) + shape), phi1.reshape(tuple(self.plates) + shape + shape)]
        u, g = self._distribution.compute_moments_and_cgf(self.phi)
        self._fused = None
        self._store(u, g, np.logical_not(self.observed))

    def initialize_from_parameters(self, mu, alpha):
        mu = np.asarray(mu, dtype=np.float64) * np.ones(np.shape(alpha))
        alpha = np.asarray(alpha, dtype=np.float64) * np.ones(np.shape(mu))
        u_mu = gaussian_constant(mu, 0).get_moments()
        u_al = gamma_constant(alpha).get_moments()
        self.phi = self._canonical_phi(self._distribution.compute_phi_from_parents(u_mu, u_al))
        u, g = self._distribution.compute_moments_and_cgf(self.phi)
        self._store(u, g, np.logical_not(self.observed))

    def initialize_from_mean_and_covariance(self, mu, Cov):
        nd = len(self.dims[0])
        mu = np.asarray(mu, dtype=np.float64)
        Cov = np.asarray(Cov, dtype=np.float64)
        outer = mu.reshape(mu.shape + (1,) * nd) * mu.reshape(mu.shape[:mu.ndim - nd] + (1,) * nd + mu.shape[mu.ndim - nd:])
        self._store([D.asarray(mu), D.asarray(Cov + outer)], None, np.logical_not(self.observed))
        self.g = D.asarray(np.nan)


# --------------------------------------------------------------------------------------------
# Gaussian (full precision matrix)
# --------------------------------------------------------------------------------------------
class GaussianDistribution(Distribution):
    """x ~ N(mu, Lambda^-1), x in R^D (gaussian.py:293-573).  Parents: mu (Gaussian moments,
    ndim 1) and Lambda (Wishart moments).  The reference routes both through
    WrapToGaussianWishart (gaussian.py:2374-2527); the products with <Lambda> computed there
    are computed here, so the messages reaching mu and Lambda are the same."""

    def __init__(self, D_):
        self.D = int(D_)
        self.shape = (self.D,)
        self.ndim = 1

    def compute_gradient(self, g, u, phi):
        """gaussian.py:489-555."""
        d0, d1 = gaussian_fisher_apply(g, u, self.D)
        return [d0.reshape(D.asarray(g[0]).shape), d1.reshape(D.asarray(g[1]).shape)]

    def compute_phi_from_parents(self, u_mu, u_Lambda, mask=True):
        """[<Lambda><mu>, -1/2 <Lambda>]  (gaussian.py:377-394 with :2438-2457)."""
        Lam = u_Lambda[0]
        mu = u_mu[0]
        npl = max(Lam.ndim - 2, mu.ndim - 1)
        pk = [("p", j) for j in range(npl, 0, -1)]
        phi0 = D.sum_product([Lam, mu], [pk[npl - (Lam.ndim - 2):] + ["i", "j"], pk[npl - (mu.ndim - 1):] + ["j"]],
                             pk + ["i"])
        return [phi0, D.mul(Lam, -0.5)]

    def compute_cgf_from_parents(self, u_mu, u_Lambda):
        """-1/2 tr(<mu mu^T><Lambda>) + 1/2 <log|Lambda|>  (gaussian.py:449-463)."""
        Lam, logdet = u_Lambda
        mumu = dense(u_mu[1])
        npl = max(Lam.ndim - 2, mumu.ndim - 2)
        pk = [("p", j) for j in range(npl, 0, -1)]
        t = D.sum_product([Lam, mumu], [pk[npl - (Lam.ndim - 2):] + ["i", "j"],
                                        pk[npl - (mumu.ndim - 2):] + ["i", "j"]], pk)
        return D.axpby(-0.5, t, 0.5, logdet)

    def compute_moments_and_cgf(self, phi, mask=True):
        """gaussian.py:397-446 (no truncation)."""
        phi = [D.asarray(dense(v)) for v in phi]
        u0, cov, g, _ = gaussian_moments_device(phi[0], phi[1], self.D)
        return [u0, FactoredSecondMoment(u0, cov, self.shape)], g

    def compute_message_to_parent(self, parent, index, u, u_mu, u_Lambda):
        """To mu: [<Lambda> x, -1/2 <Lambda>]; to Lambda: [-1/2 (xx^T - x mu^T - mu x^T + mu mu^T), 1/2]
        (gaussian.py:341-375 with :2496-2522)."""
        x = u[0]
        if index == 0:
            Lam = u_Lambda[0]
            npl = max(Lam.ndim - 2, x.ndim - 1)
            pk = [("p", j) for j in range(npl, 0, -1)]
            m0 = D.sum_product([Lam, x], [pk[npl - (Lam.ndim - 2):] + ["i", "j"], pk[npl - (x.ndim - 1):] + ["j"]],
                               pk + ["i"])
            return [m0, D.mul(Lam, -0.5)]
        elif index == 1:
            xx = dense(u[1])
            mu, mumu = u_mu[0], dense(u_mu[1])
            xm = D.mul(x.add_trailing(1), mu.expand_dims(-2))          # x mu^T
            mx = D.mul(mu.add_trailing(1), x.expand_dims(-2))          # mu x^T
            t = D.add(D.sub(D.sub(xx, xm), mx), mumu)
            return [D.mul(t, -0.5), D.asarray(0.5)]
        raise ValueError("Index out of bounds")

    def compute_fixed_moments_and_f(self, x, mask=True):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim < 1 or x.shape[-1] != self.D:
            raise ValueError("Invalid shape")
        xd = D.asarray(x)
        from .plans import LazyArray
        xx = LazyArray(tuple(x.shape) + (self.D,), lambda: D.mul(xd.add_trailing(1), xd.expand_dims(-2)))
        return [xd, xx], -0.5 * self.D * LOG2PI

    def random(self, *phi, plates=None):
        import scipy.linalg
        N = self.D
        plates = tuple(plates)
        V = -2 * np.asarray(phi[1])
        z = np.random.randn(*(plates + (N,)))
        Vb = np.broadcast_to(V, plates + (N, N))
        pb = np.broadcast_to(phi[0], plates + (N,))
        x = np.empty(plates + (N,))
        for idx in np.ndindex(*plates):
            U = scipy.linalg.cho_factor(Vb[idx])[0]
            mu = scipy.linalg.cho_solve((U, False), pb[idx])
            x[idx] = mu + scipy.linalg.solve_triangular(U, z[idx], trans="N", lower=False)
        return x


class Gaussian(_GaussianNode):
    """``Gaussian(mu, Lambda, plates=None, name="")`` (gaussian.py:1346-1417)."""

    def __init__(self, mu, Lambda, plates=None, name="", initialize=True, plates_multiplier=None):
        from .wishart import ensure_wishart
        Lambda = ensure_wishart(Lambda)
        Dm = Lambda.dims[0][-1]
        if isinstance(mu, Node) and hasattr(mu, "_to_gaussian"):
            mu = mu._to_gaussian()        # a Gaussian Markov chain as the mean: Gaussian vectors plated over time
        dist = GaussianDistribution(Dm)
        if isinstance(mu, Node):
            if mu.moment_kind not in ("gaussian", "gaussian_gamma") or tuple(mu.dims[0]) != (Dm,):
                raise ValueError("Mean and precision have inconsistent shapes: {0} and {1}".format(
                    mu.dims, Lambda.dims))
            if mu.moment_kind == "gaussian_gamma":
                from .gaussian_gamma import GaussianScaledMeanDistribution
                dist = GaussianScaledMeanDistribution(Dm)
        else:
            mu = gaussian_constant(mu, 1)
            if tuple(mu.dims[0]) != (Dm,):
                raise ValueError("Mean and precision have inconsistent shapes: {0} and {1}".format(
                    mu.dims, Lambda.dims))
        super().__init__(mu, Lambda, dims=((Dm,), (Dm, Dm)), distribution=dist,
                         plates=plates, name=name, initialize=initialize, plates_multiplier=plates_multiplier)

    def rotate(self, R, inv=None, logdet=None, Q=None):
        """q(x) -> q(R x) (gaussian.py:1451-1520 without the plate mixing Q): natural parameters by R^-T, moments by R,
        log-normaliser by -log|det R|; the plated arrays are transformed on the device."""
        if Q is not None:
            raise NotImplementedError("Plate mixing of a Gaussian node is not implemented")
        R = np.asarray(R, dtype=np.float64)
        invR = np.linalg.inv(R) if inv is None else np.asarray(inv, dtype=np.float64)
        logdetR = np.linalg.slogdet(R)[1] if logdet is None else float(logdet)
        K = self.dims[0][0]
        Rd, iRT = D.asarray(R), D.asarray(np.ascontiguousarray(invR.T))

        def rot_vec(a, Mx):
            a = D.asarray(dense(a))
            return D.sum_product([Mx, a.reshape((-1, K))], [["i", "k"], ["n", "k"]], ["n", "i"]).reshape(a.shape)

        def rot_mat(a, Mx):
            a = D.asarray(dense(a))
            return D.sum_product([Mx, a.reshape((-1, K, K)), Mx], [["i", "k"], ["n", "k", "l"], ["j", "l"]],
                                 ["n", "i", "j"]).reshape(a.shape)
        self.phi = [rot_vec(self.phi[0], iRT), rot_mat(self.phi[1], iRT)]
        self.u = [rot_vec(self.u[0], Rd), rot_mat(self.u[1], Rd)]
        self.g = D.affine(D.asarray(self.g), 1.0, -logdetR)

    def initialize_from_parameters(self, mu, Lambda):
        """q <- N(mu, Lambda^-1)  (gaussian.py:1420-1423)."""
        from .wishart import wishart_constant
        u_mu = gaussian_constant(mu, 1).get_moments()
        u_L = wishart_constant(Lambda).get_moments()
        self.phi = self._canonical_phi(self._distribution.compute_phi_from_parents(u_mu, u_L))
        u, g = self._distribution.compute_moments_and_cgf(self.phi)
        self._store(u, g, np.logical_not(self.observed))
