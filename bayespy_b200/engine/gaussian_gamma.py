"""Gaussian-gamma nodes on device: a Gaussian vector whose precision carries a common gamma-distributed scale.

Replaces the array math of ``bayespy/inference/vmp/nodes/gaussian.py``: GaussianGammaMoments :161-226,
GaussianGammaDistribution :892-1136, the node GaussianGamma :1777-2142 and the converter GaussianToGaussianGamma
:2226-2276.  As for ``Gaussian`` (engine/gaussian.py) the reference's glue nodes are folded in: a GaussianGamma has its
four parents (mu, Lambda, a, b) directly and forms the products with <Lambda> that WrapToGaussianWishart :2374-2527 forms
in the reference, so the messages reaching the user-visible parents are the same.

Moment kind ``"gaussian_gamma"``: [<tau x>, <tau x x^T>, <tau>, <log tau>], dims (S, S+S, (), ()).

    p(x, tau | mu, Lambda, a, b) = N(x | mu, (tau Lambda)^-1) Gamma(tau | a, b)
    log q(x, tau) = tau x^T phi0 + tau x^T Phi1 x + tau phi2 + log(tau) phi3 + g + f

All per-plate K x K work goes through ``bpk_gaussian_moments`` (the kernel of the Gaussian nodes), the gamma part through
``bpk_gamma_moments``; the products are ``bpk_sum_multiply`` launches.  ndim 0 (scalars, Lambda a gamma-like precision)
and ndim 1 (vectors, Lambda Wishart-like) are supported, like the reference ("currently, supports only vector
variables").
"""
import numpy as np

from .. import darray as D
from ..darray import DArray
from .expfam import Distribution, ExponentialFamily
from .gamma import GammaDistribution, gamma_prior_constant
from .gaussian import (GaussianARDDistribution, GaussianDimsToPlates, GaussianDistribution, LOG2PI, dense,
                       ensure_gamma, gaussian_constant, gaussian_moments_device)
from .node import Deterministic, Node


# --------------------------------------------------------------------------------------------
# small contraction helpers over broadcast plates
# --------------------------------------------------------------------------------------------
def _pk(n):
    return [("p", j) for j in range(n, 0, -1)]


def _mvdot(A, x):
    """A (.., K, K) x (.., K) -> (.., K) with broadcast plates."""
    npl = max(A.ndim - 2, x.ndim - 1)
    pk = _pk(npl)
    return D.sum_product([A, x], [pk[npl - (A.ndim - 2):] + ["i", "j"], pk[npl - (x.ndim - 1):] + ["j"]], pk + ["i"])


def _inner(x, y, nd):
    """Sum over the last ``nd`` axes of x * y with broadcast plates."""
    if nd == 0:
        return D.mul(x, y)
    npl = max(x.ndim, y.ndim) - nd
    pk = _pk(npl)
    dk = ["d%d" % i for i in range(nd)]
    return D.sum_product([x, y], [pk[npl - (x.ndim - nd):] + dk, pk[npl - (y.ndim - nd):] + dk], pk)


def _outer(x, y):
    return D.mul(x.add_trailing(1), y.expand_dims(-2))


def _scale(a, s, nd):
    """a (.., dims) times s (..) broadcast over the last nd axes."""
    s = D.asarray(s)
    return D.mul(a, s.add_trailing(nd) if nd and s.ndim else s)


# --------------------------------------------------------------------------------------------
# parents
# --------------------------------------------------------------------------------------------
class GaussianToGaussianGamma(Deterministic):
    """A Gaussian node seen as Gaussian-gamma with tau = 1: moments [x, xx, 1, 0]; only the first two entries of a
    message reach the parent (gaussian.py:2226-2276)."""
    moment_kind = "gaussian_gamma"

    def __init__(self, X, name=""):
        shape = tuple(X.dims[0])
        super().__init__(X, dims=(shape, shape + shape, (), ()), name=name or X.name)

    def _compute_moments(self, u):
        return [u[0], dense(u[1]), D.asarray(1.0), D.asarray(0.0)]

    def _compute_message_to_parent(self, index, m, u):
        return [m[0], m[1]]


def gaussian_gamma_constant(x, alpha, ndim):
    """Fixed (x, alpha): [alpha x, alpha x x^T, alpha, log alpha] (GaussianGammaMoments.compute_fixed_moments :182-203)."""
    from .node import Constant
    x = np.asarray(x, dtype=np.float64)
    alpha = np.asarray(alpha, dtype=np.float64)
    shape = x.shape[x.ndim - ndim:] if ndim else ()
    u0 = x * alpha.reshape(alpha.shape + (1,) * ndim)
    xx = x.reshape(x.shape + (1,) * ndim) * x.reshape(x.shape[:x.ndim - ndim] + (1,) * ndim + shape)
    u1 = xx * alpha.reshape(alpha.shape + (1,) * 2 * ndim)
    plates = np.broadcast_shapes(x.shape[:x.ndim - ndim], alpha.shape)
    with np.errstate(divide="ignore"):
        u = [D.asarray(u0), D.asarray(u1), D.asarray(alpha), D.asarray(np.log(alpha))]
    return Constant("gaussian_gamma", u, dims=(shape, shape + shape, (), ()), plates=plates, value=(x, alpha))


def ensure_gaussian_gamma(x, ndim):
    """Gaussian-gamma view of a parent given as an array, a Gaussian-like node or a Gaussian-gamma-like node."""
    if isinstance(x, Node) and hasattr(x, "_to_gaussian") and x.moment_kind not in ("gaussian", "gaussian_gamma"):
        x = x._to_gaussian()
    if not isinstance(x, Node):
        x = gaussian_constant(x, ndim)
    if x.moment_kind == "gaussian_gamma":
        if len(x.dims[0]) != ndim:
            raise NotImplementedError("Conversion of Gaussian-gamma moments to a different ndim is not implemented")
        return x
    if x.moment_kind != "gaussian":
        raise ValueError("Expected a Gaussian-like or Gaussian-gamma-like node, got %s" % type(x).__name__)
    nd = len(x.dims[0])
    if nd != ndim:
        if ndim == 0:
            x = GaussianDimsToPlates(x)
        else:
            raise ValueError("The parent has %d variable axes, %d expected" % (nd, ndim))
    return GaussianToGaussianGamma(x)


def wrap_with_precision(u_mu, u_Lam, ndim, K):
    """Moments of (mu, alpha) joined with a precision Lambda: [alpha Lambda mu, alpha mu^T Lambda mu, alpha Lambda,
    log|alpha Lambda|] (WrapToGaussianWishart._compute_moments, gaussian.py:2438-2457)."""
    ax, axx, alpha, logalpha = u_mu[0], dense(u_mu[1]), u_mu[2], u_mu[3]
    Lam, logdet = u_Lam
    if ndim == 0:
        return [D.mul(Lam, ax), D.mul(Lam, axx), D.mul(Lam, alpha), D.add(logdet, logalpha)]
    return [_mvdot(Lam, ax), _inner(Lam, axx, 2), _scale(Lam, alpha, 2), D.axpby(1.0, logdet, float(K), logalpha)]


def message_through_precision(index, m, u_mu, u_Lam, ndim, K):
    """Split the message [m0, m1, m2, m3] meant for the joined (mu, alpha, Lambda) moments between the Gaussian-gamma
    parent (index 0) and the precision parent (index 1) (WrapToGaussianWishart._compute_message_to_parent :2462-2527)."""
    m = [None if mi is None else D.asarray(mi) for mi in m]
    if index == 0:
        Lam = u_Lam[0]
        if ndim == 0:
            return [D.mul(Lam, m[0]), D.mul(Lam, m[1]), D.mul(Lam, m[2]), m[3]]
        return [_mvdot(Lam, m[0]), _scale(Lam, m[1], 2), _inner(Lam, m[2], 2), D.mul(m[3], float(K))]
    ax, axx, alpha = u_mu[0], dense(u_mu[1]), u_mu[2]
    if ndim == 0:
        return [D.add(D.add(D.mul(ax, m[0]), D.mul(axx, m[1])), D.mul(alpha, m[2])), m[3]]
    sym = D.add(_outer(ax, m[0]), _outer(m[0], ax))
    t = D.add(D.mul(sym, 0.5), _scale(axx, m[1], 2))
    return [D.add(t, _scale(m[2], alpha, 2)), m[3]]


# --------------------------------------------------------------------------------------------
# the distribution
# --------------------------------------------------------------------------------------------
class GaussianGammaDistribution(Distribution):
    """gaussian.py:892-1136 on device arrays.  Parents: (mu as Gaussian-gamma moments, Lambda, a, b)."""

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.ndim = len(self.shape)
        if self.ndim > 1:
            raise NotImplementedError("GaussianGamma supports scalar (ndim=0) and vector (ndim=1) variables")
        self.K = int(self.shape[0]) if self.ndim else 1

    def compute_phi_from_parents(self, u_mu, u_Lam, u_a, u_b, mask=True):
        """[Lambda mu, -Lambda/2, -mu^T Lambda mu / 2 - b, a]  (:1035-1048)."""
        w = wrap_with_precision(u_mu, u_Lam, self.ndim, self.K)
        return [w[0], D.mul(w[2], -0.5), D.axpby(-0.5, w[1], -1.0, u_b[0]), u_a[0]]

    def compute_cgf_from_parents(self, u_mu, u_Lam, u_a, u_b):
        """log|Lambda| / 2 + a <log b> - lgamma(a)  (:1086-1096)."""
        w = wrap_with_precision(u_mu, u_Lam, self.ndim, self.K)
        return D.sub(D.axpby(0.5, w[3], 1.0, D.mul(u_a[0], u_b[1])), u_a[1])

    def compute_moments_and_cgf(self, phi, mask=True):
        """:1051-1083: the Gaussian part with V = -2 Phi1, then a gamma with rate b = -phi2 - mu^T phi0 / 2."""
        phi = [D.asarray(dense(v)) for v in phi]
        nd = self.ndim
        if nd == 0:
            V = D.mul(phi[1], -2.0)
            cov = D._unary("RECIP", V, 1.0)
            mu = D.mul(phi[0], cov)
            logdet = D.log(V)
        else:
            mu, cov, _, logdet = gaussian_moments_device(phi[0], phi[1], self.K)
        b = D.axpby(-1.0, phi[2], -0.5, _inner(mu, phi[0], nd))
        (tau, logtau), g_gamma = GammaDistribution().compute_moments_and_cgf([D.mul(b, -1.0), phi[3]])
        u0 = _scale(mu, tau, nd)
        mumu = D.square(mu) if nd == 0 else _outer(mu, mu)
        u1 = D.add(cov, _scale(mumu, tau, 2 * nd))
        g = D.axpby(0.5, logdet, 1.0, g_gamma)
        return [u0, u1, tau, logtau], g

    def compute_message_to_parent(self, parent, index, u, u_mu, u_Lam, u_a, u_b):
        """:937-1032, with the split between mu and Lambda of :2462-2527."""
        if index in (0, 1):
            m = [u[0], D.mul(u[2], -0.5), D.mul(dense(u[1]), -0.5), D.asarray(0.5)]
            return message_through_precision(index, m, u_mu, u_Lam, self.ndim, self.K)
        if index == 2:
            return [D.add(u[3], u_b[1]), D.asarray(-1.0)]
        if index == 3:
            return [D.mul(u[2], -1.0), u_a[0]]
        raise ValueError("Index out of bounds")

    def compute_fixed_moments_and_f(self, x_alpha, mask=True):
        """:1099-1114."""
        x, alpha = x_alpha
        x = np.asarray(x, dtype=np.float64)
        alpha = np.asarray(alpha, dtype=np.float64)
        if self.ndim and (x.ndim < 1 or x.shape[-1] != self.K):
            raise ValueError("Invalid shape")
        c = gaussian_gamma_constant(x, alpha, self.ndim)
        f = D.affine(c.u[3], self.K / 2.0 - 1.0, -0.5 * self.K * LOG2PI)
        return list(c.u), f

    def random(self, *phi, plates=None):
        """Host draw with the reference's call pattern (:1117-1136, including its stated simplification of the gamma
        part)."""
        alpha = GammaDistribution().random(phi[2], phi[3], plates=plates)
        nd = self.ndim
        a1 = np.reshape(alpha, np.shape(alpha) + (1,) * nd)
        a2 = np.reshape(alpha, np.shape(alpha) + (1,) * 2 * nd)
        mu = GaussianARDDistribution(self.shape).random(a1 * phi[0], a2 * phi[1], plates=plates)
        return (mu, alpha)


class GaussianGamma(ExponentialFamily):
    """``GaussianGamma(mu, Lambda, a, b, ndim=1, plates=None, name="")`` (gaussian.py:1777-1850).  ``mu`` may be an
    array, a Gaussian-like or a Gaussian-gamma-like node; ``Lambda`` Wishart-like (ndim=1) or gamma-like (ndim=0)."""
    moment_kind = "gaussian_gamma"

    def __init__(self, mu, Lambda, a, b, ndim=1, plates=None, name="", initialize=True, plates_multiplier=None):
        if ndim not in (0, 1):
            raise NotImplementedError("GaussianGamma supports ndim 0 and 1")
        mu = ensure_gaussian_gamma(mu, ndim)
        shape = tuple(mu.dims[0])
        if ndim == 0:
            Lambda = ensure_gamma(Lambda)
        else:
            from .wishart import ensure_wishart
            Lambda = ensure_wishart(Lambda)
            if tuple(Lambda.dims[0]) != shape + shape:
                raise ValueError("Mean and precision have inconsistent shapes: {0} and {1}".format(mu.dims, Lambda.dims))
        if isinstance(a, Node):
            if a.moment_kind != "gamma_prior":
                raise ValueError("a has wrong shape")
        else:
            a = gamma_prior_constant(a)
        b = ensure_gamma(b)
        super().__init__(mu, Lambda, a, b, dims=(shape, shape + shape, (), ()),
                         distribution=GaussianGammaDistribution(shape), plates=plates, name=name,
                         initialize=initialize, plates_multiplier=plates_multiplier)

    @property
    def ndim(self):
        return len(self.dims[0])

    # ---- transformations of q (gaussian.py:1853-1945) ----------------------------------------------------------
    def translate(self, b):
        """q(x, tau) -> q(x + b, tau)."""
        if self.ndim != 1:
            raise NotImplementedError("Only ndim=1 supported at the moment")
        b = D.asarray(np.asarray(b, dtype=np.float64))
        tau = D.asarray(self.u[2])
        x = D.div(self.u[0], tau.add_trailing(1))
        xb = _outer(x, b)
        bx = _outer(b, x)
        bb = _outer(b, b)
        Lam = D.mul(D.asarray(dense(self.phi[1])), -2.0)
        Lb = _mvdot(Lam, b)
        dtau = D.axpby(-0.5, _inner(Lb, b, 1), -1.0, _inner(Lb, x, 1))
        u0 = D.add(self.u[0], D.mul(tau.add_trailing(1), b))
        u1 = D.add(dense(self.u[1]), _scale(D.add(D.add(xb, bx), bb), tau, 2))
        self.phi = self._canonical_phi([D.add(self.phi[0], Lb), self.phi[1], D.add(self.phi[2], dtau), self.phi[3]])
        self.u = [u0, u1, self.u[2], self.u[3]]

    def rotate(self, R, inv=None, logdet=None):
        """q(x, tau) -> q(R x, tau)."""
        if self.ndim != 1:
            raise NotImplementedError("Only ndim=1 supported at the moment")
        R = np.asarray(R, dtype=np.float64)
        inv = np.linalg.inv(R) if inv is None else np.asarray(inv, dtype=np.float64)
        logdet = np.linalg.slogdet(R)[1] if logdet is None else float(logdet)
        K = self.dims[0][0]
        Rd, iRT = D.asarray(R), D.asarray(np.ascontiguousarray(inv.T))

        def rot_vec(a, Mx):
            a = D.asarray(a)
            return D.sum_product([Mx, a.reshape((-1, K))], [["i", "k"], ["n", "k"]], ["n", "i"]).reshape(a.shape)

        def rot_mat(a, Mx):
            a = D.asarray(dense(a))
            return D.sum_product([Mx, a.reshape((-1, K, K)), Mx], [["i", "k"], ["n", "k", "l"], ["j", "l"]],
                                 ["n", "i", "j"]).reshape(a.shape)
        self.phi = [rot_vec(self.phi[0], iRT), rot_mat(self.phi[1], iRT), self.phi[2], self.phi[3]]
        self.u = [rot_vec(self.u[0], Rd), rot_mat(self.u[1], Rd), self.u[2], self.u[3]]
        self.g = D.affine(D.asarray(self.g), 1.0, -logdet)

    # ---- summaries on the host (gaussian.py:2009-2141) -----------------------------------------------------------
    def get_gaussian_location(self):
        if self.ndim != 1:
            raise NotImplementedError("Only ndim=1 supported at the moment")
        return np.asarray(self.u[0]) / np.asarray(self.u[2])[..., None]

    def get_gaussian_mean_and_variance(self):
        """Mean and variance of the Student-t marginal of x."""
        if self.ndim != 1:
            raise NotImplementedError("Only ndim=1 supported at the moment")
        a = np.asarray(self.phi[3])
        nu = 2 * a
        if np.any(nu <= 1):
            raise ValueError("Mean not defined for degrees of freedom <= 1")
        if np.any(nu <= 2):
            raise ValueError("Variance not defined if degrees of freedom <= 2")
        tau = np.asarray(self.u[2])[..., None]
        tau_mu = np.asarray(self.u[0])
        mu = tau_mu / tau
        var = (np.einsum("...ii->...i", np.asarray(dense(self.u[1]))) - tau_mu * mu) / tau
        return mu, (nu / (nu - 2))[..., None] * var if np.ndim(nu) else nu / (nu - 2) * var

    def __str__(self):
        return "%s ~ GaussianGamma" % self.name


# --------------------------------------------------------------------------------------------
# Gaussian nodes whose mean parent carries a gamma scale
# --------------------------------------------------------------------------------------------
class GaussianScaledMeanDistribution(GaussianDistribution):
    """x ~ N(mu, (alpha Lambda)^-1) for a Gaussian-gamma (mu, alpha) parent: the Gaussian formulas of gaussian.py:341-463
    on the joined moments of WrapToGaussianWishart :2438-2457."""

    def compute_phi_from_parents(self, u_mu, u_Lambda, mask=True):
        w = wrap_with_precision(u_mu, u_Lambda, 1, self.D)
        return [w[0], D.mul(w[2], -0.5)]

    def compute_cgf_from_parents(self, u_mu, u_Lambda):
        w = wrap_with_precision(u_mu, u_Lambda, 1, self.D)
        return D.axpby(-0.5, w[1], 0.5, w[3])

    def compute_message_to_parent(self, parent, index, u, u_mu, u_Lambda):
        if index not in (0, 1):
            raise ValueError("Index out of bounds")
        m = [u[0], D.asarray(-0.5), D.mul(dense(u[1]), -0.5), D.asarray(0.5)]
        return message_through_precision(index, m, u_mu, u_Lambda, 1, self.D)


class GaussianARDScaledMeanDistribution(GaussianARDDistribution):
    """x ~ N(mu, diag(tau alpha)^-1) for a scalar Gaussian-gamma (mu, tau) parent plated over plates + shape: the formulas
    of gaussian.py:609-732 on the joined moments of WrapToGaussianGamma :2339-2371."""

    def _wrap(self, u_mu, u_alpha):
        alpha, logalpha = u_alpha
        return [D.mul(u_mu[0], alpha), D.mul(u_mu[1], alpha), D.mul(u_mu[2], alpha), D.add(u_mu[3], logalpha)]

    def compute_phi_from_parents(self, u_mu, u_alpha, mask=True):
        w = self._wrap(u_mu, u_alpha)
        if self.ndim == 0:
            return [w[0], D.mul(w[2], -0.5)]
        phi0 = self._expand_var_axes(w[0])
        a = self._expand_var_axes(w[2])
        phi1 = DArray.zeros(tuple(a.shape) + self.shape)
        D._ew("AFFINE", a.shape, phi1.diag_view(self.ndim), [a], alpha=-0.5, beta=0.0)
        return [phi0, phi1]

    def compute_cgf_from_parents(self, u_mu, u_alpha):
        w = self._wrap(u_mu, u_alpha)
        if self.ndim == 0:
            return D.axpby(-0.5, w[1], 0.5, w[3])
        nd, full = self.ndim, self.shape

        def sum_dims(a):
            a = a if a.ndim >= nd else a.add_leading(nd - a.ndim)
            npl = a.ndim - nd
            return D.reduce_to_shape(a, tuple(a.shape[:npl]) + (1,) * nd,
                                     from_shape=tuple(a.shape[:npl]) + full).reshape(a.shape[:npl])
        return D.axpby(-0.5, sum_dims(w[1]), 0.5, sum_dims(w[3]))

    def compute_message_to_parent(self, parent, index, u, u_mu, u_alpha):
        x = u[0]
        x2 = dense(u[1]).diag_view(self.ndim) if self.ndim else u[1]
        half = DArray.full(self.shape, 0.5) if self.ndim else D.asarray(0.5)
        if index == 0:
            alpha = u_alpha[0]
            return [D.mul(alpha, x), D.mul(alpha, -0.5), D.mul(D.mul(alpha, x2), -0.5), half]
        if index == 1:
            t = D.sub(D.mul(x, u_mu[0]), D.mul(u_mu[1], 0.5))
            return [D.sub(t, D.mul(D.mul(x2, u_mu[2]), 0.5)), half]
        raise ValueError("Invalid parent index")
