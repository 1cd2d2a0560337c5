"""Gamma node on device (replaces the array math of nodes/gamma.py:90-241).

x ~ Gamma(a, b) with fixed shape ``a`` and a gamma-like (or fixed) rate ``b``.
Natural parameters phi = [-<b>, a]; moments u = [<x>, <log x>];
all of it is one fused kernel, ``bpk_gamma_moments``.
"""
import numpy as np

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .expfam import Distribution, ExponentialFamily
from .gaussian import ensure_gamma
from .node import Constant, Deterministic, Node


def gamma_prior_constant(a):
    """[a, lgamma(a)]  (GammaPriorMoments, gamma.py:33-59)."""
    a = np.asarray(a, dtype=np.float64)
    if np.any(a <= 0):
        raise ValueError("Shape parameter must be positive")
    ad = D.asarray(a)
    return Constant("gamma_prior", [ad, D.gammaln(ad)], dims=((), ()), plates=a.shape, value=a)


class GammaToDiagonalWishart(Deterministic):
    """Gamma scalars whose last plate axis becomes the diagonal of a Wishart-like matrix: [diag(x), sum log x]
    (gamma.py:337-397)."""
    moment_kind = "wishart"

    def __init__(self, alpha, name=""):
        alpha = ensure_gamma(alpha)
        if len(alpha.plates) == 0:
            raise Exception("Gamma variable needs to have plates in order to be used as a diagonal Wishart.")
        self.Dm = int(alpha.plates[-1])
        super().__init__(alpha, dims=((self.Dm, self.Dm), ()), name=name)

    def _plates_to_parent(self, index):
        return tuple(self.plates) + (self.Dm,)

    def _plates_from_parent(self, index):
        return tuple(self.parents[index].plates[:-1])

    def _map_parent_axes(self, index, values):
        return tuple(values[:-1])

    def _weights_to_parent(self, index, mask):
        return np.asarray(mask)[..., np.newaxis]

    def _compute_moments(self, u):
        x = D.asarray(u[0]).broadcast_to(tuple(D.asarray(u[0]).shape[:-1]) + (self.Dm,))
        out = DArray.zeros(tuple(x.shape) + (self.Dm,))
        D.copy_into(out.diag_view(1), x)
        lx = D.asarray(u[1])
        return [out, D.reduce_to_shape(lx, tuple(lx.shape[:-1]) + (1,),
                                       from_shape=tuple(lx.shape[:-1]) + (self.Dm,)).reshape(lx.shape[:-1])]

    def _compute_message_to_parent(self, index, m, u):
        m0 = None if m[0] is None else D.asarray(m[0]).diag_view(1)
        m1 = None if m[1] is None else D.asarray(m[1]).add_trailing(1)
        return [m0, m1]


class GammaDistribution(Distribution):

    def compute_gradient(self, g, u, phi):
        """gamma.py:183-211.  With b = -phi0, a = phi1: <u> = [a / b, psi(a) - log b], so the Fisher information is
        [[a / b^2, 1 / b], [1 / b, psi'(a)]]."""
        b = D.mul(D.asarray(phi[0]), -1.0)
        a = D.asarray(phi[1])
        g0, g1 = D.asarray(g[0]), D.asarray(g[1])
        inv_b = D.div(1.0, b)
        tri = D.trigamma(a)
        d0 = D.add(D.mul(g0, D.mul(a, D.mul(inv_b, inv_b))), D.mul(g1, inv_b))
        d1 = D.add(D.mul(g1, tri), D.mul(g0, inv_b))
        return [d0, d1]

    def compute_message_to_parent(self, parent, index, u_self, u_a, u_b):
        """gamma.py:96-113."""
        if index == 0:
            # to the shape a, whose moments are [a, lgamma(a)]: [<log x> + <log b>, -1]  (gamma.py:96-106)
            return [D.add(u_self[1], u_b[1]), D.asarray(-1.0)]
        elif index == 1:
            return [D.mul(u_self[0], -1.0), u_a[0]]
        raise ValueError("Index out of bounds")

    def compute_phi_from_parents(self, u_a, u_b, mask=True):
        """gamma.py:116-121."""
        return [D.mul(u_b[0], -1.0), u_a[0]]

    def compute_cgf_from_parents(self, u_a, u_b):
        """a <log b> - lgamma(a)   (gamma.py:151-160)."""
        return D.sub(D.mul(u_a[0], u_b[1]), u_a[1])

    def compute_moments_and_cgf(self, phi, mask=True):
        """gamma.py:124-148 as one kernel."""
        phi = [D.asarray(v) for v in phi]
        be = _bpk.get()
        P = tuple(np.broadcast_shapes(phi[0].shape, phi[1].shape))
        n = int(np.prod(P, dtype=np.int64)) if P else 1

        def prep(p):
            cnt = p.size
            if cnt != 1 and tuple(p.shape) != P and (1,) * (len(P) - p.ndim) + tuple(p.shape) != P:
                p = p.broadcast_to(P)
                cnt = n
            return p.contiguous(), (1 if cnt == 1 else n)
        p0, n0 = prep(phi[0])
        p1, n1 = prep(phi[1])
        u0, u1, g = DArray.empty(P), DArray.empty(P), DArray.empty(P)
        be.gamma_moments(p0.ptr, n0, p1.ptr, n1, n, u0.ptr, u1.ptr, g.ptr, True)
        return [u0, u1], g

    def compute_fixed_moments_and_f(self, x, mask=True):
        """gamma.py:163-173."""
        x = np.asarray(x, dtype=np.float64)
        if np.any(x < 0):
            raise ValueError("Values must be positive")
        xd = D.asarray(x)
        logx = D.log(xd)
        return [xd, logx], D.mul(logx, -1.0)

    def random(self, *phi, plates=None):
        return np.random.gamma(phi[1], -1 / phi[0], size=plates)


class Gamma(ExponentialFamily):
    """``Gamma(a, b, plates=None, name="")`` as in the reference (gamma.py:214-241)."""
    moment_kind = "gamma"

    def __init__(self, a, b, plates=None, name="", initialize=True):
        if isinstance(a, Node):
            if a.moment_kind != "gamma_prior":
                raise ValueError("Shape parameter must be a fixed value or a gamma-prior node")
        else:
            a = gamma_prior_constant(a)
        b = ensure_gamma(b)
        super().__init__(a, b, dims=((), ()), distribution=GammaDistribution(), plates=plates, name=name,
                         initialize=initialize)

    def initialize_from_parameters(self, a, b):
        """q <- Gamma(a, b)  (expfamily.py:166-180 with the fixed moments of the parents)."""
        u_a = gamma_prior_constant(np.asarray(a, dtype=np.float64) * np.ones(np.shape(b))).get_moments()
        u_b = ensure_gamma(np.asarray(b, dtype=np.float64) * np.ones(np.shape(a))).get_moments()
        self.phi = self._canonical_phi(self._distribution.compute_phi_from_parents(u_a, u_b))
        u, g = self._distribution.compute_moments_and_cgf(self.phi)
        self._store(u, g, np.logical_not(self.observed))

    def as_wishart(self, ndim=0):
        """[x, log x] are the moments of a 0-dimensional Wishart variable as they stand (gamma.py:258-261, :399-430)."""
        if ndim != 0:
            raise NotImplementedError()
        return self

    def as_diagonal_wishart(self):
        return GammaToDiagonalWishart(self, name=self.name + " as Wishart")

    def diag(self):
        return self.as_diagonal_wishart()

    def __str__(self):
        a = self.phi[1].numpy()
        b = -self.phi[0].numpy()
        return "%s ~ Gamma(a, b)\n  a =\n%s\n  b =\n%s\n" % (self.name, a, b)


def invpsi(x):
    """Inverse digamma function on host arrays: Minka's initial guess and five Newton steps (misc.py:1404-1429,
    the same recipe so that estimates agree digit for digit)."""
    import scipy.special as sp
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore", divide="ignore"):
        y = np.where(x >= -2.22, np.exp(x) + 0.5, -1.0 / (x - sp.psi(1)))
    for _ in range(5):
        y = y - (sp.psi(y) - x) / sp.polygamma(1, y)
    return y


class GammaShape(Node):
    """Maximum-likelihood point estimate of the shape parameter of gamma variables (gamma.py:273-335).  The children's
    messages enter the bound as m0 a + m1 lgamma(a), so a = psi^-1(-m0 / m1): a root of the size of the hyper-parameter
    plates, found on the host with the reference's Newton recipe; the moments [a, lgamma(a)] live on the device."""
    moment_kind = "gamma_prior"

    def __init__(self, m0=0, m1=0, plates=None, name=""):
        super().__init__(dims=((), ()), plates=plates if plates is not None else (), name=name)
        self._id = Node._id_counter
        Node._id_counter += 1
        self._m0, self._m1 = m0, m1
        self.observed = False
        self.initialize_from_value(1.0)

    def _ids(self):
        return [self._id]

    def initialize_from_value(self, x):
        self.u = list(gamma_prior_constant(np.asarray(x, dtype=np.float64) * np.ones(self.plates)).u)
        self._version += 1

    def get_moments(self):
        return list(self.u)

    def update(self, annealing=1.0):
        m = self.message_from_children()
        m0 = self._m0 + np.asarray(m[0])
        m1 = self._m1 + np.asarray(m[1])
        self.initialize_from_value(invpsi(-m0 / m1))

    def lower_bound_contribution(self, ignore_masked=True):
        return 0.0

    def _update_mask(self):
        mask = np.array(False)
        for child, index in self.children:
            mask = np.logical_or(mask, child._mask_to_parent(index))
        self._set_mask(mask)
