"""Dirichlet node on device (replaces the array math of nodes/dirichlet.py:107-389).
phi = [alpha]; u = [<log p>] = psi(phi) - psi(sum phi); one kernel, ``bpk_dirichlet_moments``."""
import numpy as np

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .expfam import Distribution, ExponentialFamily
from .node import Constant, Node


def concentration_constant(alpha):
    """[alpha, lgamma(sum alpha) - sum lgamma(alpha)]  (ConcentrationMoments, dirichlet.py:25-62)."""
    alpha = np.asarray(alpha, dtype=np.float64)
    if alpha.ndim < 1:
        raise ValueError("The prior sample sizes must be a vector")
    if np.any(alpha < 0):
        raise ValueError("The prior sample sizes must be non-negative")
    ad = D.asarray(alpha)
    K = alpha.shape[-1]
    s = D.sum_product([ad], [list(range(ad.ndim))], list(range(ad.ndim - 1)))
    lg = D.sum_product([D.gammaln(ad)], [list(range(ad.ndim))], list(range(ad.ndim - 1)))
    z = D.sub(D.gammaln(s), lg)
    return Constant("dirichlet_prior", [ad, z], dims=((K,), ()), plates=alpha.shape[:-1], value=alpha)


def dirichlet_constant(p):
    """[log p] of fixed probabilities (DirichletMoments.compute_fixed_moments)."""
    p = np.asarray(p, dtype=np.float64)
    if p.ndim < 1:
        raise ValueError("Probabilities must be given as a vector")
    if np.any(p < 0) or np.any(p > 1):
        raise ValueError("Probabilities must be in range [0,1]")
    if not np.allclose(np.sum(p, axis=-1), 1.0):
        raise ValueError("Probabilities must sum to one")
    p = p / np.sum(p, axis=-1, keepdims=True)
    with np.errstate(divide="ignore"):
        return Constant("dirichlet", [D.log(D.asarray(p))], dims=((p.shape[-1],),), plates=p.shape[:-1], value=p)


class DirichletDistribution(Distribution):

    def compute_gradient(self, g, u, phi):
        """dirichlet.py:213-231, as the reference computes it: g * (psi'(phi) - psi'(sum_d phi_d)), elementwise."""
        ph, g0 = D.asarray(phi[0]), D.asarray(g[0])
        keys = list(range(ph.ndim))
        tot = D.sum_product([ph], [keys], keys[:-1])
        return [D.mul(g0, D.sub(D.trigamma(ph), D.trigamma(tot).add_trailing(1)))]

    def compute_message_to_parent(self, parent, index, u_self, u_alpha):
        return [u_self[0], D.asarray(1.0)]

    def compute_phi_from_parents(self, u_alpha, mask=True):
        return [u_alpha[0]]

    def compute_cgf_from_parents(self, u_alpha):
        return u_alpha[1]

    def compute_moments_and_cgf(self, phi, mask=True):
        """dirichlet.py:130-160; raises ValueError("Natural parameters should be positive")."""
        phi = [D.asarray(v) for v in phi]
        p = phi[0].contiguous()
        K = p.shape[-1]
        P = tuple(p.shape[:-1])
        n = int(np.prod(P, dtype=np.int64)) if P else 1
        u, g = DArray.empty(P + (K,)), DArray.empty(P)
        _bpk.get().dirichlet_moments(p.ptr, n, K, u.ptr, g.ptr, True)
        return [u], g

    def compute_fixed_moments_and_f(self, p, mask=True):
        c = dirichlet_constant(p)
        logp = c.u[0]
        f = D.sum_product([logp], [list(range(logp.ndim))], list(range(logp.ndim - 1)), scale=-1.0)
        return [logp], f

    def random(self, *phi, plates=None):
        # utils/random.py:329-347: normalised gamma draws (host RNG, so that seeded initialisations reproduce the
        # reference draw for draw), also for plated concentration parameters
        alpha = np.asarray(phi[0], dtype=np.float64)
        size = tuple(plates) + alpha.shape[-1:] if plates is not None else alpha.shape
        p = np.random.gamma(alpha, size=size)
        total = np.sum(p, axis=-1, keepdims=True)
        if np.any(total == 0):
            raise RuntimeError("Numerically zero samples. Try using a larger Dirichlet concentration parameter value.")
        return p / total


class Dirichlet(ExponentialFamily):
    """``Dirichlet(alpha, plates=None, name="")`` (dirichlet.py:333-389)."""
    moment_kind = "dirichlet"
    _guard_zero_times_inf = True

    def __init__(self, alpha, plates=None, name="", initialize=True):
        if isinstance(alpha, Node):
            if alpha.moment_kind != "dirichlet_prior":
                raise ValueError("Concentration must be a fixed array")
        else:
            alpha = concentration_constant(alpha)
        K = alpha.dims[0][0]
        super().__init__(alpha, dims=((K,),), distribution=DirichletDistribution(), plates=plates, name=name,
                         initialize=initialize)

    def __str__(self):
        return "%s ~ Dirichlet(alpha)\n  alpha =\n%s" % (self.name, self.phi[0].numpy())


class Concentration(Node):
    """Maximum-likelihood point estimate of the concentration parameters of Dirichlet variables (dirichlet.py:234-330).
    The children's messages are [sum <log p>, count]; with the optional regularisation ("prior" log-probability and
    sample number) the estimate solves psi(a_k) = psi(sum a) + mean <log p_k> by the reference's fixed-point iteration —
    a root over the hyper-parameter plates found on the host; the moments [a, lgamma(sum a) - sum lgamma(a)] live on
    the device."""
    moment_kind = "dirichlet_prior"

    def __init__(self, D_, regularization=True, plates=None, name=""):
        self.D = int(D_)
        super().__init__(dims=((self.D,), ()), plates=plates if plates is not None else (), name=name)
        self._id = Node._id_counter
        Node._id_counter += 1
        self.observed = False
        if regularization is None or regularization is False:
            regularization = [0, 0]
        elif regularization is True:
            regularization = [np.log(1 / self.D), 1]
        if len(regularization) != 2:
            raise ValueError("Regularization must 2-tuple")
        self.regularization = regularization
        self.initialize_from_value(np.ones(self.D))

    def _ids(self):
        return [self._id]

    def initialize_from_value(self, x):
        x = np.asarray(x, dtype=np.float64) * np.ones(tuple(self.plates) + (self.D,))
        self.u = list(concentration_constant(x).u)
        self._version += 1

    def get_moments(self):
        return list(self.u)

    def update(self, annealing=1.0):
        import scipy.special as sp
        from .gamma import invpsi
        m = self.message_from_children()
        logp = np.asarray(m[0]) + self.regularization[0]
        n = np.asarray(m[1]) + self.regularization[1]
        mean_logp = logp / np.asarray(n)[..., None]
        if np.any(np.isinf(mean_logp)):
            raise ValueError("Cannot estimate DirichletConcentration because of infs. This means that there are "
                             "numerically zero probabilities in the child Dirichlet node.")
        a = np.ones(self.D)
        da = np.inf
        while np.any(np.abs(da / a) > 1e-5):
            a_new = invpsi(sp.psi(np.sum(a, axis=-1, keepdims=True)) + mean_logp)
            da = a_new - a
            a = a_new
        self.initialize_from_value(a)

    def lower_bound_contribution(self, ignore_masked=True):
        u0, u1 = np.asarray(self.u[0]), np.asarray(self.u[1])
        return float(np.sum(np.sum(u0 * self.regularization[0], axis=-1) + u1 * self.regularization[1]))

    def _update_mask(self):
        mask = np.array(False)
        for child, index in self.children:
            mask = np.logical_or(mask, child._mask_to_parent(index))
        self._set_mask(mask)


DirichletConcentration = Concentration


def BetaConcentration(**kwargs):
    return Concentration(2, **kwargs)
