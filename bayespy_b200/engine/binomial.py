"""Binomial, Bernoulli and Beta nodes (nodes/binomial.py:46-210, bernoulli.py:37-102, beta.py:44-175).

A binomial count is the first component of a two-category multinomial, so the device work is the same kernel:
with the log-odds phi = <log p> - <log(1-p)>, the max-shifted softmax of [phi, 0] gives sigmoid(phi) and its
log-sum-exp gives log(1 + e^phi) (``bpk_softmax_moments``): u = [N sigmoid(phi)], g = -N log(1 + e^phi).  A Beta
variable is a two-component Dirichlet whose moments are [<log p>, <log(1-p)>]."""
import numpy as np

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .dirichlet import Dirichlet, DirichletDistribution
from .expfam import Distribution, ExponentialFamily
from .node import Constant, Deterministic, Node, broadcast_plates


def beta_constant(p):
    """[log p, log(1-p)] of fixed probabilities (BetaMoments.compute_fixed_moments, beta.py:30-37)."""
    p = np.asarray(p, dtype=np.float64)
    if np.any(p < 0) or np.any(p > 1):
        raise ValueError("Probabilities must be in range [0,1]")
    pp = np.stack([p, 1.0 - p], axis=-1)
    with np.errstate(divide="ignore"):
        return Constant("dirichlet", [D.log(D.asarray(pp))], dims=((2,),), plates=p.shape, value=p)


def ensure_beta(p):
    if isinstance(p, Node):
        if p.moment_kind != "dirichlet" or tuple(p.dims) != ((2,),):
            raise ValueError("Expected a beta-like node (a two-component Dirichlet)")
        return p
    return beta_constant(p)


class BetaDistribution(DirichletDistribution):
    """Realisations are scalars p; the moments are the two-vector [log p, log(1-p)] (beta.py:44-96)."""

    def compute_fixed_moments_and_f(self, p, mask=True):
        p = np.asarray(p, dtype=np.float64)
        return super().compute_fixed_moments_and_f(np.stack([p, 1.0 - p], axis=-1), mask=mask)

    def random(self, *phi, plates=None):
        return super().random(*phi, plates=plates)[..., 0]


class Beta(Dirichlet):
    """``Beta(alpha)`` with ``alpha[..., 0:2] = (a, b)``, prior counts of success and failure (beta.py:99-160)."""

    def __init__(self, alpha, plates=None, name="", initialize=True):
        super().__init__(alpha, plates=plates, name=name, initialize=False)
        if tuple(self.dims) != ((2,),):
            raise ValueError("Parent has wrong dimensionality. Must be a two-dimensional vector.")
        self._distribution = BetaDistribution()
        if initialize:
            self.initialize_from_prior()

    def complement(self):
        return Complement(self)

    def __str__(self):
        a = self.phi[0].numpy()
        return "%s ~ Beta(a, b)\n  a = \n%s\n  b = \n%s\n" % (self.name, a[..., 0], a[..., 1])


class Complement(Deterministic):
    """1 - p of a beta-like node: the two moments [log p, log(1-p)] swap places, and so does a message (beta.py:178-203)."""
    moment_kind = "dirichlet"

    def __init__(self, p, name=""):
        p = ensure_beta(p)
        super().__init__(p, dims=p.dims, name=name)

    @staticmethod
    def _swap(a):
        a = D.asarray(a)
        st = list(a.strides)
        esz = 8
        st[-1] = -st[-1]
        return DArray(a.owner, a.ptr + (a.shape[-1] - 1) * a.strides[-1] * esz, a.shape, st, a.dtype)

    def _compute_moments(self, u_p):
        return [self._swap(u_p[0]).contiguous()]

    def _compute_message_to_parent(self, index, m, u_p):
        if index != 0:
            raise IndexError()
        return [None if m[0] is None else self._swap(m[0]).contiguous()]


class BinomialDistribution(Distribution):
    zero_times_inf = True          # a probability of exactly 0 or 1 has an infinite log-odds

    def __init__(self, N):
        N = np.asarray(N)
        if not issubclass(N.dtype.type, np.integer):
            raise ValueError("Number of trials must be integer")
        if np.any(N < 0):
            raise ValueError("Number of trials must be non-negative")
        self.N = N
        self._Nd = D.asarray(N.astype(np.float64))

    @staticmethod
    def _split(logp):
        logp = D.asarray(logp)
        last = logp.ndim - 1
        sh = tuple(logp.shape[:-1])
        return logp.slice_axis(last, 0, 1).reshape(sh), logp.slice_axis(last, 1, 2).reshape(sh)

    def compute_message_to_parent(self, parent, index, u_self, u_p):
        """[x, n - x] (binomial.py:62-72)."""
        if index != 0:
            raise ValueError("Incorrect parent index")
        x = D.asarray(u_self[0])
        P = tuple(np.broadcast_shapes(tuple(x.shape), self.N.shape))
        out = DArray.empty(P + (2,))
        D.copy_into(out.slice_axis(len(P), 0, 1).reshape(P), x)
        D.copy_into(out.slice_axis(len(P), 1, 2).reshape(P), D.sub(self._Nd, x))
        return [out]

    def compute_phi_from_parents(self, u_p, mask=True):
        l0, l1 = self._split(u_p[0])
        return [D.sub(l0, l1)]

    def compute_moments_and_cgf(self, phi, mask=True):
        """u = [N / (1 + e^-phi)], g = -N log(1 + e^phi) (binomial.py:82-88)."""
        ph = D.asarray(phi[0])
        P = tuple(ph.shape)
        z = DArray.zeros(P + (2,))
        D.copy_into(z.slice_axis(len(P), 0, 1).reshape(P), ph)
        n = int(np.prod(P, dtype=np.int64)) if P else 1
        soft, g = DArray.empty(P + (2,)), DArray.empty(P)
        _bpk.get().softmax_moments(z.ptr, n, 2, soft.ptr, g.ptr)
        p1 = soft.slice_axis(len(P), 0, 1).reshape(P)
        return [D.mul(p1, self._Nd)], D.mul(g, self._Nd)

    def compute_cgf_from_parents(self, u_p):
        _, l1 = self._split(u_p[0])
        return D.mul(l1, self._Nd)

    def compute_fixed_moments_and_f(self, x, mask=True):
        x = np.asarray(x)
        if not issubclass(x.dtype.type, (np.integer, np.bool_)):
            raise ValueError("Counts must be integer")
        x = x.astype(np.int64)
        if np.any(x < 0) or np.any(x > self.N):
            raise ValueError("Invalid count")
        xd = D.asarray(x.astype(np.float64))
        f = D.sub(D.sub(D.gammaln(D.affine(self._Nd, 1.0, 1.0)), D.gammaln(D.affine(xd, 1.0, 1.0))),
                  D.gammaln(D.affine(D.sub(self._Nd, xd), 1.0, 1.0)))
        return [xd], f

    def random(self, *phi, plates=None):
        p = 1.0 / (1.0 + np.exp(-np.asarray(phi[0])))
        return np.random.binomial(self.N, p, size=plates)

    def squeeze(self, axis):
        if self.N.ndim < -axis:
            return self
        try:
            N = np.squeeze(self.N, axis)
        except ValueError as err:
            raise ValueError("The number of trials must be constant over a squeezed axis, so the corresponding array "
                             "axis must be singleton. Cannot squeeze axis {0} from a binomial distribution because "
                             "the number of trials arrays has shape {2}, so the given axis has length {1} != 1. "
                             .format(axis, np.shape(self.N)[axis], np.shape(self.N))) from err
        return type(self)(N) if type(self) is BinomialDistribution else self


class BernoulliDistribution(BinomialDistribution):

    def __init__(self, N=1):
        super().__init__(np.asarray(1))

    def squeeze(self, axis):
        return self


class Binomial(ExponentialFamily):
    """``Binomial(n, p, plates=None, name="")`` (binomial.py:135-210); p a beta-like node or an array."""
    moment_kind = "binomial"
    _guard_zero_times_inf = True

    def __init__(self, n, p, plates=None, name="", initialize=True, plates_multiplier=None):
        p = ensure_beta(p)
        dist = BinomialDistribution(n)
        total = broadcast_plates(tuple(p.plates), np.shape(n))
        if plates is not None:
            plates = tuple(int(v) for v in plates)
            if broadcast_plates(total, plates) != plates:
                raise ValueError("The plates %s of the parents are not broadcastable to the given plates %s."
                                 % (total, plates))
            total = plates
        super().__init__(p, dims=((),), distribution=dist, plates=total, name=name, initialize=initialize,
                         plates_multiplier=plates_multiplier)

    def __str__(self):
        p = 1 / (1 + np.exp(-self.phi[0].numpy()))
        return "%s ~ Binomial(n, p)\n  n = \n%s\n  p = \n%s\n" % (self.name, self._distribution.N, p)


class Bernoulli(ExponentialFamily):
    """``Bernoulli(p, plates=None, name="")`` (bernoulli.py:44-102): a binomial with one trial."""
    moment_kind = "binomial"
    _guard_zero_times_inf = True

    def __init__(self, p, plates=None, name="", initialize=True, plates_multiplier=None):
        p = ensure_beta(p)
        super().__init__(p, dims=((),), distribution=BernoulliDistribution(), plates=plates, name=name,
                         initialize=initialize, plates_multiplier=plates_multiplier)

    def __str__(self):
        return "%s ~ Bernoulli(p)\n  p = \n%s\n" % (self.name, 1 / (1 + np.exp(-self.phi[0].numpy())))
