"""Multinomial node on device (nodes/multinomial.py:60-319): counts x over K categories in N trials with log-probabilities
from a Dirichlet-like parent.  phi = [<log p>]; u = [N softmax(phi)]; g = -N logsumexp(phi).  The softmax with the
reference's max-shift and second renormalisation is the kernel of the Categorical node (``bpk_softmax_moments``,
multinomial.py:101-121 / misc.py:1366-1401); Categorical is this node with one trial."""
import numpy as np

from .. import darray as D
from .categorical import CategoricalDistribution
from .dirichlet import dirichlet_constant
from .expfam import ExponentialFamily
from .node import Node


class MultinomialDistribution(CategoricalDistribution):

    def __init__(self, trials, categories):
        trials = np.asarray(trials)
        if not issubclass(trials.dtype.type, np.integer):
            raise ValueError("Number of trials must be integer")
        if np.any(trials < 0):
            raise ValueError("Number of trials must be non-negative")
        super().__init__(categories)
        self.N = trials
        self._Nd = D.asarray(trials.astype(np.float64))

    def compute_moments_and_cgf(self, phi, mask=True):
        """multinomial.py:101-121."""
        (p,), g1 = super().compute_moments_and_cgf(phi, mask=mask)
        N = self._Nd
        return [D.mul(p, N.add_trailing(1) if N.ndim else N)], D.mul(g1, N)

    def compute_fixed_moments_and_f(self, x, mask=True):
        """multinomial.py:131-155: the counts themselves; f = log N! - sum log x_k!."""
        x = np.asarray(x)
        if not issubclass(x.dtype.type, np.integer):
            raise ValueError("Counts must be integers")
        if np.any(x < 0):
            raise ValueError("Counts must be non-negative")
        if x.ndim < 1 or x.shape[-1] != self.D:
            raise ValueError("Counts must be given for %d categories" % self.D)
        if np.any(np.sum(x, axis=-1) != self.N):
            raise ValueError("Counts must sum to the number of trials")
        xd = D.asarray(x.astype(np.float64))
        lg = D.gammaln(D.affine(xd, 1.0, 1.0))
        keys = list(range(lg.ndim))
        f = D.sub(D.gammaln(D.affine(self._Nd, 1.0, 1.0)), D.sum_product([lg], [keys], keys[:-1]))
        return [xd], f

    def compute_gradient(self, g, u, phi):
        """multinomial.py:161-218: grad_i = g_i u_i - u_i / N sum_j g_j u_j."""
        ui, g0 = D.asarray(u[0]), D.asarray(g[0])
        gu = D.mul(g0, ui)
        keys = list(range(gu.ndim))
        tot = D.div(D.sum_product([gu], [keys], keys[:-1]), self._Nd)
        return [D.sub(gu, D.mul(ui, tot.add_trailing(1)))]

    def squeeze(self, axis):
        """The distribution without plate axis ``axis`` (for a mixture over that axis, multinomial.py:221-239)."""
        if self.N.ndim < -axis:
            return self
        try:
            N = np.squeeze(self.N, axis)
        except ValueError as err:
            raise ValueError("The number of trials must be constant over a squeezed axis, so the corresponding array "
                             "axis must be singleton. Cannot squeeze axis {0} from a multinomial distribution because "
                             "the number of trials arrays has shape {2}, so the given axis has length {1} != 1. "
                             .format(axis, np.shape(self.N)[axis], np.shape(self.N))) from err
        return MultinomialDistribution(N, self.D)

    def random(self, *phi, plates=None):
        """Host draw with NumPy's global RNG, one ``np.random.multinomial`` per plate like utils/random.py:290-316."""
        logp = np.array(phi[0])
        logp -= np.amax(logp, axis=-1, keepdims=True)
        p = np.exp(logp)
        p = p / np.sum(p, axis=-1, keepdims=True)
        k = p.shape[-1]
        size = tuple(plates) if plates is not None else np.broadcast_shapes(np.shape(self.N), p.shape[:-1])
        n = np.broadcast_to(self.N, size)
        p = np.broadcast_to(p, size + (k,))
        x = np.empty(size + (k,))
        for i in np.ndindex(*size):
            x[i] = np.random.multinomial(n[i], p[i])
        return x.astype(int)


class Multinomial(ExponentialFamily):
    """``Multinomial(n, p, plates=None, name="")`` (multinomial.py:236-319)."""
    moment_kind = "multinomial"
    _guard_zero_times_inf = True

    def __init__(self, n, p, plates=None, name="", initialize=True, plates_multiplier=None):
        if isinstance(p, Node):
            if p.moment_kind != "dirichlet":
                raise ValueError("Expected a Dirichlet-like node")
        else:
            p = dirichlet_constant(p)
        K = p.dims[0][0]
        dist = MultinomialDistribution(n, K)
        from .node import broadcast_plates
        total = broadcast_plates(tuple(p.plates), np.shape(n))
        if plates is not None:
            plates = tuple(int(v) for v in plates)
            if broadcast_plates(total, plates) != plates:
                raise ValueError("The plates %s of the parents are not broadcastable to the given plates %s."
                                 % (total, plates))
            total = plates
        super().__init__(p, dims=((K,),), distribution=dist, plates=total, name=name, initialize=initialize,
                         plates_multiplier=plates_multiplier)

    def __str__(self):
        return "%s ~ Multinomial(p)\n  p = \n%s\n" % (self.name, self.u[0].numpy() / np.asarray(self._distribution.N)[..., None])
