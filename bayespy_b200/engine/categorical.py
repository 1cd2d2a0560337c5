"""Categorical node on device (replaces nodes/categorical.py:87-201 and the softmax of
nodes/multinomial.py:101-121 / misc.py:1366-1401).  One-hot encoding of integer labels is
index work and is bit-exact (``bpk_one_hot``)."""
import numpy as np

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .dirichlet import dirichlet_constant
from .expfam import Distribution, ExponentialFamily
from .node import Constant, Node


def one_hot(x, K):
    x = np.asarray(x)
    if not np.issubdtype(x.dtype, np.integer):
        if np.any(x != np.floor(x)):
            raise ValueError("Values must be integers")
        x = x.astype(np.int64)
    if np.any(x < 0) or np.any(x >= K):
        raise ValueError("Invalid category index")
    lab = DArray.from_numpy(np.ascontiguousarray(x, dtype=np.int64).reshape(x.shape), "i8")
    out = DArray.empty(tuple(x.shape) + (K,))
    _bpk.get().one_hot(lab.ptr, int(x.size), K, out.ptr, True)
    return out


def categorical_constant(x, K):
    x = np.asarray(x)
    return Constant("categorical", [one_hot(x, K)], dims=((K,),), plates=x.shape, value=x)


class CategoricalMoments:
    """``CategoricalMoments(K)`` as the reference exposes it for ``Constant`` (categorical.py:20-90): integer class
    labels become one-hot moments on the device (bit-exact)."""
    kind = "categorical"

    def __init__(self, categories):
        if not isinstance(categories, (int, np.integer)) or categories < 0:
            raise ValueError("Number of categories must be a non-negative integer")
        self.categories = int(categories)

    def fixed_moments(self, x):
        x = np.asarray(x)
        if not issubclass(x.dtype.type, np.integer):
            raise ValueError("Values must be integers")
        if np.any(x < 0) or np.any(x >= self.categories):
            raise ValueError("Invalid category index")
        K = self.categories
        return "categorical", [one_hot(x, K)], ((K,),), x.shape, x


class CategoricalDistribution(Distribution):
    zero_times_inf = True          # log-probabilities may be -inf where the one-hot moment is 0

    def compute_gradient(self, g, u, phi):
        """multinomial.py:161-218 with one trial: the Fisher information of the softmax is diag(p) - p p^T, so
        grad_i = g_i p_i - p_i sum_j g_j p_j."""
        p, g0 = D.asarray(u[0]), D.asarray(g[0])
        gp = D.mul(g0, p)
        keys = list(range(gp.ndim))
        tot = D.sum_product([gp], [keys], keys[:-1])
        return [D.sub(gp, D.mul(p, tot.add_trailing(1)))]

    def __init__(self, categories):
        if not isinstance(categories, (int, np.integer)):
            raise ValueError("Number of categories must be integer")
        if categories < 0:
            raise ValueError("Number of categoriess must be non-negative")
        self.D = int(categories)

    def compute_message_to_parent(self, parent, index, u, u_p):
        if index == 0:
            return [u[0]]
        raise ValueError("Index out of bounds")

    def compute_phi_from_parents(self, u_p, mask=True):
        return [u_p[0]]

    def compute_cgf_from_parents(self, u_p):
        return D.asarray(0.0)

    def compute_moments_and_cgf(self, phi, mask=True):
        """softmax with the reference's max-shift and second renormalisation; g = -logsumexp."""
        phi = [D.asarray(v) for v in phi]         # the protocol also takes host arrays (seam 2)
        p = phi[0].contiguous()
        K = p.shape[-1]
        P = tuple(p.shape[:-1])
        n = int(np.prod(P, dtype=np.int64)) if P else 1
        u, g = DArray.empty(P + (K,)), DArray.empty(P)
        _bpk.get().softmax_moments(p.ptr, n, K, u.ptr, g.ptr)
        return [u], g

    def compute_fixed_moments_and_f(self, x, mask=True):
        return [one_hot(x, self.D)], 0.0

    def random(self, *phi, plates=None):
        """Host draw, same call pattern as categorical.py:117-124 + utils/random.py:247-288."""
        logp = np.array(phi[0])
        logp -= np.amax(logp, axis=-1, keepdims=True)
        p = np.exp(logp)
        size = tuple(plates) if plates is not None else p.shape[:-1]
        p = p / np.sum(p, axis=-1, keepdims=True)
        P = np.cumsum(p, axis=-1)
        x = np.random.rand(*size)
        P = P * np.ones(size + (p.shape[-1],))
        if size == ():
            return int(np.searchsorted(P, x))
        z = np.zeros(size)
        for ind in np.ndindex(*size):
            z[ind] = np.searchsorted(P[ind], x[ind])
        return z.astype(int)


class Categorical(ExponentialFamily):
    """``Categorical(p, plates=None, name="")`` (categorical.py:127-201)."""
    moment_kind = "categorical"
    _guard_zero_times_inf = True

    def __init__(self, p, plates=None, name="", initialize=True, plates_multiplier=None, **kwargs):
        if isinstance(p, Node):
            if p.moment_kind != "dirichlet":
                raise ValueError("Expected a Dirichlet-like node")
        else:
            p = dirichlet_constant(p)               # invalid probabilities are refused before anything else
        if plates is not None:
            from .node import is_subshape
            if not is_subshape(tuple(p.plates), tuple(plates)):
                raise ValueError("The plates %s of the parents are not broadcastable to the given plates %s."
                                 % (tuple(p.plates), tuple(plates)))
        if kwargs:
            raise TypeError("Categorical got unexpected keyword arguments: %s" % ", ".join(sorted(kwargs)))
        K = p.dims[0][0]
        super().__init__(p, dims=((K,),), distribution=CategoricalDistribution(K), plates=plates, name=name,
                         initialize=initialize, plates_multiplier=plates_multiplier)

    def __str__(self):
        return "%s ~ Categorical(p)\n  p = \n%s\n" % (self.name, self.u[0].numpy())
