"""Poisson and Exponential nodes (nodes/poisson.py:46-140, exponential.py:18-60): elementwise exponential-family
updates on device arrays.  Poisson: rate from a gamma-like parent, phi = [<log lambda>], u = [e^phi], g = -e^phi.
Exponential(l) is Gamma(1, l)."""
import numpy as np

from .. import darray as D
from .expfam import Distribution, ExponentialFamily
from .gamma import Gamma
from .gaussian import ensure_gamma


class PoissonDistribution(Distribution):

    def compute_message_to_parent(self, parent, index, u, u_lambda):
        """[-1, x] for the gamma moments [lambda, log lambda] of the rate (poisson.py:52-63)."""
        if index != 0:
            raise ValueError("Index out of bounds")
        return [D.asarray(-1.0), u[0]]

    def compute_phi_from_parents(self, u_lambda, mask=True):
        return [u_lambda[1]]

    def compute_moments_and_cgf(self, phi, mask=True):
        u0 = D.exp(D.asarray(phi[0]))
        return [u0], D.mul(u0, -1.0)

    def compute_cgf_from_parents(self, u_lambda):
        return D.mul(u_lambda[0], -1.0)

    def compute_fixed_moments_and_f(self, x, mask=True):
        x = np.asarray(x)
        if not issubclass(x.dtype.type, np.integer):
            raise ValueError("Count not integer")
        if np.any(x < 0):
            raise ValueError("Counts must be non-negative")
        xd = D.asarray(x.astype(np.float64))
        return [xd], D.mul(D.gammaln(D.affine(xd, 1.0, 1.0)), -1.0)

    def random(self, *phi, plates=None):
        return np.random.poisson(np.exp(np.asarray(phi[0])), size=plates)


class Poisson(ExponentialFamily):
    """``Poisson(l, plates=None, name="")`` (poisson.py:104-140); ``l`` gamma-like or an array of rates."""
    moment_kind = "poisson"

    def __init__(self, l, plates=None, name="", initialize=True, plates_multiplier=None):
        super().__init__(ensure_gamma(l), dims=((),), distribution=PoissonDistribution(), plates=plates, name=name,
                         initialize=initialize, plates_multiplier=plates_multiplier)

    def __str__(self):
        return "%s ~ Poisson(lambda)\n  lambda = \n%s\n" % (self.name, self.u[0].numpy())


class Exponential(Gamma):
    """``Exponential(l)`` = ``Gamma(1, l)`` (exponential.py:18-60)."""

    def __init__(self, l, plates=None, name="", initialize=True):
        super().__init__(1, l, plates=plates, name=name, initialize=initialize)
