"""Wishart node on device (replaces the array math of nodes/wishart.py:118-306).

Lambda ~ Wishart(n, V) over D x D SPD matrices; phi = [-V/2, n/2];
u = [<Lambda>, <log|Lambda|>].  The per-plate Cholesky / inverse / log-det
(linalg.py via wishart.py:165-188) and the multivariate digamma / log-gamma are
one fused kernel, ``bpk_wishart_moments``.
"""
import numpy as np
import scipy.special as sp

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .expfam import Distribution, ExponentialFamily
from .node import Constant, Node


def wishart_constant(Lam):
    """[Lambda, log|Lambda|] of a fixed SPD matrix stack (WishartMoments.compute_fixed_moments,
    wishart.py:55-75); the factorisation runs on device."""
    from ..utils import linalg
    Lam = np.asarray(Lam, dtype=np.float64)
    if Lam.ndim < 2 or Lam.shape[-1] != Lam.shape[-2]:
        raise ValueError("Values for Wishart distribution must be square matrices, thus the two last "
                         "axes must have equal length.")
    Ld = D.asarray(Lam)
    ldet = linalg.chol_logdet(linalg.chol(Ld))
    Dm = Lam.shape[-1]
    return Constant("wishart", [Ld, ldet], dims=((Dm, Dm), ()), plates=Lam.shape[:-2], value=Lam)


def wishart_prior_constant(n, d):
    """[n, log Gamma_d(n/2)]  (WishartPriorMoments, wishart.py:23-42; host special function of a constant)."""
    n = np.asarray(n, dtype=np.float64)
    nd = D.asarray(n)
    return Constant("wishart_prior", [nd, D.multigammaln(D.mul(nd, 0.5), d)], dims=((), ()), plates=n.shape,
                    value=n)


def ensure_wishart(V):
    if isinstance(V, Node):
        if V.moment_kind != "wishart":
            raise ValueError("Expected a Wishart-like node, got %s" % type(V).__name__)
        return V
    return wishart_constant(V)


class WishartDistribution(Distribution):

    def compute_message_to_parent(self, parent, index, u_self, u_n, u_V):
        """wishart.py:136-147."""
        if index == 0:
            raise NotImplementedError("Message from Wishart to degrees of freedom parameter (first parent) "
                                      "not yet implemented")
        elif index == 1:
            return [D.mul(u_self[0], -0.5), D.mul(u_n[0], 0.5)]
        raise ValueError("Invalid parent index {0}".format(index))

    def compute_phi_from_parents(self, u_n, u_V, mask=True):
        """wishart.py:149-163."""
        return [D.mul(u_V[0], -0.5), D.mul(u_n[0], 0.5)]

    def compute_cgf_from_parents(self, u_n, u_V):
        """n/2 log|V| - nk/2 log 2 - log Gamma_k(n/2)   (wishart.py:190-205)."""
        n, gammaln_n = u_n
        V, logdet_V = u_V
        k = V.shape[-1]
        t = D.mul(D.mul(n, logdet_V), 0.5)
        return D.sub(D.axpby(1.0, t, -0.5 * k * float(np.log(2)), n), gammaln_n)

    def compute_moments_and_cgf(self, phi, mask=True):
        """wishart.py:165-188 as one kernel."""
        be = _bpk.get()
        p0, p1 = [D.asarray(v) for v in phi]
        Dm = p0.shape[-1]
        P = tuple(np.broadcast_shapes(p0.shape[:-2], p1.shape))
        n = int(np.prod(P, dtype=np.int64)) if P else 1
        if (1,) * (len(P) - (p0.ndim - 2)) + tuple(p0.shape[:-2]) != P:
            p0 = p0.broadcast_to(P + (Dm, Dm))
        p0c = p0.contiguous()
        n1 = 1 if p1.size == 1 else n
        if n1 == n and p1.size != 1 and (1,) * (len(P) - p1.ndim) + tuple(p1.shape) != P:
            p1 = p1.broadcast_to(P)
        p1c = p1.contiguous()
        u0, u1, g = DArray.empty(P + (Dm, Dm)), DArray.empty(P), DArray.empty(P)
        be.wishart_moments(p0c.ptr, p1c.ptr, n1, n, Dm, u0.ptr, u1.ptr, g.ptr, True)
        return [u0, u1], g

    def compute_fixed_moments_and_f(self, Lambda, mask=True):
        """wishart.py:207-225."""
        c = wishart_constant(Lambda)
        k = np.shape(Lambda)[-1]
        return c.u, D.mul(c.u[1], -(k + 1) / 2)


class Wishart(ExponentialFamily):
    """``Wishart(n, V, plates=None, name="")`` as in the reference (wishart.py:228-306)."""
    moment_kind = "wishart"

    def __init__(self, n, V, plates=None, name="", initialize=True):
        V = ensure_wishart(V)
        Dm = V.dims[0][-1]
        if isinstance(n, Node):
            if n.moment_kind != "wishart_prior":
                raise ValueError("Degrees of freedom must be a fixed value")
        else:
            n = wishart_prior_constant(n, Dm)
        super().__init__(n, V, dims=((Dm, Dm), ()), distribution=WishartDistribution(), plates=plates,
                         name=name, initialize=initialize)

    def __str__(self):
        n = 2 * self.phi[1].numpy()
        A = 0.5 * self.u[0].numpy() / self.phi[1].numpy()[..., None, None]
        return "%s ~ Wishart(n, A)\n  n =\n%s\n  A =\n%s\n" % (self.name, n, A)
