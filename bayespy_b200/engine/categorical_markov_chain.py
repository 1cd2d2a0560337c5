"""Categorical Markov chain node (nodes/categorical_markov_chain.py:80-330): a chain of N discrete states,

    p(z_0 = k) = pi_k,      p(z_n = j | z_{n-1} = i) = [A_{n-1}]_ij,

the latent state sequence of a hidden Markov model when a ``Mixture`` hangs below it.  Moments u = [<z_0> (K,),
<z_n z_{n+1}^T> (N-1, K, K)]; natural parameters phi = [<log pi> + messages, <log A_n> + messages].

The posterior is the forward-backward (alpha-beta) recursion of utils/random.py:357-422 in log space.  Every step is
device work on the plated K x K tables — broadcast adds and the max-shifted, re-normalised softmax / log-sum-exp kernel
of the Categorical node (``bpk_softmax_moments``) — driven by a host loop over time like the reference's; nothing is
read back.  (The recursion is sequential in n by nature; a fused scan kernel is the next step for long chains.)
"""
import numpy as np

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .dirichlet import dirichlet_constant
from .expfam import Distribution, ExponentialFamily
from .node import Deterministic, Node, broadcast_plates


def _softmax_lse(a):
    """(softmax over the last axis, log-sum-exp over the last axis) of a device array; the softmax is normalised twice
    like misc.normalized_exp (misc.py:1388-1401)."""
    a = D.asarray(a).contiguous()
    K = a.shape[-1]
    P = tuple(a.shape[:-1])
    n = int(np.prod(P, dtype=np.int64)) if P else 1
    u, g = DArray.empty(P + (K,)), DArray.empty(P)
    _bpk.get().softmax_moments(a.ptr, n, K, u.ptr, g.ptr)
    return u, D.mul(g, -1.0)


def alpha_beta_recursion(logp0, logP, plates):
    """Forward-backward recursion (utils/random.py:357-422) on device arrays.

    logp0: plates + (K,)      log P(z_0) + log P(y_0 | z_0)           (not normalised)
    logP:  plates + (T, K, K) log P(z_{n+1} | z_n) + log P(y_{n+1} | z_{n+1})
    Returns <z_0> (plates + (K,)), <z_n z_{n+1}^T> (plates + (T, K, K)) and the log-normaliser g (plates)."""
    P = tuple(plates)
    npl = len(P)
    K = logp0.shape[-1]
    T = logP.shape[-3]
    logp0 = D.asarray(logp0).broadcast_to(P + (K,)).contiguous()
    logP = D.asarray(logP).broadcast_to(P + (T, K, K)).contiguous()

    def step(n):
        return logP.slice_axis(npl, n, n + 1).reshape(P + (K, K))
    logalpha = DArray.empty(P + (T, K))
    logbeta = DArray.zeros(P + (T, K))

    def row(buf, n):
        return buf.slice_axis(npl, n, n + 1).reshape(P + (K,))
    D.copy_into(row(logalpha, 0), logp0)
    g = DArray.zeros(P)
    # forward: log P(z_n | y_0..y_n) up to the constants collected in g
    for n in range(1, T):
        v = D.add(row(logalpha, n - 1).add_trailing(1), step(n - 1))              # (.., i, j)
        _, c = _softmax_lse(v.reshape(P + (K * K,)))
        _, s = _softmax_lse(v.swap_last2())                                        # log sum_i, per j
        D.copy_into(row(logalpha, n), D.sub(s, c.add_trailing(1)))
        g = D.sub(g, c)
    v = D.add(row(logalpha, T - 1).add_trailing(1), step(T - 1))
    _, c = _softmax_lse(v.reshape(P + (K * K,)))
    g = D.sub(g, c)
    # backward
    for n in reversed(range(T - 1)):
        v = D.add(row(logbeta, n + 1).expand_dims(-2), step(n + 1))                # (.., i, j)
        _, c = _softmax_lse(v.reshape(P + (K * K,)))
        _, s = _softmax_lse(v)                                                     # log sum_j, per i
        D.copy_into(row(logbeta, n), D.sub(s, c.add_trailing(1)))
    # pairwise marginals: softmax over (i, j) of alpha_n[i] + beta_n[j] + logP_n[i, j]
    v = D.add(D.add(logalpha.add_trailing(1), logbeta.expand_dims(-2)), logP)
    zz, _ = _softmax_lse(v.reshape(P + (T, K * K)))
    zz = zz.reshape(P + (T, K, K))
    first = zz.slice_axis(npl, 0, 1).reshape(P + (K, K))
    pk = [("p", j) for j in range(npl)]
    z0 = D.sum_product([first], [pk + ["i", "j"]], pk + ["i"])
    tot = D.sum_product([z0], [pk + ["i"]], pk)
    z0 = D.div(z0, tot.add_trailing(1))
    return z0, zz, g


class CategoricalMarkovChainDistribution(Distribution):
    zero_times_inf = True          # log-probabilities may be -inf where a moment is 0

    def __init__(self, categories, states):
        self.K, self.N = int(categories), int(states)

    def compute_message_to_parent(self, parent, index, u, u_p0, u_P):
        if index == 0:
            return [u[0]]
        if index == 1:
            return [u[1]]
        raise ValueError("Parent index out of bounds")

    def compute_weights_to_parent(self, index, weights):
        if index == 0:
            return weights
        if index == 1:
            w = np.asarray(weights)
            return w.reshape(w.shape + (1, 1))
        raise ValueError("Parent index out of bounds")

    def compute_phi_from_parents(self, u_p0, u_P, mask=True):
        logP = D.asarray(u_P[0])
        if logP.ndim < 3:
            logP = logP.add_leading(3 - logP.ndim)
        return [u_p0[0], logP.broadcast_to(tuple(logP.shape[:-3]) + (self.N - 1, self.K, self.K))]

    def compute_moments_and_cgf(self, phi, mask=True):
        logp0, logP = D.asarray(phi[0]), D.asarray(phi[1])
        plates = np.broadcast_shapes(tuple(logp0.shape[:-1]), tuple(logP.shape[:-3]))
        z0, zz, g = alpha_beta_recursion(logp0, logP, plates)
        return [z0, zz], g

    def compute_cgf_from_parents(self, u_p0, u_P):
        return D.asarray(0.0)

    def compute_fixed_moments_and_f(self, x, mask=True):
        raise NotImplementedError()

    def plates_to_parent(self, index, plates):
        if index == 0:
            return tuple(plates)
        if index == 1:
            return tuple(plates) + (self.N - 1, self.K)
        raise ValueError("Parent index out of bounds")

    def plates_from_parent(self, index, plates):
        if index == 0:
            return tuple(plates)
        if index == 1:
            return tuple(plates[:-2])
        raise ValueError("Parent index out of bounds")

    def random(self, *phi, plates=None):
        """Host draw with NumPy's global RNG in the reference's order (categorical_markov_chain.py:186-215,
        utils/random.py:247-288): the first state, then every next state given the previous one."""
        from .categorical import CategoricalDistribution
        plates = tuple(plates)
        draw = CategoricalDistribution(self.K).random
        logP = np.array(phi[1]) * np.ones(plates + (1, 1, 1))
        Z = np.zeros(plates + (self.N,), dtype=np.int64)
        Z[..., 0] = draw(np.asarray(phi[0]), plates=plates)
        idx = tuple(np.arange(n).reshape((n,) + (1,) * (len(plates) - i - 1)) for i, n in enumerate(plates))
        for n in range(self.N - 1):
            t = min(n, logP.shape[-3] - 1)
            Z[..., n + 1] = draw(logP[idx + (t, Z[..., n], Ellipsis)], plates=plates)
        return Z


class CategoricalMarkovChain(ExponentialFamily):
    """``CategoricalMarkovChain(pi, A, states=None, plates=None, name="")`` (categorical_markov_chain.py:218-330): pi a
    Dirichlet-like node or (..., K) array, A Dirichlet-like with plates (K,), (..., 1, K) or (..., N-1, K)."""
    moment_kind = "categorical_markov_chain"
    _guard_zero_times_inf = True

    def __init__(self, pi, A, states=None, plates=None, name="", initialize=True):
        def dirichlet_like(p):
            if isinstance(p, Node):
                if p.moment_kind != "dirichlet":
                    raise ValueError("Expected a Dirichlet-like node")
                return p
            return dirichlet_constant(p)
        pi, A = dirichlet_like(pi), dirichlet_like(A)
        K = pi.dims[0][0]
        if len(A.plates) < 2:
            if states is None:
                raise ValueError("Could not infer the length of the Markov chain")
            N = int(states)
        elif A.plates[-2] == 1:
            N = 2 if states is None else int(states)
        else:
            if states is not None and A.plates[-2] + 1 != states:
                raise ValueError("Given length of the Markov chain is inconsistent with the transition probability "
                                 "matrix")
            N = A.plates[-2] + 1
        if tuple(pi.dims) != tuple(A.dims):
            raise ValueError("Initial state probability vector and state transition probability matrix have different "
                             "size")
        if len(A.plates) < 1 or A.plates[-1] != K:
            raise ValueError("Transition probability matrix is not square")
        self.K, self.N = K, N
        super().__init__(pi, A, dims=((K,), (N - 1, K, K)), distribution=CategoricalMarkovChainDistribution(K, N),
                         plates=plates, name=name, initialize=initialize)

    def _to_categorical(self):
        if getattr(self, "_as_categorical", None) is None:
            self._as_categorical = CategoricalMarkovChainToCategorical(self, name=self.name)
        return self._as_categorical


class CategoricalMarkovChainToCategorical(Deterministic):
    """The chain seen as N categorical variables plated over time (categorical_markov_chain.py:333-410): the state
    marginals are <z_0> followed by the column sums of the pairwise moments."""
    moment_kind = "categorical"

    def __init__(self, Z, name=""):
        self.K, self.N = Z.K, Z.N
        super().__init__(Z, dims=((Z.K,),), plates=tuple(Z.plates) + (Z.N,), name=name)

    def _plates_from_parent(self, index):
        return tuple(self.parents[0].plates) + (self.N,)

    def _map_parent_axes(self, index, values):
        return tuple(values) + (1,)

    def _plates_to_parent(self, index):
        return tuple(self.plates[:-1])

    def _weights_to_parent(self, index, mask):
        mask = np.asarray(mask)
        return np.any(mask, axis=-1) if mask.ndim >= 1 else mask

    def _compute_moments(self, u):
        P = tuple(self.parents[0].plates)
        npl = len(P)
        K, N = self.K, self.N
        z0 = D.asarray(u[0]).broadcast_to(P + (K,))
        zz = D.asarray(u[1]).broadcast_to(P + (N - 1, K, K))
        out = DArray.empty(P + (N, K))
        D.copy_into(out.slice_axis(npl, 0, 1), z0.reshape(P + (1, K)))
        pk = [("p", j) for j in range(npl)]
        D.sum_product([zz], [pk + ["n", "i", "j"]], pk + ["n", "j"], out=out.slice_axis(npl, 1, N))
        return [out]

    def message_to_parent(self, index):
        # every child masks its own message, so the time axis simply turns from plate into variable axis
        (m,) = self.message_from_children()
        if m is None:
            return [None, None]
        P = tuple(self.parents[0].plates)
        npl = len(P)
        K, N = self.K, self.N
        m = D.asarray(m)
        if m.ndim < 2:
            m = m.add_leading(2 - m.ndim)
        m = m.broadcast_to(tuple(m.shape[:-2]) + (N, K))
        q = m.ndim - 2
        m0 = m.slice_axis(q, 0, 1).reshape(tuple(m.shape[:q]) + (K,))
        m1 = m.slice_axis(q, 1, N).expand_dims(-2)                   # (.., N-1, 1, K): the same for every previous state
        return [m0, m1]
