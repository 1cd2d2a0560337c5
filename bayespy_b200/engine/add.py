"""``Add``: the sum of Gaussian nodes that are independent in the posterior approximation (nodes/add.py:19-135).

    <y> = sum_i <x_i>,     <y y^T> = sum_i <x_i x_i^T> + sum_{i != j} <x_i> <x_j>^T
    message to x_i:  [m1 + 2 M2 sum_{j != i} <x_j>,  M2]

Moments and messages are broadcast adds and one ``bpk_sum_multiply`` contraction per parent on the device; the variable
block is flattened so any number of variable axes is handled alike."""
import numpy as np

from .. import darray as D
from .gaussian import dense, ensure_gaussian
from .node import Deterministic, Node


class Add(Deterministic):
    moment_kind = "gaussian"

    def __init__(self, *nodes, plates=None, name=""):
        if len(nodes) < 2:
            raise ValueError("Give at least two parents")
        ndim = None
        for n in nodes:
            if isinstance(n, Node):
                if hasattr(n, "_to_gaussian") and n.moment_kind != "gaussian":
                    n = n._to_gaussian()
                ndim = len(n.dims[0])
                break
        if ndim is None:
            raise ValueError("At least one parent must be a node")
        nodes = [ensure_gaussian(n, ndim) for n in nodes]
        for a, b in zip(nodes[:-1], nodes[1:]):
            if tuple(a.dims) != tuple(b.dims):
                raise ValueError("Nodes do not have identical shapes")
        self.shape = tuple(nodes[0].dims[0])
        self.nd = len(self.shape)
        self.K = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        super().__init__(*nodes, dims=nodes[0].dims, plates=plates, name=name)

    def _outer(self, a, b):
        nd = self.nd
        if nd == 0:
            return D.mul(a, b)
        return D.mul(a.add_trailing(nd), b.reshape(tuple(b.shape[:b.ndim - nd]) + (1,) * nd + tuple(b.shape[b.ndim - nd:])))

    def _compute_moments(self, *u_parents):
        x = [D.asarray(u[0]) for u in u_parents]
        u0, u1 = x[0], D.asarray(dense(u_parents[0][1]))
        for xi, u in zip(x[1:], u_parents[1:]):
            u0 = D.add(u0, xi)
            u1 = D.add(u1, D.asarray(dense(u[1])))
        for i in range(len(x)):
            for j in range(i + 1, len(x)):
                u1 = D.add(u1, D.add(self._outer(x[i], x[j]), self._outer(x[j], x[i])))
        return [u0, u1]

    def _compute_message_to_parent(self, index, m, *u_parents):
        m0, m1 = m
        if m1 is None:
            return [m0, None]
        others = [D.asarray(u[0]) for k, u in enumerate(u_parents) if k != index]
        s = others[0]
        for o in others[1:]:
            s = D.add(s, o)
        nd, K = self.nd, self.K
        m1 = D.asarray(m1)
        if nd == 0:
            t = D.mul(D.mul(m1, s), 2.0)
        else:
            # flatten the variable block: (.., K, K) times (.., K)
            M = m1.contiguous().reshape(tuple(m1.shape[:m1.ndim - 2 * nd]) + (K, K))
            v = s.contiguous().reshape(tuple(s.shape[:s.ndim - nd]) + (K,))
            npl = max(M.ndim - 2, v.ndim - 1)
            pk = [("p", j) for j in range(npl, 0, -1)]
            t = D.sum_product([M, v], [pk[npl - (M.ndim - 2):] + ["i", "j"], pk[npl - (v.ndim - 1):] + ["j"]],
                              pk + ["i"], scale=2.0)
            t = t.reshape(tuple(t.shape[:-1]) + self.shape)
        return [t if m0 is None else D.add(D.asarray(m0), t), m1]
