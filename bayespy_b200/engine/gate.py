"""``Gate``: deterministic gating of one node over a plate axis by a categorical variable (nodes/gate.py:20-222).

    u_i[...] = sum_k <z_k> u_i(X)[..., k, ...]

Moments and both messages are broadcast product-sums on the device (``D.sum_product``, i.e. ``bpk_sum_multiply``):
the gated plate is just one more summation key, so nothing is moved or copied (the reference moves the axis to the end
and calls ``misc.sum_product``)."""
import numpy as np

from .. import darray as D
from .categorical import categorical_constant
from .gaussian import dense
from .node import Deterministic, Node


def _pk(n):
    return [("p", j) for j in range(n, 0, -1)]


class Gate(Deterministic):

    def __init__(self, Z, X, gated_plate=-1, moments=None, name=""):
        if gated_plate >= 0:
            raise ValueError("Cluster plate must be negative integer")
        self.gated_plate = int(gated_plate)
        if moments is not None:
            from . import moments as _moments
            X = _moments.ensure(X, moments)              # e.g. an array with ``moments=GaussianMoments(())`` (gate.py:56-62)
        if not isinstance(X, Node):
            raise ValueError("X must be a node or moments should be provided")
        if len(X.plates) < abs(gated_plate):
            raise ValueError("The gated node does not have a plate axis is gated")
        self.K = K = int(X.plates[gated_plate])
        if isinstance(Z, Node) and hasattr(Z, "_to_categorical"):
            Z = Z._to_categorical()
        if not isinstance(Z, Node):
            Z = categorical_constant(Z, K)
        if Z.moment_kind != "categorical" or tuple(Z.dims) != ((K,),):
            raise ValueError("Inconsistent number of clusters")
        self.moment_kind = X.moment_kind
        super().__init__(Z, X, dims=X.dims, name=name)

    # ---- plates (gate.py:170-205) ----------------------------------------------------------------------------------
    def _plates_from_parent(self, index):
        plates = list(self.parents[index].plates)
        if index == 1 and len(plates) >= abs(self.gated_plate):
            plates.pop(self.gated_plate)
        return tuple(plates)

    def _plates_to_parent(self, index):
        plates = list(self.plates)
        if index == 1:
            plates.insert(len(plates) + self.gated_plate + 1, self.K)
        return tuple(plates)

    def _weights_to_parent(self, index, mask):
        if index == 0:
            return mask
        mask = np.asarray(mask)
        return np.expand_dims(mask, axis=self.gated_plate) if mask.ndim >= abs(self.gated_plate) else mask

    # ---- keys: result plates p_n..p_1, X plates the same with "k" inserted at the gated position -------------------
    def _keys(self, i):
        npl = len(self.plates)
        p = _pk(npl)
        px = list(p)
        px.insert(npl + self.gated_plate + 1, "k")
        d = [("d", i, j) for j in range(len(self.dims[i]))]
        return p, px, d

    @staticmethod
    def _tail(keys, arr):
        """Right-aligned key list for an array that may lack leading (broadcast) axes."""
        return keys[len(keys) - arr.ndim:] if arr.ndim <= len(keys) else None

    def _compute_moments(self, u_Z, u_X):
        z = D.asarray(u_Z[0])
        out = []
        for i in range(len(self.dims)):
            p, px, d = self._keys(i)
            x = D.asarray(dense(u_X[i]))
            out.append(D.sum_product([z, x], [self._tail(p + ["k"], z), self._tail(px + d, x)], p + d,
                                     sizes=self._sizes(i)))
        return out

    def _sizes(self, i):
        p, px, d = self._keys(i)
        s = {k: n for k, n in zip(p, self.plates)}
        s.update({k: n for k, n in zip(d, self.dims[i])})
        s["k"] = self.K
        return s

    def _compute_message_to_parent(self, index, m_child, u_Z, u_X):
        if index == 0:
            # to Z: <child message, moments of X> summed over the variable axes, gated plate last (gate.py:109-133)
            m0 = None
            for i in range(len(self.dims)):
                if m_child[i] is None:
                    continue
                p, px, d = self._keys(i)
                c = D.asarray(m_child[i])
                x = D.asarray(dense(u_X[i]))
                t = D.sum_product([c, x], [self._tail(p + d, c), self._tail(px + d, x)], p + ["k"], sizes=self._sizes(i))
                m0 = t if m0 is None else D.add(m0, t)
            return [m0]
        if index == 1:
            # to X: <z_k> times the child message, gated plate at its place (gate.py:135-165)
            z = D.asarray(u_Z[0])
            out = []
            for i in range(len(self.dims)):
                if m_child[i] is None:
                    out.append(None)
                    continue
                p, px, d = self._keys(i)
                c = D.asarray(m_child[i])
                out.append(D.sum_product([z, c], [self._tail(p + ["k"], z), self._tail(p + d, c)], px + d,
                                         sizes=self._sizes(i)))
            return out
        raise ValueError("Invalid parent index")


def Choose(z, *nodes):
    """Plate elements picked from several nodes by a categorical variable (gate.py:219-250): a ``Gate`` over the
    ``Concatenate`` of the nodes along a new last plate axis.

        z = [0, 0, 2, 1];  Choose(z, x0, x1, x2).get_moments()[0]  ->  [<x0>, <x0>, <x2>, <x1>]"""
    from .concatenate import Concatenate
    from . import moments
    z = moments.ensure(z, moments._categorical_moments()(len(nodes)))
    return Gate(z, Concatenate(*[n[..., None] for n in nodes]))
