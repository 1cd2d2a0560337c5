"""``Concatenate``: nodes of one moment kind joined along a plate axis (nodes/concatenate.py:14-190), like
``numpy.concatenate`` for the plates.  Moments are copied side by side into one device array (a parent that broadcasts
over other plates is expanded only along those the result needs); a message to a parent is the strided view of its
stretch of the axis, so nothing is copied on the way up."""
import numpy as np

from .. import darray as D
from ..darray import DArray
from .node import Deterministic, Node, broadcast_plates


class Concatenate(Deterministic):

    def __init__(self, *nodes, axis=-1, plates=None, name=""):
        if axis >= 0:
            raise ValueError("Currently, only negative axis indeces are allowed.")
        first = next((n for n in nodes if isinstance(n, Node)), None)
        if first is None:
            raise ValueError("Couldn't determine parent moments")
        from . import moments
        tag = first._moments
        try:
            nodes = [n if isinstance(n, Node) and n.moment_kind == first.moment_kind else moments.ensure(n, tag)
                     for n in nodes]
        except (ValueError, KeyError, NotImplementedError):
            raise ValueError("Parents have different moments")
        for n in nodes:
            if tuple(n.dims) != tuple(first.dims):
                raise ValueError("Parents have different dimensionalities")
            if len(n.plates) < -axis:
                raise ValueError("Every parent needs the plate axis %d" % axis)
        self._axis = int(axis)
        self.moment_kind = first.moment_kind
        self._lengths = [int(n.plates[axis]) for n in nodes]
        self._indices = np.concatenate(([0], np.cumsum(self._lengths))).astype(int)
        super().__init__(*nodes, dims=first.dims, plates=plates, name=name)

    # parents are kept apart along the axis, so they need not be independent (concatenate.py:77-83)
    def _check_independent_parents(self):
        pass

    def _ids(self):
        return list(dict.fromkeys(super()._ids()))

    def _plates_from_parent(self, index):
        p = list(self.parents[index].plates)
        p[self._axis] = int(self._indices[-1])
        return tuple(p)

    def _plates_to_parent(self, index):
        p = list(self.plates)
        p[self._axis] = self._lengths[index]
        return tuple(p)

    def _plates_multiplier_from_parent(self, index):
        m = tuple(getattr(self.parents[index], "plates_multiplier", ()))
        if any(v != 1 for v in m):
            raise ValueError("Concatenation node does not support plate multipliers.")
        return ()

    def _weights_to_parent(self, index, mask):
        mask = np.asarray(mask)
        ax = self._axis
        if mask.ndim >= -ax and mask.shape[ax] > 1:
            return np.take(mask, range(self._indices[index], self._indices[index + 1]), axis=ax)
        return mask

    def _compute_moments(self, *u_parents):
        out = []
        for i, dims in enumerate(self.dims):
            nd = len(dims)
            ax = self._axis - nd                                  # the concatenated axis among the moment's axes
            parts = [D.asarray(u[i].materialize() if hasattr(u[i], "materialize") else u[i]) for u in u_parents]
            need = max(max(p.ndim for p in parts), -ax)
            parts = [p.add_leading(need - p.ndim) for p in parts]
            shapes = []
            for p in parts:
                sh = list(p.shape)
                sh[ax] = 1
                shapes.append(tuple(sh))
            bc = list(np.broadcast_shapes(*shapes))
            total = list(bc)
            total[ax] = int(self._indices[-1])
            res = DArray.empty(tuple(total))
            pos = need + ax
            for k, (p, n) in enumerate(zip(parts, self._lengths)):
                tgt = list(bc)
                tgt[ax] = n
                view = res.slice_axis(pos, int(self._indices[k]), int(self._indices[k + 1]))
                D.copy_into(view, p.broadcast_to(tuple(tgt)))
            out.append(res)
        return out

    def _compute_message_to_parent(self, index, m, *u_parents):
        msg = []
        lo, hi = int(self._indices[index]), int(self._indices[index + 1])
        for mi, dims in zip(m, self.dims):
            if mi is None:
                msg.append(None)
                continue
            mi = D.asarray(mi)
            ax = self._axis - len(dims)
            if mi.ndim >= -ax and mi.shape[ax] > 1:
                mi = mi.slice_axis(mi.ndim + ax, lo, hi)
            msg.append(mi)
        return msg
