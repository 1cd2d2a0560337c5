"""``Take``: pick elements / sub-arrays along a plate axis (nodes/take.py:13-140) — ``np.take`` on the plates.

Moments are gathered on the device by integer index (``bpk_take``, pure data movement: bit-exact); the message to the
parent is the inverse with accumulation (``misc.put_simple``, misc.py:549-585): every parent plate receives the sum of
the messages of the child plates that picked it, added in increasing child order (``bpk_put_add`` with a CSR of the
index array built once on the host: deterministic, no atomics)."""
import numpy as np

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .gaussian import dense
from .node import Deterministic


class Take(Deterministic):

    def __init__(self, node, indices, plate_axis=-1, name=""):
        self._indices = np.array(indices)
        self._plate_axis = plate_axis
        if not isinstance(plate_axis, (int, np.integer)):
            raise ValueError("Plate axis must be integer")
        if plate_axis >= 0:
            raise ValueError("plate_axis must be negative index")
        if plate_axis < -len(node.plates):
            raise ValueError("plate_axis out of bounds")
        if not issubclass(self._indices.dtype.type, np.integer):
            raise ValueError("Indices must be integers")
        self._original_length = L = int(node.plates[plate_axis])
        if np.any(self._indices < -L) or np.any(self._indices >= L):
            raise ValueError("Index out of bounds")
        self.moment_kind = node.moment_kind
        flat = np.where(self._indices < 0, self._indices + L, self._indices).astype(np.int64).reshape(-1)
        self._J = int(flat.size)
        order = np.argsort(flat, kind="stable").astype(np.int64)
        start = np.searchsorted(flat[order], np.arange(L + 1), side="left").astype(np.int64)
        self._flat = flat
        self._dev = None
        self._host_csr = (order, start)
        super().__init__(node, dims=node.dims, name=name)

    # the index arrays live on the device once the first kernel needs them
    def _index_arrays(self):
        if self._dev is None:
            order, start = self._host_csr
            self._dev = (DArray.from_numpy(self._flat, "i8"), DArray.from_numpy(order, "i8"), DArray.from_numpy(start, "i8"))
        return self._dev

    # ---- plates (take.py:91-124) ---------------------------------------------------------------------------------
    def _plates_from_parent(self, index):
        pp = tuple(self.parents[index].plates)
        ax = self._plate_axis
        plates = pp[:ax] + tuple(np.shape(self._indices))
        if ax != -1:
            plates = plates + pp[ax + 1:]
        return plates

    def _plates_to_parent(self, index):
        plates = tuple(self.plates)
        nd = np.ndim(self._indices)
        end_before = self._plate_axis - nd + 1
        start_after = self._plate_axis + 1
        if end_before == 0:
            return plates + (self._original_length,)
        if start_after == 0:
            return plates[:end_before] + (self._original_length,)
        return plates[:end_before] + (self._original_length,) + plates[start_after:]

    def _weights_to_parent(self, index, mask):
        """misc.put_simple on the host mask: how many active child plates picked each parent plate."""
        w = np.broadcast_to(np.asarray(mask, dtype=np.float64), tuple(self.plates))
        nd = np.ndim(self._indices)
        ax = len(self.plates) + self._plate_axis - nd + 1          # first of the axes the indices created
        pre = int(np.prod(self.plates[:ax], dtype=np.int64))
        post = int(np.prod(self.plates[ax + nd:], dtype=np.int64))
        w3 = np.ascontiguousarray(w).reshape(pre, self._J, post)
        out = np.zeros((pre, self._original_length, post))
        np.add.at(out, (slice(None), self._flat, slice(None)), w3)
        return out.reshape(tuple(self.plates[:ax]) + (self._original_length,) + tuple(self.plates[ax + nd:]))

    # ---- moments and messages -------------------------------------------------------------------------------------
    def _split(self, plates_full, ndim, n_taken_axes, taken_len):
        ax = len(plates_full) + self._plate_axis - n_taken_axes + 1
        pre = int(np.prod(plates_full[:ax], dtype=np.int64))
        post = int(np.prod(plates_full[ax + n_taken_axes:], dtype=np.int64))
        return ax, pre, post

    def _compute_moments(self, u_parent):
        be = _bpk.get()
        idx, _, _ = self._index_arrays()
        pp = tuple(self.parents[0].plates)
        out = []
        for ui, dims in zip(u_parent, self.dims):
            ui = D.asarray(dense(ui))
            full = pp + tuple(dims)
            src = ui.broadcast_to(full).contiguous()                 # a broadcast (unit / missing) taken axis gets its length
            ax = len(pp) + self._plate_axis
            pre = int(np.prod(pp[:ax], dtype=np.int64))
            post = int(np.prod(full[ax + 1:], dtype=np.int64))
            res = DArray.empty(tuple(self.plates) + tuple(dims))
            be.take(src.ptr, pre, self._original_length, post, idx.ptr, self._J, res.ptr)
            out.append(res)
        return out

    def _compute_message_to_parent(self, index, m_child, u_parent):
        be = _bpk.get()
        _, order, start = self._index_arrays()
        nd = np.ndim(self._indices)
        plates = tuple(self.plates)
        to_plates = self._plates_to_parent(0)
        msg = []
        for mi, dims in zip(m_child, self.dims):
            if mi is None:
                msg.append(None)
                continue
            full = plates + tuple(dims)
            src = D.asarray(mi).broadcast_to(full).contiguous()
            ax = len(plates) + self._plate_axis - nd + 1
            pre = int(np.prod(plates[:ax], dtype=np.int64))
            post = int(np.prod(full[ax + nd:], dtype=np.int64))
            res = DArray.empty(tuple(to_plates) + tuple(dims))
            be.put_add(src.ptr, pre, self._J, post, order.ptr, start.ptr, self._original_length, res.ptr)
            msg.append(res)
        return msg
