"""``ConcatGaussian``: independent Gaussian vectors stacked into one longer vector (nodes/concat_gaussian.py:17-95).

    <y> = [<x_1>; <x_2>; ...],      <y y^T> blocks: <x_i x_i^T> on the diagonal, <x_i> <x_j>^T off it
    message to x_i:  [m1_i + 2 sum_{j != i} M2_ij <x_j>,  M2_ii]

Blocks are written through strided views of one device array; the messages are views of the child's message."""
import numpy as np

from .. import darray as D
from ..darray import DArray
from .gaussian import dense, ensure_gaussian
from .node import Deterministic


class ConcatGaussian(Deterministic):
    moment_kind = "gaussian"

    def __init__(self, *nodes, plates=None, name=""):
        nodes = [ensure_gaussian(n, 1) for n in nodes]
        if any(len(n.dims[0]) != 1 for n in nodes):
            raise ValueError("Input nodes must be (Gaussian) vectors")
        self.slices = tuple(int(v) for v in np.cumsum([0] + [n.dims[0][0] for n in nodes]))
        Dm = self.slices[-1]
        super().__init__(*nodes, dims=((Dm,), (Dm, Dm)), plates=plates, name=name)

    def _compute_moments(self, *u_nodes):
        xs = [D.asarray(u[0]) for u in u_nodes]
        xxs = [D.asarray(dense(u[1])) for u in u_nodes]
        P = tuple(np.broadcast_shapes(*[tuple(x.shape[:-1]) for x in xs], *[tuple(a.shape[:-2]) for a in xxs]))
        npl, r, Dm = len(P), self.slices, self.slices[-1]
        x = DArray.empty(P + (Dm,))
        xx = DArray.zeros(P + (Dm, Dm))
        for i, (xi, xxi) in enumerate(zip(xs, xxs)):
            D.copy_into(x.slice_axis(npl, r[i], r[i + 1]), xi.broadcast_to(P + (r[i + 1] - r[i],)))
            rows = xx.slice_axis(npl, r[i], r[i + 1])
            D.copy_into(rows.slice_axis(npl + 1, r[i], r[i + 1]), xxi.broadcast_to(P + (r[i + 1] - r[i],) * 2))
            for j in range(i):
                xi_xj = D.mul(xi.add_trailing(1), xs[j].expand_dims(-2))
                D.copy_into(rows.slice_axis(npl + 1, r[j], r[j + 1]), xi_xj)
                D.copy_into(xx.slice_axis(npl, r[j], r[j + 1]).slice_axis(npl + 1, r[i], r[i + 1]), xi_xj.swap_last2())
        return [x, xx]

    def _compute_message_to_parent(self, i, m, *u_nodes):
        r = self.slices
        m0, m1 = (None if v is None else D.asarray(v) for v in m)
        out0 = None if m0 is None else m0.slice_axis(m0.ndim - 1, r[i], r[i + 1])
        if m1 is None:
            return [out0, None]
        rows = m1.slice_axis(m1.ndim - 2, r[i], r[i + 1])
        for j, u in enumerate(u_nodes):
            if j == i:
                continue
            Mij = rows.slice_axis(rows.ndim - 1, r[j], r[j + 1])
            xj = D.asarray(u[0])
            npl = max(Mij.ndim - 2, xj.ndim - 1)
            pk = [("p", k) for k in range(npl, 0, -1)]
            t = D.sum_product([Mij, xj], [pk[npl - (Mij.ndim - 2):] + ["a", "b"], pk[npl - (xj.ndim - 1):] + ["b"]],
                              pk + ["a"], scale=2.0)
            out0 = t if out0 is None else D.add(out0, t)
        return [out0, rows.slice_axis(rows.ndim - 1, r[i], r[i + 1])]
