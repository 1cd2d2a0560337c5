"""SumMultiply / Dot: einsum-like deterministic node over Gaussian parents
(replaces nodes/dot.py:19-644).

All three einsums of the reference — the mean (:355), the second moment (:403)
and "THE BEEF" message contraction (:581) — become ``sum_product`` calls, i.e.
``bpk_sum_multiply`` launches with broadcast strides; nothing is materialised on
the host and no operand is expanded to the full plate space.
"""
import numpy as np

from .. import darray as D
from .gaussian import dense, ensure_gaussian
from .node import Constant, Deterministic, Node


def _parse(args):
    args = list(args)
    if len(args) < 2:
        raise ValueError("Not enough inputs")
    if isinstance(args[0], str):
        spec = "".join(args[0].split())
        parts = spec.split("->")
        if len(parts) > 2:
            raise ValueError("The string contains too many ->")
        ins = parts[0].split(",")
        nodes = args[1:]
        if len(ins) != len(nodes):
            raise ValueError("Number of given input nodes is different from the input keys in the string")
        keysets = [list(s) for s in ins]
        keys_out = list(parts[1]) if len(parts) == 2 else []
    else:
        keys_out = list(args.pop(-1)) if len(args) % 2 == 1 else []
        nodes = args[::2]
        keysets = [list(k) for k in args[1::2]]
    return list(nodes), keysets, keys_out


class SumMultiply(Deterministic):
    """``SumMultiply('ik,k->i', A, x)`` or ``SumMultiply(A, [0,1], x, [1], [0])`` —
    the call syntax of the reference node (dot.py:115-301)."""
    moment_kind = "gaussian"

    def __init__(self, *args, iterator_axis=None, plates=None, name=""):
        if iterator_axis is not None:
            raise NotImplementedError("Iterator axis not implemented yet")
        nodes, keysets, keys_out = _parse(args)
        # inputs and output are Gaussian unless at least one parent is Gaussian-gamma; then fixed arrays stay plain
        # values (no scale of their own) and Gaussian parents count with tau = 1 (dot.py:177-222)
        self.gaussian_gamma = any(isinstance(n, Node) and n.moment_kind == "gaussian_gamma" for n in nodes)
        self.is_constant = [not isinstance(n, Node) or isinstance(n, Constant) for n in nodes]
        if self.gaussian_gamma:
            from .gaussian_gamma import ensure_gaussian_gamma
            self.moment_kind = "gaussian_gamma"
            nodes = [ensure_gaussian(n, len(k)) if c else ensure_gaussian_gamma(n, len(k))
                     for n, k, c in zip(nodes, keysets, self.is_constant)]
        else:
            nodes = [ensure_gaussian(n, len(k)) for n, k in zip(nodes, keysets)]
        full = []
        for ks in keysets:
            for k in ks:
                if k not in full:
                    full.append(k)
        for n, (node, ks) in enumerate(zip(nodes, keysets)):
            if len(node.dims[0]) != len(ks):
                raise ValueError("Wrong number of keys (%d) for the node number %d with %d dimensions"
                                 % (len(ks), n, len(node.dims[0])))
            if len(set(ks)) != len(ks):
                raise ValueError("Axis keys for node number %d are not unique" % n)
        if len(keys_out) != len(set(keys_out)):
            raise ValueError("Output keys are not unique")
        for k in keys_out:
            if k not in full:
                raise ValueError("Output key %s does not appear in any input" % k)
        size = {}
        for k in full:
            size[k] = 1
            for node, ks in zip(nodes, keysets):
                if k in ks:
                    n = node.dims[0][ks.index(k)]
                    if n != size[k]:
                        if size[k] == 1:
                            size[k] = n
                        elif n != 1:
                            raise ValueError("Axes using key %s do not broadcast properly" % k)
        self.key_size = size
        self.in_keys = keysets
        self.out_keys = keys_out
        shape = tuple(size[k] for k in keys_out)
        dims = (shape, shape + shape, (), ()) if self.gaussian_gamma else (shape, shape + shape)
        super().__init__(*nodes, dims=dims, plates=plates, name=name)

    # -- label helpers: plate axis j (counted from the right, 1-based) -> ('p', j)
    @staticmethod
    def _plate_keys(n):
        return [("p", j) for j in range(n, 0, -1)]

    def _keys(self, dim_keys, ind):
        d = [("d", k) for k in dim_keys]
        if ind == 0:
            return d
        return [("e", k) for k in dim_keys] + d

    def _operand(self, u, dim_keys, ind):
        a = dense(u[ind])
        ndims = (ind + 1) * len(dim_keys)
        return a, self._plate_keys(a.ndim - ndims) + self._keys(dim_keys, ind)

    def _compute_moments(self, *u_parents):
        """<f> and <f f^T> (dot.py:316-415)."""
        out = []
        for ind in range(2):
            ops, ksets, npl = [], [], 0
            for u, ks in zip(u_parents, self.in_keys):
                a, k = self._operand(u, ks, ind)
                ops.append(a)
                ksets.append(k)
                npl = max(npl, a.ndim - (ind + 1) * len(ks))
            out_keys = self._plate_keys(npl) + self._keys(self.out_keys, ind)
            out.append(D.sum_product(ops, ksets, out_keys))
        if self.gaussian_gamma:
            # <tau> and <log tau> of the product: the scales multiply (dot.py:406-413)
            tau, logtau = D.asarray(1.0), D.asarray(0.0)
            for u, c in zip(u_parents, self.is_constant):
                if not c:
                    tau, logtau = D.mul(tau, u[2]), D.add(logtau, u[3])
            out += [tau, logtau]
        return out

    def message_to_parent(self, index):
        """Message to parent[index], already summed to its plates (dot.py:425-633)."""
        if index >= len(self.parents):
            raise ValueError("Parent index larger than the number of parents")
        parent = self.parents[index]
        u_parents = self.moments_from_parents(exclude=index)
        m = self.message_from_children()
        npl_self = len(self.plates)
        npl_par = len(parent.plates)
        msg = []
        for ind in range(2, 4 if self.gaussian_gamma and not self.is_constant[index] else 2):
            # the scale part: m2 times the other parents' <tau>, m3 as it is, summed to the parent's plates (dot.py:617-631)
            if m[ind] is None:
                msg.append(None)
                continue
            ops, ksets = [D.asarray(m[ind])], [self._plate_keys(D.asarray(m[ind]).ndim)]
            if ind == 2:
                for k, u in enumerate(u_parents):
                    if k != index and not self.is_constant[k]:
                        a = D.asarray(u[2])
                        ops.append(a)
                        ksets.append(self._plate_keys(a.ndim))
            pk = [("p", j) for j in range(npl_par, 0, -1) if parent.plates[npl_par - j] != 1]
            sizes = {("p", j): self.plates[npl_self - j] for j in range(1, npl_self + 1) if ("p", j) not in pk}
            r = D.sum_product(ops, ksets, pk, sizes=sizes)
            it = iter(tuple(r.shape))
            msg.append(r.reshape(tuple(next(it) if parent.plates[a] != 1 else 1 for a in range(npl_par))))
        scale_msgs, msg = msg, []
        for ind in range(2):
            if m[ind] is None:
                msg.append(None)
                continue
            ops, ksets = [], []
            for k, u in enumerate(u_parents):
                if k == index:
                    continue
                a, ks = self._operand(u, self.in_keys[k], ind)
                ops.append(a)
                ksets.append(ks)
            mc = m[ind]
            nd_c = (ind + 1) * len(self.out_keys)
            ops.append(mc)
            ksets.append(self._plate_keys(mc.ndim - nd_c) + self._keys(self.out_keys, ind))
            # keep a plate key only where the parent really has that plate
            pk = []
            for j in range(npl_par, 0, -1):
                if parent.plates[npl_par - j] != 1:
                    pk.append(("p", j))
            dk_all = self._keys(self.in_keys[index], ind)
            want = tuple(parent.dims[ind])
            # a variable axis the parent holds with length 1 under a longer key is summed like a plate
            bcast = [n == 1 and self.key_size[k[1]] != 1 for k, n in zip(dk_all, want)]
            dk = [k for k, b in zip(dk_all, bcast) if not b]
            # plates of this node that are summed: axes no operand spans still count (multiplier)
            sizes = {("p", j): self.plates[npl_self - j] for j in range(1, npl_self + 1) if ("p", j) not in pk}
            sizes.update({k: self.key_size[k[1]] for k, b in zip(dk_all, bcast) if b})
            r = D.sum_product(ops, ksets, pk + dk, sizes=sizes)
            # restore unit plate axes and force explicit variable dims
            pshape = tuple(r.shape[:len(pk)])
            it = iter(pshape)
            full_pl = tuple(next(it) if parent.plates[a] != 1 else 1 for a in range(npl_par))
            itd = iter(tuple(r.shape[len(pk):]))
            r = r.reshape(full_pl + tuple(1 if b else next(itd) for b in bcast))
            if tuple(r.shape[npl_par:]) != want:
                r = r.broadcast_to(full_pl + want).contiguous()
            msg.append(r)
        return msg + scale_msgs


def Dot(*args, **kwargs):
    """Inner product of Gaussian vectors (dot.py:636-644)."""
    return SumMultiply("i" + ",i" * (len(args) - 1), *args, **kwargs)
