"""Host-side node graph: plates, masks and message routing.

This is bookkeeping only (shapes, who-is-whose-parent, boolean masks); every
floating-point array is a :class:`bayespy_b200.darray.DArray` in HBM and every
arithmetic step is a libbpk kernel.  It plays the role of the reference's
``bayespy/inference/vmp/nodes/node.py`` (Node :223-857), ``deterministic.py``
and ``constant.py`` but is organised differently: there are no Moments /
converter classes — a node advertises a *moment kind* string and parents given
as plain arrays become :class:`Constant` nodes of the kind the child expects.

Moment kinds and their per-plate dims
    "gaussian"     [<x>, <x x^T>]          dims (S, S+S)
    "gamma"        [<a>, <log a>]          dims ((), ())
    "gamma_prior"  [a, lgamma(a)]          dims ((), ())        (gamma.py:33-59)
    "wishart"      [<L>, <log|L|>]         dims ((D,D), ())
    "wishart_prior" [n, lgamma_D(n/2)]     dims ((), ())        (wishart.py:23-42)
    "dirichlet"    [<log p>]               dims ((K,),)
    "dirichlet_prior" [alpha, ...]         dims ((K,), ())      (dirichlet.py:25-62)
    "categorical"  [<one-hot>]             dims ((K,),)
"""
import numpy as np

from .. import darray as D
from ..darray import DArray


# --------------------------------------------------------------------------------------------
# shape helpers (pure host integer work)
# --------------------------------------------------------------------------------------------
def broadcast_plates(*plates):
    try:
        return tuple(int(n) for n in np.broadcast_shapes(*[tuple(p) for p in plates]))
    except ValueError:
        raise ValueError("The plates of the parents do not broadcast: %s" % (plates,))


def is_subshape(sub, full):
    """True if ``sub`` broadcasts to ``full`` without enlarging it (right-aligned)."""
    sub, full = tuple(sub), tuple(full)
    if len(sub) > len(full):
        return False
    for a, b in zip(reversed(sub), reversed(full)):
        if a != 1 and a != b:
            return False
    return True


def mask_is_full(mask):
    return mask is True or bool(np.all(mask))


def combine_multipliers(*mults):
    """Plate multipliers (tuples aligned with the trailing plates, 1 = not scaled) of a node and its parents combined
    like plates are: right-aligned, a unit entry yields to the other (node.py:295-301 ``_total_plates``)."""
    out = []
    for m in mults:
        if m is None:
            continue
        m = tuple(m)
        n = max(len(out), len(m))
        a = [1] * (n - len(out)) + list(out)
        b = [1] * (n - len(m)) + list(m)
        out = []
        for x, y in zip(a, b):
            if x != 1 and y != 1 and x != y:
                raise ValueError("The plate multipliers are incompatible: %s and %s" % (tuple(a), tuple(b)))
            out.append(y if x == 1 else x)
    while out and out[0] == 1:
        out.pop(0)
    return tuple(out)


def multiplier_to_parent(own, parent):
    """Factor by which a message grows on its way to a parent: the product of this node's plate multipliers on the
    axes where the parent (as seen from this node) has none (node.py:604-633)."""
    own, parent = tuple(own), tuple(parent)
    r = 1.0
    for j in range(1, len(own) + 1):
        if own[-j] != 1 and (j > len(parent) or parent[-j] == 1):
            r *= own[-j]
    return r


class _NodeType(type):
    """Every node class takes the reference's ``plotter=`` keyword (node.py:241-246, :860-866); it is kept for
    ``plot()`` and never reaches the class's own constructor."""

    def __call__(cls, *args, plotter=None, **kwargs):
        obj = super().__call__(*args, **kwargs)
        obj._plotter = plotter
        return obj


class Node(metaclass=_NodeType):
    """Base class: plates, parents/children, masks, message reduction."""

    moment_kind = None
    _plotter = None

    def set_plotter(self, plotter):
        self._plotter = plotter

    def plot(self, fig=None, **kwargs):
        """Plot with the node's plotter, if one was given (node.py:850-866)."""
        if self._plotter is None:
            raise Exception("No plotter defined, can not plot")
        return self._plotter(self, fig=fig, **kwargs)

    def show(self):
        print(str(self))

    def has_plotter(self):
        return self._plotter is not None

    def lowerbound(self):
        """This node's term of the lower bound as a float (node.py:844-848)."""
        return float(self.lower_bound_contribution())
    _id_counter = 0

    def __init__(self, *parents, dims=None, plates=None, name="", notify_parents=True, plates_multiplier=None):
        self.parents = list(parents)
        self.dims = tuple(tuple(d) for d in dims)
        self.name = name
        self.children = []          # (child, index) in registration order
        parent_plates = [self._plates_from_parent(i) for i in range(len(self.parents))]
        if plates is None:
            self.plates = broadcast_plates(*parent_plates) if parent_plates else ()
        else:
            plates = tuple(int(n) for n in plates)
            for p in parent_plates:
                if not is_subshape(p, plates):
                    raise ValueError("The plates %s of the parents are not broadcastable to the given "
                                     "plates %s." % (p, plates))
            self.plates = plates
        # stochastic VI: this node stands for `multiplier` times as many plates as it holds (node.py:257, :295-301)
        self.plates_multiplier = combine_multipliers(
            plates_multiplier, *[self._plates_multiplier_from_parent(i) for i in range(len(self.parents))])
        # which plates take part in inference at all (OR of the children's masks)
        self.mask = np.array(False)
        self._mask_dev = None
        self._version = 0           # bumped whenever the moments change (message caches key on it)
        if notify_parents:
            for i, p in enumerate(self.parents):
                p._add_child(self, i)

    # what the reference calls ``node._moments``: a tag naming the moment kind (engine/moments.py)
    @property
    def _moments(self):
        if "_moments_tag" in self.__dict__:
            return self.__dict__["_moments_tag"]
        from . import moments
        return moments.of(self)

    @_moments.setter
    def _moments(self, value):
        self.__dict__["_moments_tag"] = value

    @staticmethod
    def _ensure_moments(node, moments_class, **kwargs):
        """node.py:330-372: ``node`` (a node or an array) as a node with the wanted kind of moments."""
        from . import moments
        return moments.ensure(node, moments_class, **kwargs)

    # ---- graph ------------------------------------------------------------------------------
    def _add_child(self, child, index):
        self.children.append((child, index))

    def _remove_child(self, child, index):
        self.children.remove((child, index))

    def delete(self):
        for i, p in enumerate(self.parents):
            p._remove_child(self, i)
        for c, _ in list(self.children):
            c.delete()

    def _ids(self):
        """IDs of the stochastic factors this node depends on (independence check); ``_get_id_list`` is the
        reference's name for it (node.py:443-444)."""
        if hasattr(self, "_get_id_list"):
            return self._get_id_list()
        raise NotImplementedError

    def _check_independent_parents(self):
        ids = []
        for p in self.parents:
            ids += list(p._ids())
        if len(ids) != len(set(ids)):
            raise ValueError("Parent nodes are not independent")

    # ---- plates -----------------------------------------------------------------------------
    def _plates_from_parent(self, index):
        """Plates of parent[index] as seen from this node."""
        return tuple(self.parents[index].plates)

    def _plates_to_parent(self, index):
        """This node's plates as seen from parent[index]."""
        return tuple(self.plates)

    def _plates_multiplier_from_parent(self, index):
        """Plate multiplier of parent[index] mapped to this node's plate axes (the same map as for the plates)."""
        m = tuple(getattr(self.parents[index], "plates_multiplier", ()))
        if not m or all(v == 1 for v in m):
            return ()
        pp = tuple(self.parents[index].plates)
        full = (1,) * (len(pp) - len(m)) + m
        # map the parent's plates once with the multiplier values in place of the extents
        try:
            mapped = self._map_parent_axes(index, full)
        except Exception:
            raise ValueError("The plate multiplier %s of parent %d cannot be mapped to the plates of this node" % (m, index))
        return tuple(mapped)

    def _map_parent_axes(self, index, values):
        """Carry per-axis values of parent[index]'s plates over to this node's plate axes; the default is the identity
        (sub-classes whose plates differ from their parents' override ``_plates_from_parent`` and this with it)."""
        return tuple(values)

    def get_shape(self, i):
        return tuple(self.plates) + tuple(self.dims[i])

    # the names the reference's own unit tests call (node.py:429-441, :570, :657; deterministic.py:62-82)
    def _message_to_child(self):
        return self.get_moments()

    def _message_to_parent(self, index, u_parent=None):
        return self.message_to_parent(index)

    def _message_from_children(self, *args, **kwargs):
        return self.message_from_children()

    def _message_from_parents(self, exclude=None):
        return self.moments_from_parents(exclude=exclude)

    def _compute_weights_to_parent(self, index, weights):
        return self._weights_to_parent(index, np.asarray(weights))

    def __getitem__(self, index):
        """Basic slicing of the plates (node.py:761-763)."""
        return Slice(self, index, name=self.name + ".__getitem__")

    # ---- masks (host booleans; nodes.node.py:446-526) ------------------------------------------
    def get_mask(self):
        return self.mask

    def _set_mask(self, mask):
        self.mask = mask
        self._mask_dev = None

    def _weights_to_parent(self, index, mask):
        """Map this node's plate mask to the plate space of parent[index]."""
        return mask

    def _mask_to_parent(self, index):
        mask = np.asarray(self._weights_to_parent(index, self.mask)) != 0
        tgt = tuple(self.parents[index].plates)
        nd = mask.ndim
        full = (1,) * (nd - len(tgt)) + tgt if nd >= len(tgt) else tgt[len(tgt) - nd:]
        axes = tuple(i for i in range(nd) if full[i] == 1 and mask.shape[i] != 1)
        if axes:
            mask = np.any(mask, axis=axes, keepdims=True)
        while mask.ndim > len(tgt):
            mask = mask[0]
        return mask

    def _update_mask(self):
        mask = np.array(False)
        for child, index in self.children:
            mask = np.logical_or(mask, child._mask_to_parent(index))
        self._set_mask(mask)
        if not is_subshape(np.shape(self.mask), self.plates):
            raise ValueError("The mask of the node %s has updated incorrectly. The plates in the mask %s "
                             "are not a subset of the plates of the node %s."
                             % (self.name, np.shape(self.mask), self.plates))
        for p in self.parents:
            p._update_mask()

    def mask_device(self, mask=None):
        """u8 device copy of a host mask, or None when every plate is active."""
        if mask is None:
            if mask_is_full(self.mask):
                return None
            if self._mask_dev is None:
                self._mask_dev = DArray.from_numpy(np.asarray(self.mask, dtype=bool), "u1")
            return self._mask_dev
        if mask_is_full(mask):
            return None
        return DArray.from_numpy(np.asarray(mask, dtype=bool), "u1")

    # ---- messages -----------------------------------------------------------------------------
    def get_moments(self):
        # a node written against the reference's base class provides `_message_to_child` instead
        if type(self)._message_to_child is not Node._message_to_child:
            return [D.asarray(ui) for ui in self._message_to_child()]
        raise NotImplementedError

    def _message_and_mask_to_parent(self, index):
        """Un-reduced message list (device) and the host mask that gates it."""
        raise NotImplementedError

    def message_to_parent(self, index):
        """Message to parent[index], masked and summed to the parent's plates
        (node.py:570-655).  Entries may be None (no contribution)."""
        if index >= len(self.parents):
            raise ValueError("Parent index larger than the number of parents")
        m, mask = self._message_and_mask_to_parent(index)
        parent = self.parents[index]
        plates_self = self._plates_to_parent(index)
        scale = multiplier_to_parent(self.plates_multiplier, self._plates_multiplier_from_parent(index)) \
            if self.plates_multiplier else 1.0
        mdev = self.mask_device() if mask is self.mask else self.mask_device(mask)
        out = []
        for i, mi in enumerate(m):
            if mi is None:
                out.append(None)
                continue
            nd = len(parent.dims[i])
            mi = D.asarray(mi)
            dims = tuple(parent.dims[i])
            from_shape = tuple(plates_self) + dims
            to_shape = tuple(parent.plates) + dims
            mk = mdev.add_trailing(nd) if mdev is not None else None
            out.append(D.reduce_to_shape(mi, to_shape, mask=mk, from_shape=from_shape, scale=scale))
        return out

    def message_from_children(self):
        """Sum of the children's messages, one entry per moment (None = no message)."""
        msg = [None] * len(self.dims)
        for child, index in self.children:
            if type(child)._message_to_parent is not Node._message_to_parent:
                m = [None if mi is None else D.asarray(mi) for mi in child._message_to_parent(index)]
            else:
                m = child.message_to_parent(index)
            for i in range(len(self.dims)):
                if m[i] is not None:
                    msg[i] = m[i] if msg[i] is None else D.add(msg[i], m[i])
        return msg

    def moments_from_parents(self, exclude=None):
        return [p.get_moments() if i != exclude else None for i, p in enumerate(self.parents)]

    def lower_bound_contribution(self, ignore_masked=True):
        return 0.0


# --------------------------------------------------------------------------------------------
class Constant(Node):
    """Fixed moments (nodes/constant.py).  ``u`` is a list of DArrays shaped plates+dims."""

    def __init__(self, kind, u=None, dims=None, plates=None, name="const", value=None):
        if hasattr(kind, "fixed_moments"):
            # the reference's public form ``Constant(moments, x, name=...)`` (nodes/constant.py:14-35)
            moments, x = kind, u
            self._moments_spec = moments
            kind, u, dims, plates, value = moments.fixed_moments(x)
        else:
            self._moments_spec = None
        self.moment_kind = kind
        self.u = list(u)
        self.value = value
        super().__init__(dims=dims, plates=plates, name=name)

    def set_value(self, x):
        """Replace the fixed value (nodes/constant.py:46-57); the plates must not change."""
        if self._moments_spec is None:
            raise NotImplementedError("set_value needs a Constant created from a moments specification")
        kind, u, dims, plates, value = self._moments_spec.fixed_moments(x)
        if tuple(plates) != tuple(self.plates) or tuple(dims) != tuple(self.dims):
            raise ValueError("Incorrect shape of the value: plates %s expected" % (tuple(self.plates),))
        self.u = list(u)
        self.value = value
        self._version += 1

    def _ids(self):
        return []

    def get_moments(self):
        return list(self.u)

    def _update_mask(self):
        pass


# --------------------------------------------------------------------------------------------
class Deterministic(Node):
    """Node whose moments are a function of the parents' moments
    (nodes/deterministic.py:16-153).  Sub-classes implement
    ``_compute_moments(*u_parents)`` and
    ``_compute_message_to_parent(index, m_children, *u_parents)``."""

    def __init__(self, *parents, dims, plates=None, name=""):
        super().__init__(*parents, dims=dims, plates=plates, name=name)
        self._cache = None
        self._check_independent_parents()

    def _ids(self):
        ids = []
        for p in self.parents:
            ids += list(p._ids())
        return ids

    def _parent_versions(self):
        return tuple(getattr(p, "_version", 0) if not isinstance(p, Deterministic) else p._parent_versions()
                     for p in self.parents)

    def get_moments(self):
        key = self._parent_versions()
        if self._cache is None or self._cache[0] != key:
            u = self._compute_moments(*[p.get_moments() for p in self.parents])
            self._cache = (key, u)
        return list(self._cache[1])

    def _message_and_mask_to_parent(self, index):
        u_parents = self.moments_from_parents(exclude=index)
        m_children = self.message_from_children()
        m = self._compute_message_to_parent(index, m_children, *u_parents)
        mask = self._weights_to_parent(index, self.mask)
        return m, mask

    def lower_bound_contribution(self, ignore_masked=True):
        return 0.0


# --------------------------------------------------------------------------------------------
def _np_index(slices):
    """Normalised slices (ranges) as NumPy index entries."""
    out = []
    for s in slices:
        if isinstance(s, range):
            out.append(slice(s.start, s.stop if s.stop >= 0 else None, s.step))
        else:
            out.append(s)
    return tuple(out)


class Slice(Deterministic):
    """Basic slicing of a node's plates: integers, slices, ``None`` (new unit plate) and one ``Ellipsis``
    (node.py:868-1160).  The moments are strided VIEWS of the parent's (no copy); the message to the parent is the
    child's message placed into a zero array of the parent's plates through the same view."""

    def __init__(self, X, slices, name=""):
        self.moment_kind = X.moment_kind
        slices = list(slices) if isinstance(slices, tuple) else [slices]
        num_axis, ellipsis_index = 0, None
        for k, s in enumerate(slices):
            if isinstance(s, (int, np.integer)) and not isinstance(s, bool) or isinstance(s, slice):
                num_axis += 1
            elif s is None:
                pass
            elif s is Ellipsis:
                if ellipsis_index is None:
                    ellipsis_index = k
                else:
                    num_axis += 1
                    slices[k] = slice(None)
            else:
                raise TypeError("Invalid argument type: {0}".format(s.__class__))
        if num_axis > len(X.plates):
            raise IndexError("Too many indices")
        expand = len(X.plates) - num_axis
        if ellipsis_index is not None:
            k = ellipsis_index
            slices = slices[:k] + [slice(None)] * expand + slices[k + 1:]
        else:
            slices = slices + [slice(None)] * expand
        j = 0
        for k, s in enumerate(slices):
            if s is None:
                continue
            n = X.plates[j]
            if isinstance(s, slice):
                s = range(*s.indices(n))          # normalised: start, stop, step (negative steps included)
                if len(s) <= 0:
                    raise IndexError("Slicing leads to empty plates")
            else:
                s = int(s)
                if s < 0:
                    s += n
                if s < 0 or s >= n:
                    raise IndexError("Index out of range")
            slices[k] = s
            j += 1
        self.slices = slices
        super().__init__(X, dims=X.dims, name=name)

    def _plates_to_parent(self, index):
        return tuple(self.parents[index].plates)

    def _plates_from_parent(self, index):
        plates, k = list(self.parents[index].plates), 0
        for s in self.slices:
            if isinstance(s, range):
                plates[k] = len(s)
                k += 1
            elif s is None:
                plates = plates[:k] + [1] + plates[k:]
                k += 1
            else:
                del plates[k]
        return tuple(plates)

    def _map_parent_axes(self, index, values):
        out, j = [], 0
        for s in self.slices:
            if s is None:
                out.append(1)
            else:
                if isinstance(s, range):
                    out.append(values[j])
                j += 1
        return tuple(out)

    def _compute_moments(self, u):
        pp = tuple(self.parents[0].plates)
        out = []
        for ui, dims in zip(u, self.dims):
            ui = D.asarray(ui.materialize() if hasattr(ui, "materialize") else ui)
            out.append(ui.broadcast_to(pp + tuple(dims)).basic_index(tuple(self.slices)))
        return out

    def _place(self, child_value, dims, dtype_mask=False):
        """Zero array over the parent's plates (+ dims) with ``child_value`` written through the slicing view."""
        pp = tuple(self.parents[0].plates)
        full = pp + tuple(dims)
        out = DArray.zeros(full)
        view = out.basic_index(tuple(self.slices))
        src = D.asarray(child_value)
        D.copy_into(view, src.broadcast_to(view.shape))
        return out

    def _weights_to_parent(self, index, mask):
        pp = tuple(self.parents[0].plates)
        w = np.zeros(pp, dtype=bool)
        idx = _np_index(self.slices)
        w[idx] = np.broadcast_to(np.asarray(mask, dtype=bool), w[idx].shape)
        return w

    def _compute_message_to_parent(self, index, m_children, u):
        if index != 0:
            raise ValueError("Invalid index")
        return [None if mi is None else self._place(mi, dims) for mi, dims in zip(m_children, self.dims)]
