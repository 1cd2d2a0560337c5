"""Device-resident fp64 arrays for the host-side node graph.

A :class:`DArray` is a (pointer, shape, element-strides) handle on HBM memory
owned by libbpk's stream-ordered pool.  All arithmetic is dispatched to the
sm_100a kernels through :mod:`bayespy_b200._bpk`; nothing here computes on the
host.  ``numpy()`` / ``__array__`` materialise a copy on demand (that is the
only D2H path), so ``node.u`` / ``node.phi`` can be inspected like the
reference's NumPy arrays.

Views (reshape of contiguous data, broadcast, diagonal, transposed last axes)
are free: the generic kernels take arbitrary element strides.
"""
import numpy as np

from . import _bpk

_DT = {"f8": (np.float64, _bpk.F64), "u1": (np.uint8, _bpk.U8), "i8": (np.int64, None)}


def _contig_strides(shape):
    st, acc = [], 1
    for n in reversed(shape):
        st.append(acc)
        acc *= int(n)
    return tuple(reversed(st))


class _Owner:
    """Frees the allocation when the last view dies (stream-ordered free)."""
    __slots__ = ("ptr", "nbytes", "be")

    def __init__(self, be, nbytes):
        self.be = be
        self.nbytes = int(nbytes)
        self.ptr = be.malloc(self.nbytes)

    def __del__(self):
        try:
            self.be.free(self.ptr)
        except Exception:
            pass


class DArray:
    __slots__ = ("owner", "ptr", "shape", "strides", "dtype")
    __array_priority__ = 1000

    def __init__(self, owner, ptr, shape, strides, dtype="f8"):
        self.owner = owner
        self.ptr = ptr
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self.dtype = dtype

    # ---- construction -----------------------------------------------------------------
    @staticmethod
    def empty(shape, dtype="f8"):
        shape = tuple(int(s) for s in shape)
        be = _bpk.get()
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        own = _Owner(be, n * np.dtype(_DT[dtype][0]).itemsize)
        return DArray(own, own.ptr, shape, _contig_strides(shape), dtype)

    @staticmethod
    def zeros(shape, dtype="f8"):
        a = DArray.empty(shape, dtype)
        _bpk.get().memset(a.ptr, 0, a.owner.nbytes)
        return a

    @staticmethod
    def full(shape, value):
        a = DArray.empty(shape)
        if value == 0.0:
            _bpk.get().memset(a.ptr, 0, a.owner.nbytes)
        else:
            z = DArray.zeros(())
            _ew("AFFINE", a.shape, a, [z], alpha=0.0, beta=float(value))
        return a

    @staticmethod
    def from_numpy(x, dtype=None):
        x = np.asarray(x)
        if dtype is None:
            dtype = "u1" if x.dtype == np.bool_ else ("i8" if np.issubdtype(x.dtype, np.integer) else "f8")
        h = np.ascontiguousarray(x, dtype=_DT[dtype][0]).reshape(x.shape)   # ascontiguousarray makes 0-d -> 1-d
        a = DArray.empty(h.shape, dtype)
        if h.size:
            _bpk.get().h2d(a.ptr, h)
        return a

    # ---- inspection --------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    def is_contiguous(self):
        exp = _contig_strides(self.shape)
        return all(n == 1 or s == e for n, s, e in zip(self.shape, self.strides, exp))

    def contiguous(self):
        if self.is_contiguous():
            return self
        if self.dtype != "f8":
            return DArray.from_numpy(self.numpy(), self.dtype)
        out = DArray.empty(self.shape)
        _ew("COPY", self.shape, out, [self])
        return out

    def copy(self):
        out = DArray.empty(self.shape, self.dtype)
        if self.dtype == "f8":
            _ew("COPY", self.shape, out, [self])
        else:
            c = self.contiguous()
            _bpk.get().d2d(out.ptr, c.ptr, out.owner.nbytes)
        return out

    def numpy(self):
        npdt = _DT[self.dtype][0]
        c = self if self.is_contiguous() else self.contiguous()
        h = np.empty(c.shape, dtype=npdt)
        if h.size:
            _bpk.get().d2h(h, c.ptr)
        if self.dtype == "u1":
            return h.astype(bool)
        return h

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def item(self):
        return float(self.numpy().reshape(-1)[0])

    def __float__(self):
        return self.item()

    def __repr__(self):
        return "DArray(shape=%s, dtype=%s)" % (self.shape, self.dtype)

    def __len__(self):
        return self.shape[0]

    # ---- free views ----------------------------------------------------------------------
    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = tuple(int(s) for s in shape)
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
            shape = tuple(self.size // max(known, 1) if s == -1 else s for s in shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ValueError("cannot reshape array of size %d into shape %s" % (self.size, shape))
        # a broadcast (stride-0) or permuted view must be materialised first,
        # except when only unit axes are added/removed
        if tuple(s for s in shape if s != 1) == tuple(s for s in self.shape if s != 1):
            core = [st for n, st in zip(self.shape, self.strides) if n != 1]
            it = iter(core)
            strides = [next(it) if n != 1 else 0 for n in shape]
            return DArray(self.owner, self.ptr, shape, strides, self.dtype)
        c = self.contiguous()
        return DArray(c.owner, c.ptr, shape, _contig_strides(shape), c.dtype)

    def expand_dims(self, axis):
        nd = self.ndim + 1
        if axis < 0:
            axis += nd
        shape = self.shape[:axis] + (1,) + self.shape[axis:]
        strides = self.strides[:axis] + (0,) + self.strides[axis:]
        return DArray(self.owner, self.ptr, shape, strides, self.dtype)

    def add_leading(self, n):
        return DArray(self.owner, self.ptr, (1,) * n + self.shape, (0,) * n + self.strides, self.dtype)

    def add_trailing(self, n):
        return DArray(self.owner, self.ptr, self.shape + (1,) * n, self.strides + (0,) * n, self.dtype)

    def squeeze_leading(self, ndim_target):
        """Drop leading unit axes until ndim == ndim_target."""
        a = self
        while a.ndim > ndim_target:
            if a.shape[0] != 1:
                raise ValueError("cannot squeeze non-unit leading axis of %s" % (a.shape,))
            a = DArray(a.owner, a.ptr, a.shape[1:], a.strides[1:], a.dtype)
        return a

    def broadcast_to(self, shape):
        shape = tuple(int(s) for s in shape)
        a = self.add_leading(len(shape) - self.ndim) if len(shape) > self.ndim else self
        strides = []
        for n, s, t in zip(a.shape, a.strides, shape):
            if n == t:
                strides.append(s if n != 1 else 0)
            elif n == 1:
                strides.append(0)
            else:
                raise ValueError("cannot broadcast %s to %s" % (self.shape, shape))
        return DArray(a.owner, a.ptr, shape, strides, a.dtype)

    def swap_last2(self):
        sh, st = list(self.shape), list(self.strides)
        sh[-1], sh[-2] = sh[-2], sh[-1]
        st[-1], st[-2] = st[-2], st[-1]
        return DArray(self.owner, self.ptr, sh, st, self.dtype)

    def diag_view(self, ndim=1):
        """View of the diagonal of the trailing (dims, dims) block: (..., d1..dn, d1..dn) -> (..., d1..dn)."""
        if ndim == 0:
            return self
        p = self.ndim - 2 * ndim
        sh = self.shape[:p] + self.shape[p:p + ndim]
        st = self.strides[:p] + tuple(a + b for a, b in zip(self.strides[p:p + ndim], self.strides[p + ndim:]))
        return DArray(self.owner, self.ptr, sh, st, self.dtype)

    def index0(self, i):
        """self[i] along the first axis (view)."""
        esz = np.dtype(_DT[self.dtype][0]).itemsize
        return DArray(self.owner, self.ptr + int(i) * self.strides[0] * esz, self.shape[1:], self.strides[1:],
                      self.dtype)

    def __getitem__(self, index):
        """NumPy-style indexing.  Basic indexing (integers, slices, None, Ellipsis) returns a device view; anything else
        (index arrays, boolean masks) is answered from a host copy, like ``numpy()[index]``."""
        idx = index if isinstance(index, tuple) else (index,)
        if all(isinstance(i, (int, np.integer, slice)) or i is None or i is Ellipsis for i in idx):
            n_real = sum(1 for i in idx if i is not None and i is not Ellipsis)
            if n_real > self.ndim:
                raise IndexError("too many indices for array")
            out, seen = [], False
            for i in idx:
                if i is Ellipsis:
                    if seen:
                        raise IndexError("an index can only have a single ellipsis ('...')")
                    seen = True
                    out += [slice(None)] * (self.ndim - n_real)
                else:
                    out.append(i)
            return self.basic_index(tuple(out))
        return self.numpy()[index]

    @property
    def T(self):
        return DArray(self.owner, self.ptr, self.shape[::-1], self.strides[::-1], self.dtype)

    def basic_index(self, index):
        """NumPy basic indexing (integers, slices with any non-zero step, None) as a view: pointer, shape and element
        strides only.  ``index`` is a tuple with one entry per axis in order; missing trailing axes are taken whole."""
        esz = np.dtype(_DT[self.dtype][0]).itemsize
        ptr, shape, strides = self.ptr, [], []
        ax = 0
        for s in index:
            if s is None:
                shape.append(1)
                strides.append(0)
                continue
            n, st = self.shape[ax], self.strides[ax]
            if isinstance(s, (slice, range)):
                # a range is an already normalised slice (its stop may be -1 for negative steps)
                start, stop, step = (s.start, s.stop, s.step) if isinstance(s, range) else s.indices(n)
                length = len(range(start, stop, step))
                ptr += start * st * esz
                shape.append(length)
                strides.append(st * step)
            else:
                i = int(s)
                if i < 0:
                    i += n
                if i < 0 or i >= n:
                    raise IndexError("Index out of range")
                ptr += i * st * esz
            ax += 1
        shape += list(self.shape[ax:])
        strides += list(self.strides[ax:])
        return DArray(self.owner, ptr, tuple(shape), tuple(strides), self.dtype)

    def slice_axis(self, axis, start, stop):
        esz = np.dtype(_DT[self.dtype][0]).itemsize
        if axis < 0:
            axis += self.ndim
        sh = list(self.shape)
        sh[axis] = stop - start
        return DArray(self.owner, self.ptr + int(start) * self.strides[axis] * esz, sh, self.strides, self.dtype)

    # ---- arithmetic (all on device) ---------------------------------------------------------
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(self, o)
    def __sub__(self, o): return sub(self, o)
    def __rsub__(self, o): return sub(o, self)
    def __mul__(self, o): return mul(self, o)
    def __rmul__(self, o): return mul(self, o)
    def __truediv__(self, o): return div(self, o)
    def __rtruediv__(self, o): return div(o, self)
    def __neg__(self): return affine(self, -1.0, 0.0)


# =============================================================================================
# functional API
# =============================================================================================
def is_scalar(x):
    return isinstance(x, (int, float, np.floating, np.integer)) or (isinstance(x, np.ndarray) and x.ndim == 0)


def asarray(x):
    """DArray (fp64) from a DArray / NumPy array / scalar."""
    if isinstance(x, DArray):
        return x
    if hasattr(x, "materialize"):        # lazily produced device arrays (engine.plans, engine.gaussian)
        return x.materialize()
    return DArray.from_numpy(np.asarray(x, dtype=np.float64), "f8")


def bshape(*shapes):
    return tuple(np.broadcast_shapes(*shapes))


def _bstrides(a, shape):
    """Element strides of ``a`` broadcast (right-aligned) to ``shape``."""
    off = len(shape) - a.ndim
    st = [0] * len(shape)
    for i, (n, s) in enumerate(zip(a.shape, a.strides)):
        if n != 1:
            if n != shape[off + i]:
                raise ValueError("shape %s does not broadcast to %s" % (a.shape, shape))
            st[off + i] = s
    return st


def _collapse(shape, stride_lists):
    """Merge adjacent axes that are jointly contiguous in every operand; drop unit axes."""
    shp, sts = [], [[] for _ in stride_lists]
    for d, n in enumerate(shape):
        if n == 1:
            continue
        if shp and all(st[-1] == sl[d] * n for st, sl in zip(sts, stride_lists)):
            shp[-1] *= n
            for st, sl in zip(sts, stride_lists):
                st[-1] = sl[d]
        else:
            shp.append(n)
            for st, sl in zip(sts, stride_lists):
                st.append(sl[d])
    return shp, sts


def _ew(opname, shape, out, ins, alpha=0.0, beta=0.0):
    be = _bpk.get()
    shape = tuple(shape)
    if any(n == 0 for n in shape):
        return out
    lists = [_bstrides(out, shape)] + [_bstrides(a, shape) for a in ins]
    shp, sts = _collapse(shape, lists)
    if len(shp) > _bpk.MAXD:
        raise ValueError("elementwise rank %d exceeds %d" % (len(shp), _bpk.MAXD))
    dts = [_DT[a.dtype][1] for a in ins]
    be.ewise(_bpk.OPS[opname], shp, out.ptr, sts[0], [a.ptr for a in ins], dts, sts[1:], alpha, beta)
    return out


def _binary(opname, a, b, alpha=0.0, beta=0.0):
    a, b = asarray(a), asarray(b)
    shape = bshape(a.shape, b.shape)
    return _ew(opname, shape, DArray.empty(shape), [a, b], alpha, beta)


def _unary(opname, a, alpha=0.0, beta=0.0):
    a = asarray(a)
    return _ew(opname, a.shape, DArray.empty(a.shape), [a], alpha, beta)


def affine(a, alpha, beta=0.0):
    return _unary("AFFINE", a, alpha, beta)


def add(a, b):
    if is_scalar(b):
        return affine(a, 1.0, float(b))
    if is_scalar(a):
        return affine(b, 1.0, float(a))
    return _binary("ADD", a, b)


def sub(a, b):
    if is_scalar(b):
        return affine(a, 1.0, -float(b))
    if is_scalar(a):
        return affine(b, -1.0, float(a))
    return _binary("SUB", a, b)


def mul(a, b):
    if is_scalar(b):
        return affine(a, float(b), 0.0)
    if is_scalar(a):
        return affine(b, float(a), 0.0)
    return _binary("MUL", a, b)


def div(a, b):
    if is_scalar(b):
        return affine(a, 1.0 / float(b), 0.0)
    if is_scalar(a):
        return _unary("RECIP", b, float(a))
    return _binary("DIV", a, b)


def axpby(alpha, a, beta, b):
    return _binary("AXPBY", a, b, alpha, beta)


def fma(alpha, a, b, beta, c):
    a, b, c = asarray(a), asarray(b), asarray(c)
    shape = bshape(a.shape, b.shape, c.shape)
    return _ew("FMA", shape, DArray.empty(shape), [a, b, c], alpha, beta)


def where(mask, a, b):
    """mask: u1/f8 DArray (non-zero = take a)."""
    a, b = asarray(a), asarray(b)
    shape = bshape(mask.shape, a.shape, b.shape)
    return _ew("WHERE", shape, DArray.empty(shape), [mask, a, b])


def nonzero_select(u, v):
    """u != 0 ? v : 0   (expfamily.py:463)."""
    return _binary("NONZERO", u, v)


def log(a): return _unary("LOG", a)
def exp(a): return _unary("EXP", a)
def square(a): return _unary("SQUARE", a)
def sqrt(a): return _unary("SQRT", a)
def gammaln(a): return _unary("LGAMMA", a)
def digamma(a): return _unary("DIGAMMA", a)
def trigamma(a): return _unary("TRIGAMMA", a)
def multigammaln(a, d): return _unary("MVLGAMMA", a, float(d))
def multidigamma(a, d): return _unary("MVDIGAMMA", a, float(d))


def copy_into(dst_view, src):
    """dst_view[...] = broadcast(src); dst may be a strided view (e.g. a diagonal)."""
    return _ew("COPY", dst_view.shape, dst_view, [asarray(src)])


def sum_product(arrays, keysets, out_keys, scale=1.0, out=None, accumulate=False, sizes=None):
    """Restricted einsum on device: out[out_keys] = scale * sum prod_i arrays[i][keysets[i]].

    Keys are hashable labels, unique within one operand.  Operand axes of length 1
    broadcast against the key's extent.  ``sizes`` may give extents of keys that no
    operand spans (broadcast-only keys; they multiply the sum when not kept and are
    emitted with extent ``sizes[k]`` when kept).
    """
    be = _bpk.get()
    arrays = [asarray(a) for a in arrays]
    if len(arrays) > _bpk.MAXIN:
        # fold the tail pairwise with elementwise products over the union index space
        head, hk = arrays[:_bpk.MAXIN - 1], list(keysets[:_bpk.MAXIN - 1])
        tail, tk = arrays[_bpk.MAXIN - 1:], list(keysets[_bpk.MAXIN - 1:])
        union = []
        for ks in tk:
            for k in ks:
                if k not in union:
                    union.append(k)
        prod = sum_product(tail[:2], tk[:2], union, sizes=sizes) if len(tail) >= 2 else tail[0]
        for a, ks in zip(tail[2:], tk[2:]):
            prod = sum_product([prod, a], [union, ks], union, sizes=sizes)
        return sum_product(head + [prod], hk + [union], out_keys, scale, out, accumulate, sizes)
    ext = dict(sizes or {})
    for a, ks in zip(arrays, keysets):
        if len(ks) != a.ndim:
            raise ValueError("operand of rank %d given %d keys" % (a.ndim, len(ks)))
        for k, n in zip(ks, a.shape):
            if ext.get(k, 1) == 1:
                ext[k] = n
            elif n != 1 and n != ext[k]:
                raise ValueError("key %r has inconsistent extents %d and %d" % (k, ext[k], n))
    out_keys = list(out_keys)
    for k in out_keys:
        ext.setdefault(k, 1)
    summed = [k for k in ext if k not in out_keys]
    order = out_keys + summed
    shape = [ext[k] for k in order]
    oshape = tuple(ext[k] for k in out_keys)
    if out is None:
        out = DArray.empty(oshape)
    elif tuple(out.shape) != oshape:
        raise ValueError("out has shape %s, expected %s" % (out.shape, oshape))
    ostr = list(out.strides) + [0] * len(summed)
    # kept axes of extent > 1 must have a non-zero out stride (contiguous out guarantees it)
    in_strides = []
    for a, ks in zip(arrays, keysets):
        pos = {k: i for i, k in enumerate(ks)}
        st = []
        for k in order:
            if k in pos and a.shape[pos[k]] != 1:
                st.append(a.strides[pos[k]])
            else:
                st.append(0)
        in_strides.append(st)
    if any(n == 0 for n in shape):
        if not accumulate:
            be.memset(out.ptr, 0, out.owner.nbytes)
        return out
    shp, sts = _collapse_groups(shape, [ostr] + in_strides, len(out_keys))
    if len(shp) > _bpk.MAXD:
        if len(arrays) < 3:
            raise ValueError("sum_product rank %d exceeds %d" % (len(shp), _bpk.MAXD))
        # too many distinct axes for one launch: contract the two operands that leave the smallest intermediate first
        # (keys needed later - by another operand, the output or a multiplicity in ``sizes`` - are kept)
        best = None
        for i in range(len(arrays)):
            for j in range(i + 1, len(arrays)):
                later = set(out_keys) | set(sizes or ())
                for m, ks in enumerate(keysets):
                    if m != i and m != j:
                        later |= set(ks)
                keep = [k for k in dict.fromkeys(list(keysets[i]) + list(keysets[j])) if k in later]
                cost = int(np.prod([ext[k] for k in keep], dtype=np.int64)) if keep else 1
                if best is None or cost < best[0]:
                    best = (cost, i, j, keep)
        _, i, j, keep = best
        t = sum_product([arrays[i], arrays[j]], [keysets[i], keysets[j]], keep)
        rest = [m for m in range(len(arrays)) if m != i and m != j]
        return sum_product([t] + [arrays[m] for m in rest], [keep] + [keysets[m] for m in rest], out_keys, scale, out,
                           accumulate, sizes)
    dts = [_DT[a.dtype][1] for a in arrays]
    be.sum_multiply(shp, [a.ptr for a in arrays], dts, sts[1:], out.ptr, sts[0], scale, accumulate)
    return out


def _collapse_groups(shape, stride_lists, n_kept):
    """Collapse kept axes and summed axes separately (never across the boundary)."""
    s1, t1 = _collapse(shape[:n_kept], [sl[:n_kept] for sl in stride_lists])
    s2, t2 = _collapse(shape[n_kept:], [sl[n_kept:] for sl in stride_lists])
    return s1 + s2, [a + b for a, b in zip(t1, t2)]


def reduce_to_shape(a, target_shape, mask=None, scale=1.0, from_shape=None):
    """Masked plate-sum of Node._message_to_parent (node.py:619-653).

    ``a`` is a message whose shape is broadcastable to ``from_shape`` (the sender's
    plates+dims; defaults to a's own shape).  It is multiplied by the optional 0/1
    ``mask`` (right-aligned, broadcastable to ``from_shape``) and summed over every
    axis that is missing or of unit length in ``target_shape``.  Axes that are
    summed but along which neither operand varies contribute their full extent as
    a factor (misc.broadcasting_multiplier, misc.py:761).  Kept axes along which
    nothing varies stay of unit length (the receiver broadcasts)."""
    a = asarray(a)
    if from_shape is None:
        from_shape = a.shape if mask is None else bshape(a.shape, mask.shape)
    from_shape = tuple(int(n) for n in from_shape)
    nd = len(from_shape)
    if a.ndim > nd or (mask is not None and mask.ndim > nd) or len(target_shape) > nd:
        raise ValueError("shapes %s / %s do not fit in from_shape %s" % (a.shape, target_shape, from_shape))
    a = a.add_leading(nd - a.ndim)
    tgt = (1,) * (nd - len(target_shape)) + tuple(int(n) for n in target_shape)
    keys = list(range(nd))
    ops, ksets = [a], [keys]
    if mask is not None:
        ops.append(mask.add_leading(nd - mask.ndim))
        ksets.append(keys)
    varies = [any(o.shape[k] != 1 for o in ops) for k in keys]
    out_keys = [k for k in keys if tgt[k] != 1 and varies[k]]
    sizes = {k: from_shape[k] for k in keys if k not in out_keys and tgt[k] == 1}
    r = sum_product(ops, ksets, out_keys, scale=scale, sizes=sizes)
    full = tuple(from_shape[k] if k in out_keys else 1 for k in keys)
    r = r.reshape(full)
    return r.squeeze_leading(len(target_shape))
