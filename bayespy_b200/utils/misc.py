"""Drop-in replacements of the plate-sum helpers of ``bayespy/utils/misc.py``
(sum_multiply :851-933, sum_product :935-945, sum_multiply_to_plates :805-844),
executed by ``bpk_sum_multiply`` on the GPU.

Inputs may be NumPy arrays (uploaded; the result comes back as NumPy, like the
reference's caller-owned-array convention) or device arrays (result stays on
device).  Shape / axis errors raise ``ValueError`` as in the reference.
"""
import numpy as np

from .. import darray as D
from ..darray import DArray


def _wrap(args):
    on_device = any(isinstance(a, DArray) for a in args)
    return [D.asarray(a) for a in args], on_device


def _ret(x, on_device):
    return x if on_device else x.numpy()


def sum_multiply(*args, axis=None, sumaxis=True, keepdims=False):
    """sum(arg0 * arg1 * ..., axis) without forming the product (misc.py:851)."""
    if len(args) == 0:
        raise ValueError("You must give at least one input array")
    arrs, on_device = _wrap(args)
    max_dim = max(a.ndim for a in arrs)
    if sumaxis:
        if axis is None:
            keep = []
        else:
            ax = [axis] if np.isscalar(axis) else list(axis)
            keep = [i for i in range(max_dim) if i not in ax and (i - max_dim) not in ax]
    else:
        if axis is None:
            keep = list(range(max_dim))
        else:
            ax = [axis] if np.isscalar(axis) else list(axis)
            keep = sorted(i if i >= 0 else i + max_dim for i in ax)
            if len(set(keep)) != len(keep):
                # the reference's einsum rejects a repeated output axis (misc.py:906)
                raise ValueError("Axis %s given several times" % (axis,))
    if keep and (min(keep) < 0 or max(keep) >= max_dim):
        raise ValueError("Axis index out of bounds")
    ksets = [list(range(max_dim - a.ndim, max_dim)) for a in arrs]
    y = D.sum_product(arrs, ksets, keep)
    if keepdims:
        it = iter(y.shape)
        y = y.reshape(tuple(next(it) if k in keep else 1 for k in range(max_dim)))
    return _ret(y, on_device)


def sum_product(*args, axes_to_keep=None, axes_to_sum=None, keepdims=False):
    """misc.py:935-945."""
    if axes_to_keep is not None:
        return sum_multiply(*args, axis=axes_to_keep, sumaxis=False, keepdims=keepdims)
    return sum_multiply(*args, axis=axes_to_sum, sumaxis=True, keepdims=keepdims)


def broadcasting_multiplier(plates, *args):
    """Product of the extents of ``plates`` that are unit/missing in every arg (misc.py:761)."""
    for a in args:
        np.broadcast_shapes(tuple(plates), tuple(a))
        if len(a) > len(plates) or any(x != 1 and x != y for x, y in zip(reversed(a), reversed(plates))):
            raise ValueError("The shapes in args are not a sub-shape of plates")
    r = 1
    for j in range(1, len(plates) + 1):
        if all(j > len(a) or a[-j] == 1 for a in args):
            r *= plates[-j]
    return r


def sum_multiply_to_plates(*arrays, to_plates=(), from_plates=None, ndim=0):
    """Product of the arrays summed to ``to_plates`` (+ndim trailing variable axes)
    with the broadcasting multiplier for ``from_plates`` (misc.py:805-844)."""
    arrs, on_device = _wrap(arrays)
    shapes = [a.shape[:a.ndim - ndim] if ndim else a.shape for a in arrs]
    prod_plates = tuple(np.broadcast_shapes(*shapes))
    nd = max(a.ndim for a in arrs)
    npl = nd - ndim
    if from_plates is None:
        r = 1
    else:
        r = broadcasting_multiplier(tuple(from_plates), prod_plates, tuple(to_plates))
    tgt = (1,) * (npl - len(to_plates)) + tuple(to_plates) if npl >= len(to_plates) else tuple(to_plates)[-npl:] if npl else ()
    keys = list(range(nd))
    out_keys = [k for k in keys if k >= npl or (tgt[k] != 1 and prod_plates[k - (npl - len(prod_plates))] != 1)]
    ksets = [keys[nd - a.ndim:] for a in arrs]
    y = D.sum_product(arrs, ksets, out_keys, scale=float(r))
    dims_shape = tuple(y.shape[len(out_keys) - ndim:]) if ndim else ()
    full = []
    it = iter(y.shape)
    for k in keys:
        full.append(next(it) if k in out_keys else 1)
    y = y.reshape(tuple(full))
    want = len(to_plates) + ndim
    if y.ndim > want:
        y = y.squeeze_leading(want)
    elif y.ndim < want:
        y = y.add_leading(want - y.ndim)
    return _ret(y, on_device)


# ---- small host helpers model scripts use (bookkeeping on NumPy arrays, nothing of the device path) ----------------
def trues(shape):
    return np.ones(shape, dtype=bool)


def atleast_nd(X, d):
    X = np.asarray(X)
    return X.reshape((1,) * (d - X.ndim) + X.shape) if X.ndim < d else X


def rmse(y1, y2, axis=None):
    return np.sqrt(np.mean((np.asarray(y1) - np.asarray(y2)) ** 2, axis=axis))


def grid(x1, x2):
    """All pairs (x1_i, x2_j) as an (M*N, 2) array, x1 varying fastest (misc.py:588-591)."""
    X1, X2 = np.meshgrid(x1, x2)
    return np.column_stack((X1.ravel(), X2.ravel()))


def identity(*shape):
    n = int(np.prod(shape)) if shape else 1
    return np.identity(n).reshape(tuple(shape) + tuple(shape))
