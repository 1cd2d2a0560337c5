"""Drop-in replacements of ``bayespy/utils/linalg.py`` chol :31-63,
chol_solve :66-171, chol_inv :174-207, chol_logdet :209-223 (+ inner, outer,
mvdot) on the GPU.  The reference loops over plates in Python and calls LAPACK
per matrix; here one kernel launch factors / solves the whole batch.

NumPy in -> NumPy out; device arrays in -> device arrays out.
A non-SPD input raises ``Exception("Matrix not positive definite")`` as the
reference intends (linalg.py:58-59).
"""
import numpy as np

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .._bpk import NotPositiveDefinite   # noqa: F401


def _dev(x):
    return isinstance(x, DArray)


def _flat(shape):
    return int(np.prod(shape, dtype=np.int64)) if len(shape) else 1


def _mat(C, ndim):
    """View (..., d1..dn, d1..dn) as (batch, Dn, Dn)."""
    C = D.asarray(C)
    if C.ndim < 2 * ndim:
        raise ValueError("Array does not have enough axes for ndim=%d" % ndim)
    sh1 = C.shape[C.ndim - 2 * ndim:C.ndim - ndim]
    sh2 = C.shape[C.ndim - ndim:]
    if sh1 != sh2:
        raise ValueError("Not square matrix w.r.t. ndim sense")
    Dn = _flat(sh1)
    plates = C.shape[:C.ndim - 2 * ndim]
    return C.reshape(plates + (Dn, Dn)).contiguous(), plates, Dn, sh1


def chol(C, ndim=1):
    """Upper Cholesky factor U (C = U^T U) of a stack of SPD matrices."""
    on_dev = _dev(C)
    if ndim == 0:
        r = D.sqrt(D.asarray(C))
        return r if on_dev else r.numpy()
    Cm, plates, Dn, sh = _mat(C, ndim)
    U = DArray.empty(Cm.shape)
    _bpk.get().chol(Cm.ptr, U.ptr, _flat(plates), Dn, True)
    U = U.reshape(plates + sh + sh)
    return U if on_dev else U.numpy()


def chol_solve(U, b, out=None, matrix=False, ndim=1):
    """Solve (U^T U) x = b for stacked factors; ``b`` holds vectors (or matrices when
    ``matrix=True``); plates of U and b broadcast against each other."""
    on_dev = _dev(U) or _dev(b)
    if ndim == 0:
        Ud = D.asarray(U)
        r = D.div(D.asarray(b), D.square(Ud))
        return r if on_dev else r.numpy()
    Um, pU, Dn, sh = _mat(U, ndim)
    bd = D.asarray(b)
    if matrix:
        if bd.ndim < 2 * ndim:
            raise ValueError("b has too few axes")
        pb = bd.shape[:bd.ndim - 2 * ndim]
        nrhs = _flat(bd.shape[bd.ndim - ndim:])
        if _flat(bd.shape[bd.ndim - 2 * ndim:bd.ndim - ndim]) != Dn:
            raise ValueError("Shapes of U and b do not match")
        bm = bd.reshape(pb + (Dn, nrhs))
        tail = bd.shape[bd.ndim - 2 * ndim:]
    else:
        pb = bd.shape[:bd.ndim - ndim]
        if _flat(bd.shape[bd.ndim - ndim:]) != Dn:
            raise ValueError("Shapes of U and b do not match")
        nrhs = 1
        bm = bd.reshape(pb + (Dn, 1))
        tail = bd.shape[bd.ndim - ndim:]
    P = tuple(np.broadcast_shapes(pU, pb))
    N = _flat(P)

    def batch(a, pa, t):
        if _flat(pa) == 1:
            return a.contiguous(), 1
        if (1,) * (len(P) - len(pa)) + tuple(pa) != P:
            a = a.broadcast_to(P + t)
        return a.contiguous(), N
    Uc, nU = batch(Um, pU, (Dn, Dn))
    bc, nb = batch(bm, pb, (Dn, nrhs))
    X = DArray.empty(P + (Dn, nrhs))
    _bpk.get().chol_solve(Uc.ptr, nU, bc.ptr, nb, X.ptr, N, Dn, nrhs)
    X = X.reshape(P + tuple(tail))
    if out is not None:
        out[...] = X.numpy()
        return out
    return X if on_dev else X.numpy()


def chol_inv(U, ndim=1):
    on_dev = _dev(U)
    if ndim == 0:
        r = D._unary("RECIP", D.square(D.asarray(U)), 1.0)
        return r if on_dev else r.numpy()
    Um, plates, Dn, sh = _mat(U, ndim)
    out = DArray.empty(Um.shape)
    _bpk.get().chol_inv(Um.ptr, out.ptr, _flat(plates), Dn)
    out = out.reshape(plates + sh + sh)
    return out if on_dev else out.numpy()


def chol_logdet(U, ndim=1):
    on_dev = _dev(U)
    if ndim == 0:
        r = D.mul(D.log(D.asarray(U)), 2.0)
        return r if on_dev else r.numpy()
    Um, plates, Dn, sh = _mat(U, ndim)
    out = DArray.empty(plates)
    _bpk.get().chol_logdet(Um.ptr, out.ptr, _flat(plates), Dn)
    return out if on_dev else out.numpy()


def inner(*args, ndim=1):
    """Sum of the elementwise product over the last ``ndim`` axes (linalg.py:299-306)."""
    on_dev = any(_dev(a) for a in args)
    arrs = [D.asarray(a) for a in args]
    nd = max(a.ndim for a in arrs)
    keys = list(range(nd))
    r = D.sum_product(arrs, [keys[nd - a.ndim:] for a in arrs], keys[:nd - ndim])
    return r if on_dev else r.numpy()


def outer(A, B, ndim=1):
    """A[..., i] * B[..., j] over the last ``ndim`` axes (linalg.py:309-334)."""
    on_dev = _dev(A) or _dev(B)
    A, B = D.asarray(A), D.asarray(B)
    if ndim == 0:
        r = D.mul(A, B)
    else:
        a = A.add_trailing(ndim)
        b = B.reshape(B.shape[:B.ndim - ndim] + (1,) * ndim + B.shape[B.ndim - ndim:])
        r = D.mul(a, b)
    return r if on_dev else r.numpy()


def mvdot(A, b, ndim=1):
    """Matrix-vector product over the last 2*ndim / ndim axes (linalg.py:407-423)."""
    on_dev = _dev(A) or _dev(b)
    A, b = D.asarray(A), D.asarray(b)
    pa = A.ndim - 2 * ndim
    pb = b.ndim - ndim
    npl = max(pa, pb)
    pk = [("p", j) for j in range(npl, 0, -1)]
    rows = [("r", i) for i in range(ndim)]
    cols = [("c", i) for i in range(ndim)]
    r = D.sum_product([A, b], [pk[npl - pa:] + rows + cols, pk[npl - pb:] + cols], pk + rows)
    return r if on_dev else r.numpy()


def block_banded_solve(A, B, y):
    """Symmetric block-tridiagonal SPD system (linalg.py:468-575): ``A`` (..., N, D, D) diagonal blocks,
    ``B`` (..., N-1, D, D) super-diagonal blocks, ``y`` (..., N, D).  Returns the diagonal blocks ``V`` and the
    super-diagonal blocks ``C`` of the inverse, the solution ``x`` and the log-determinant — one
    ``bpk_block_banded_solve`` launch sequence (parallel in time) instead of the reference's loop over N.
    Leading plate axes of the three arguments broadcast like in the reference."""
    on_dev = _dev(A) or _dev(B) or _dev(y)
    A, B, y = D.asarray(A), D.asarray(B), D.asarray(y)
    if y.ndim < 2 or A.ndim < 3 or B.ndim < 3:
        raise ValueError("block_banded_solve: A (...,N,D,D), B (...,N-1,D,D), y (...,N,D) expected")
    N, Dn = y.shape[-2], y.shape[-1]
    if A.shape[-3] != N:
        raise ValueError("The number of diagonal blocks is incorrect")
    if tuple(A.shape[-2:]) != (Dn, Dn):
        raise ValueError("The diagonal blocks have wrong shape")
    if B.shape[-3] != N - 1:
        raise ValueError("The number of super-diagonal blocks is incorrect")
    if N > 1 and tuple(B.shape[-2:]) != (Dn, Dn):
        raise ValueError("The diagonal blocks have wrong shape")
    plates = tuple(np.broadcast_shapes(tuple(A.shape[:-3]), tuple(B.shape[:-3]), tuple(y.shape[:-2])))
    batch = _flat(plates)
    Ab = A.broadcast_to(plates + (N, Dn, Dn)).contiguous()
    Bb = B.broadcast_to(plates + (N - 1, Dn, Dn)).contiguous() if N > 1 else DArray.empty((1,))
    yb = y.broadcast_to(plates + (N, Dn)).contiguous()
    V = DArray.empty(plates + (N, Dn, Dn))
    C = DArray.empty(plates + (max(N - 1, 1), Dn, Dn))
    x = DArray.empty(plates + (N, Dn))
    ld = DArray.empty(plates)
    _bpk.get().block_banded_solve(Ab.ptr, Bb.ptr, yb.ptr, batch, N, Dn, V.ptr, C.ptr, x.ptr, ld.ptr, True)
    C = C.slice_axis(len(plates), 0, N - 1) if N > 1 else DArray.empty(plates + (0, Dn, Dn))
    if on_dev:
        return V, C, x, ld
    return V.numpy(), C.numpy(), x.numpy(), ld.numpy()
