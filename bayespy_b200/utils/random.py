"""Host RNG helpers kept on NumPy's legacy global generator so that seeded runs
reproduce the reference's streams (bayespy/utils/random.py:45-78 ``mask``)."""
import numpy as np


def mask(*shape, p=0.5):
    """Boolean mask, True with probability ``p`` (utils/random.py:45-78)."""
    return np.random.rand(*shape) < p


def categorical(p, size=None):
    """Class labels drawn with probabilities ``p[..., k]`` (utils/random.py:247-288): one uniform draw per label,
    inverted through the cumulative probabilities, in C order of the plates."""
    p = np.asarray(p, dtype=np.float64)
    if size is None:
        size = p.shape[:-1]
    size = tuple(size) if np.ndim(size) else (int(size),)
    if np.any(p < 0) or np.any(np.isnan(p)):
        raise ValueError("Array contains negative probabilities")
    cum = np.cumsum(p / np.sum(p, axis=-1, keepdims=True), axis=-1)
    cum = np.broadcast_to(cum, size + (p.shape[-1],))
    x = np.random.rand(*size)
    if size == ():
        return int(np.searchsorted(cum, x))
    z = np.zeros(size, dtype=np.int64)
    for ind in np.ndindex(*size):
        z[ind] = np.searchsorted(cum[ind], x[ind])
    return z


def covariance(D, size=(), nu=None):
    """Random covariance matrices from an inverse-Wishart: the inverse of W W^T / nu with W of shape size + (D, nu)
    standard normal (utils/random.py:80-113)."""
    if nu is None:
        nu = D
    if nu < D:
        raise ValueError("nu must be greater than or equal to D")
    try:
        size = tuple(size)
    except TypeError:
        size = (size,)
    W = np.random.randn(*(size + (D, nu)))
    return np.linalg.inv(W @ np.swapaxes(W, -1, -2) / nu)


def bernoulli(p, size=None):
    """Boolean draws with success probabilities ``p`` (utils/random.py:236-244)."""
    if isinstance(size, int):
        size = (size,)
    if size is None:
        size = np.shape(p)
    return np.random.rand(*size) < p
