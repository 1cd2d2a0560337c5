"""Host RNG helpers kept on NumPy's legacy global generator so that seeded runs
reproduce the reference's streams (bayespy/utils/random.py:45-78 ``mask``)."""
import numpy as np


def mask(*shape, p=0.5):
    """Boolean mask, True with probability ``p`` (utils/random.py:45-78)."""
    return np.random.rand(*shape) < p
