"""bayespy_b200 — a B200-native (sm_100a) variational-message-passing engine
behind the BayesPy API (``bayespy.nodes`` / ``bayespy.inference.VB``).

Host code is Python (node graph bookkeeping); every moment / message / bound
computation runs in hand-written CUDA kernels of ``libbpk.so`` reached through
the C ABI in ``include/bpk.h``.  There is no CPU fallback: importing is cheap,
but the first array operation raises if the library or a B200 is missing.

    import bayespy_b200 as bayespy            # or bayespy_b200.install_as_bayespy()
    from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
"""
import sys as _sys

from . import _bpk            # noqa: F401
from . import darray          # noqa: F401
from . import nodes           # noqa: F401
from . import inference       # noqa: F401
from . import utils           # noqa: F401

__version__ = "0.1.0"


def install_as_bayespy():
    """Register this package under the name ``bayespy`` so that unmodified model
    scripts (``from bayespy.nodes import ...``) run on the B200 engine."""
    mods = {
        "bayespy": _sys.modules[__name__],
        "bayespy.nodes": nodes,
        "bayespy.inference": inference,
        "bayespy.utils": utils,
        "bayespy.utils.misc": utils.misc,
        "bayespy.utils.linalg": utils.linalg,
        "bayespy.utils.random": utils.random,
    }
    for k, v in mods.items():
        _sys.modules[k] = v
