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


class _Permissive:
    """Stand-in for a plotting module: every attribute exists, every call is a no-op, and it works as a decorator."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self

    def __call__(self, *args, **kwargs):
        if len(args) == 1 and not kwargs and callable(args[0]) and not isinstance(args[0], _Permissive):
            return args[0]
        return self

    def __iter__(self):
        return iter(())


def install_as_bayespy(stub_plotting=False):
    """Register this package under the name ``bayespy`` so that unmodified model scripts
    (``from bayespy.nodes import ...``, ``from bayespy.inference.vmp.vmp import VB``,
    ``from bayespy.inference.vmp import transformations``) run on the B200 engine.  ``stub_plotting=True`` also answers
    ``bayespy.plot`` (and ``matplotlib`` when it is not installed) with no-op stand-ins, for scripts that plot."""
    import types
    from .inference import vmp as _vmp, vb as _vb
    from .inference.vmp import transformations as _tr
    from .inference.vmp import nodes as _vmp_nodes
    from .inference.vmp.nodes import gaussian, gamma, constant, categorical      # noqa: F401
    vmp_vmp = types.ModuleType("bayespy.inference.vmp.vmp")
    vmp_vmp.VB = inference.VB
    mods = {
        "bayespy": _sys.modules[__name__],
        "bayespy.nodes": nodes,
        "bayespy.inference": inference,
        "bayespy.inference.vmp": _vmp,
        "bayespy.inference.vmp.vmp": vmp_vmp,
        "bayespy.inference.vmp.nodes": _vmp_nodes,
        "bayespy.inference.vmp.nodes.gaussian": _vmp_nodes.gaussian,
        "bayespy.inference.vmp.nodes.gamma": _vmp_nodes.gamma,
        "bayespy.inference.vmp.nodes.constant": _vmp_nodes.constant,
        "bayespy.inference.vmp.nodes.categorical": _vmp_nodes.categorical,
        "bayespy.inference.vmp.transformations": _tr,
        "bayespy.utils": utils,
        "bayespy.utils.misc": utils.misc,
        "bayespy.utils.linalg": utils.linalg,
        "bayespy.utils.random": utils.random,
    }
    _vmp.vmp, _vmp.transformations = vmp_vmp, _tr
    if stub_plotting:
        plot = _Permissive()
        mods["bayespy.plot"] = plot
        _sys.modules[__name__].plot = plot
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mods["matplotlib"] = _Permissive()
            mods["matplotlib.pyplot"] = _Permissive()
    for k, v in mods.items():
        _sys.modules[k] = v
