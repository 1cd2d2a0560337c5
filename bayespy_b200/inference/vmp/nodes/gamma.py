"""``bayespy.inference.vmp.nodes.gamma`` names that scripts import."""
from ....engine.gamma import Gamma, GammaToDiagonalWishart                                          # noqa: F401
from ....engine.moments import GammaMoments, GammaPriorMoments                                      # noqa: F401


def diagonal(alpha):
    """A diagonal Wishart-like node made of gamma scalars (gamma.py:25-30)."""
    return GammaToDiagonalWishart(alpha, name=getattr(alpha, "name", "") + " as Wishart")
