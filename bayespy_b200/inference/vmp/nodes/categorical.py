"""``from bayespy.inference.vmp.nodes.categorical import CategoricalMoments`` (lda.rst:100)."""
from ....engine.categorical import CategoricalMoments, Categorical            # noqa: F401
