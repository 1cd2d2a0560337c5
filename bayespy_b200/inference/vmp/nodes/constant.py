"""``bayespy.inference.vmp.nodes.constant``."""
from ....engine.node import Constant                                                                # noqa: F401
