"""``bayespy.inference.vmp.nodes.gaussian`` names that scripts import."""
from ....engine.gaussian import Gaussian, GaussianARD                                              # noqa: F401
from ....engine.gaussian_gamma import GaussianGamma, GaussianToGaussianGamma                       # noqa: F401
from ....engine.moments import GaussianMoments, GaussianGammaMoments                               # noqa: F401
