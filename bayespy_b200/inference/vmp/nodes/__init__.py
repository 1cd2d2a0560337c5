"""Import paths of the reference's node package (bayespy/inference/vmp/nodes/__init__.py): model scripts use
``from bayespy.inference.vmp import nodes`` / ``nodes.Dirichlet(...)`` as well as ``from bayespy.nodes import ...``."""
from ....nodes import *                                                         # noqa: F401,F403
from ....nodes import (Node, Constant, Deterministic, Slice, ExponentialFamily, GaussianARD, Gamma, SumMultiply, Dot,   # noqa: F401
                       Gaussian, GaussianGamma, Wishart, Dirichlet, Categorical, CategoricalMarkovChain, Multinomial,
                       Mixture, GaussianMarkovChain, VaryingGaussianMarkovChain, SwitchingGaussianMarkovChain, Take, Gate)
