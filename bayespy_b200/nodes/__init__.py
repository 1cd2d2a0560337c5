"""Node classes with the reference's public names (bayespy/nodes/__init__.py:105)."""
from ..engine.node import Node, Constant, Deterministic, Slice               # noqa: F401
from ..engine.expfam import ExponentialFamily                                 # noqa: F401
from ..engine.gaussian import GaussianARD                                     # noqa: F401
from ..engine.gamma import Gamma, GammaShape                                              # noqa: F401
from ..engine.dot import SumMultiply, Dot                                     # noqa: F401
from ..engine.gaussian import Gaussian                                        # noqa: F401
from ..engine.gaussian_gamma import GaussianGamma                            # noqa: F401
from ..engine.wishart import Wishart                                          # noqa: F401
from ..engine.dirichlet import Dirichlet, Concentration, DirichletConcentration, BetaConcentration                                      # noqa: F401
from ..engine.categorical import Categorical                                  # noqa: F401
from ..engine.multinomial import Multinomial                                # noqa: F401
from ..engine.categorical_markov_chain import CategoricalMarkovChain      # noqa: F401
from ..engine.binomial import Beta, Bernoulli, Binomial, Complement                         # noqa: F401
from ..engine.poisson import Poisson, Exponential                              # noqa: F401
from ..engine.add import Add                                                   # noqa: F401
from ..engine.concatenate import Concatenate                                  # noqa: F401
from ..engine.concat_gaussian import ConcatGaussian                           # noqa: F401
from ..engine.mixture import Mixture, MultiMixture                                          # noqa: F401
from ..engine.gmc import (GaussianMarkovChain, VaryingGaussianMarkovChain,   # noqa: F401
                          SwitchingGaussianMarkovChain)
from ..engine.take import Take                                                # noqa: F401
from ..engine.gate import Gate, Choose                                                # noqa: F401
