"""Names for the moment kinds.

The engine identifies what a node hands to its children by a string (``node.moment_kind``, see engine/node.py); the
reference uses one Moments class per kind (nodes/node.py:60-220, gaussian.py:42-290, gamma.py:33-87, wishart.py:23-115,
dirichlet.py:25-104, categorical.py:20-60).  Code written against the reference asks ``isinstance(node._moments,
WishartMoments)``: these tag classes answer that question; they carry no arithmetic."""


class Moments:
    kind = None

    def __init__(self, dims=None):
        self.dims = dims

    def __repr__(self):
        return "%s(dims=%s)" % (type(self).__name__, self.dims)


class GaussianMoments(Moments):
    kind = "gaussian"


class GaussianGammaMoments(Moments):
    kind = "gaussian_gamma"


class GammaMoments(Moments):
    kind = "gamma"


class GammaPriorMoments(Moments):
    kind = "gamma_prior"


class WishartMoments(Moments):
    kind = "wishart"


class WishartPriorMoments(Moments):
    kind = "wishart_prior"


class DirichletMoments(Moments):
    kind = "dirichlet"


class DirichletPriorMoments(Moments):
    kind = "dirichlet_prior"


BY_KIND = {c.kind: c for c in (GaussianMoments, GaussianGammaMoments, GammaMoments, GammaPriorMoments, WishartMoments,
                               WishartPriorMoments, DirichletMoments, DirichletPriorMoments)}


def of(node):
    """Tag object for a node's moment kind (a plain ``Moments`` for kinds without a class of their own)."""
    return BY_KIND.get(node.moment_kind, Moments)(getattr(node, "dims", None))
