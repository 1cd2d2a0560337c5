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

    def __init__(self, shape=None):
        # the reference's constructor takes the variable shape (gaussian.py:42-53); a node's tag carries its dims
        if shape is not None and len(shape) > 0 and isinstance(shape[0], (tuple, list)):
            dims, shape = tuple(shape), tuple(shape[0])
        else:
            shape = None if shape is None else tuple(shape)
            dims = None if shape is None else (shape, shape + shape)
        super().__init__(dims)
        self.shape = shape
        self.ndim = None if shape is None else len(shape)


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


def _categorical_moments():
    from .categorical import CategoricalMoments          # carries the one-hot encoding of labels (engine/categorical.py)
    return CategoricalMoments


def __getattr__(name):
    if name == "CategoricalMoments":
        return _categorical_moments()
    raise AttributeError(name)


def ensure(x, moments, **kwargs):
    """``Node._ensure_moments(x, MomentsClass, **kwargs)`` of the reference (node.py:330-372) for the kinds of this
    engine: a node of the wanted kind is returned as it is (through its converter when it has one, e.g. a Markov chain
    seen as Gaussian / categorical variables over time), an array becomes a constant node of that kind."""
    from .node import Node
    kind = getattr(moments, "kind", None) or str(moments)
    ndim = kwargs.get("ndim", getattr(moments, "ndim", None))
    if isinstance(x, Node) and x.moment_kind == kind:
        return x
    if kind == "gaussian":
        from .gaussian import ensure_gaussian
        return ensure_gaussian(x, 0 if ndim is None else ndim)
    if kind == "gaussian_gamma":
        from .gaussian_gamma import ensure_gaussian_gamma
        return ensure_gaussian_gamma(x, 1 if ndim is None else ndim)
    if kind == "gamma":
        from .gaussian import ensure_gamma
        return ensure_gamma(x)
    if kind == "wishart":
        from .wishart import ensure_wishart
        return ensure_wishart(x)
    if kind == "dirichlet":
        from .dirichlet import dirichlet_constant
        if isinstance(x, Node):
            raise ValueError("Expected a Dirichlet-like node, got %s" % type(x).__name__)
        return dirichlet_constant(x)
    if kind == "categorical":
        from .categorical import categorical_constant
        if isinstance(x, Node):
            if hasattr(x, "_to_categorical"):
                return x._to_categorical()
            raise ValueError("Expected a categorical-like node, got %s" % type(x).__name__)
        return categorical_constant(x, kwargs.get("categories", getattr(moments, "categories", None)))
    raise ValueError("No conversion to %s moments" % kind)


def of(node):
    """Tag object for a node's moment kind (a plain ``Moments`` for kinds without a class of their own)."""
    dims = getattr(node, "dims", None)
    if node.moment_kind == "categorical" and dims:
        return _categorical_moments()(dims[0][0])
    cls = BY_KIND.get(node.moment_kind, Moments)
    return cls(dims) if dims is not None else cls()
