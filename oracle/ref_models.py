"""TEST INFRASTRUCTURE / CPU BASELINE — the benchmark's model scripts run through the UNMODIFIED
reference package staged in ``oracle/_ref`` (oracle/make_ref.py).

Each builder follows the reference's documented example line by line:
    pca()  doc/source/examples/pca.rst:40-66 (no rotation callback), optional ``mask``
    gmm()  doc/source/examples/gmm.rst:71-98
and returns the reference ``VB`` object plus its nodes, so that callers (bench.py's reference arm
and cpu_baseline leg, tests/test_reference_arm.py) time / inspect ``Q.update()`` of the reference
itself — ``kind: "reference"`` in bench.py's JSON.
"""
import time

import numpy as np

from . import make_ref


def pca(y, K, C_init, mask=None):
    """pca.rst:40-66 with M, N taken from ``y``; C initialised from ``C_init`` (M,1,K)."""
    make_ref.import_reference()
    from bayespy.nodes import GaussianARD, Gamma, SumMultiply
    from bayespy.inference import VB
    M, N = y.shape
    X = GaussianARD(0, 1, plates=(1, N), shape=(K,), name="X")
    alpha = Gamma(1e-5, 1e-5, plates=(K,), name="alpha")
    C = GaussianARD(0, alpha, plates=(M, 1), shape=(K,), name="C")
    F = SumMultiply("d,d->", X, C, name="F")
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    if mask is None:
        Y.observe(y)
    else:
        Y.observe(y, mask=mask)
    C.initialize_from_value(np.asarray(C_init).reshape((M, 1, K)))
    Q = VB(Y, X, C, alpha, tau)
    return Q, dict(X=X, C=C, alpha=alpha, tau=tau, Y=Y, F=F)


def gmm(y, K, z_init):
    """gmm.rst:71-98 with priors 1e-5; Z initialised from integer labels ``z_init`` (N,)."""
    make_ref.import_reference()
    from bayespy.nodes import Dirichlet, Categorical, Gaussian, Wishart, Mixture
    from bayespy.inference import VB
    N, D = y.shape
    alpha = Dirichlet(1e-5 * np.ones(K), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="Z")
    mu = Gaussian(np.zeros(D), 1e-5 * np.identity(D), plates=(K,), name="mu")
    Lam = Wishart(D, 1e-5 * np.identity(D), plates=(K,), name="Lambda")
    Y = Mixture(Z, Gaussian, mu, Lam, name="Y")
    Z.initialize_from_value(np.asarray(z_init))
    Y.observe(y)
    Q = VB(Y, mu, Lam, Z, alpha)
    return Q, dict(Y=Y, mu=mu, Lam=Lam, Z=Z, alpha=alpha)


def time_sweeps(Q, steps, warmup):
    """Seconds per ``Q.update()`` sweep (bound included, vmp.py:154-172), after ``warmup`` sweeps."""
    Q.ignore_bound_checks = True
    if warmup:
        Q.update(repeat=warmup, verbose=False)
    t = time.perf_counter()
    Q.update(repeat=steps, verbose=False)
    return (time.perf_counter() - t) / max(steps, 1)
