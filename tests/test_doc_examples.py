"""The reference's PUBLISHED doctest outputs reproduced by running the documented model
scripts through this package: doc/source/examples/lssm.rst (first 10 iterations, 80 % missing values) and
doc/source/examples/pca.rst (with the rotation callback).  The scripts below are the doc examples verbatim apart
from the import lines; seeds are the doc build's (numpy.random.seed(1), set in each file's hidden testsetup)."""
import re

import numpy as np


def _loglikes(out):
    return [float(m) for m in re.findall(r"loglike=([-+0-9.e]+)", out)]


def test_lssm_doc_example_first_ten_iterations(backend, capsys):
    """lssm.rst:8-11,45-181,199-203: 'Iteration 1: loglike=-1.439704e+05 ... Iteration 10: loglike=-1.051441e+04'."""
    from bayespy_b200.nodes import GaussianARD, GaussianMarkovChain, Gamma, Dot
    from bayespy_b200.inference import VB
    from bayespy_b200.utils import random
    np.random.seed(1)
    M, N, D = 30, 400, 10
    alpha = Gamma(1e-5, 1e-5, plates=(D,), name="alpha")
    A = GaussianARD(0, alpha, shape=(D,), plates=(D,), name="A")
    X = GaussianMarkovChain(np.zeros(D), 1e-3 * np.identity(D), A, np.ones(D), n=N, name="X")
    gamma = Gamma(1e-5, 1e-5, plates=(D,), name="gamma")
    C = GaussianARD(0, gamma, shape=(D,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    assert tuple(F.plates) == (30, 400)
    C.initialize_from_random()
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Q = VB(X, C, gamma, A, alpha, tau, Y)
    w = 0.3
    a = np.array([[np.cos(w), -np.sin(w), 0, 0], [np.sin(w), np.cos(w), 0, 0], [0, 0, 1, 0], [0, 0, 0, 0]])
    c = np.random.randn(M, 4)
    x = np.empty((N, 4)); f = np.empty((M, N)); y = np.empty((M, N))
    x[0] = 10 * np.random.randn(4)
    f[:, 0] = np.dot(c, x[0])
    y[:, 0] = f[:, 0] + 3 * np.random.randn(M)
    for n in range(N - 1):
        x[n + 1] = np.dot(a, x[n]) + [1, 1, 10, 10] * np.random.randn(4)
        f[:, n + 1] = np.dot(c, x[n + 1])
        y[:, n + 1] = f[:, n + 1] + 3 * np.random.randn(M)
    mask = random.mask(M, N, p=0.2)
    Y.observe(y, mask=mask)
    Q.update(repeat=10)
    L = _loglikes(capsys.readouterr().out)
    assert len(L) == 10
    # the doc prints 7 significant digits
    np.testing.assert_allclose(L[0], -1.439704e+05, rtol=5e-7)
    np.testing.assert_allclose(L[9], -1.051441e+04, rtol=5e-7)


def test_pca_doc_example_with_rotations(backend, capsys):
    """pca.rst:8-11,26-112,114-118: 'Iteration 1: loglike=-2.33...e+03 ... loglike=6.500...e+02, Converged'."""
    from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
    from bayespy_b200.inference.vmp.transformations import RotateGaussianARD, RotationOptimizer
    np.random.seed(1)
    M, N = 20, 100
    x = np.random.randn(N, 2)
    w = np.random.randn(M, 2)
    f = np.einsum("ik,jk->ij", w, x)
    y = f + 0.1 * np.random.randn(M, N)
    D = 10
    X = GaussianARD(0, 1, plates=(1, N), shape=(D,))
    alpha = Gamma(1e-5, 1e-5, plates=(D,))
    C = GaussianARD(0, alpha, plates=(M, 1), shape=(D,))
    F = SumMultiply("d,d->", X, C)
    tau = Gamma(1e-5, 1e-5)
    Y = GaussianARD(F, tau)
    Y.observe(y)
    Q = VB(Y, X, C, alpha, tau)
    C.initialize_from_random()
    rot_X = RotateGaussianARD(X)
    rot_C = RotateGaussianARD(C, alpha)
    R = RotationOptimizer(rot_X, rot_C, D)
    Q.set_callback(R.rotate)
    Q.update(repeat=1000)
    out = capsys.readouterr().out
    L = _loglikes(out)
    np.testing.assert_allclose(L[0], -2.33e+03, rtol=2.5e-3)          # doc: -2.33...e+03
    # doc: 6.500...e+02.  The last digits depend on WHEN the relative increment first dips under tol = 1e-5: the
    # L-BFGS rotation amplifies 1e-12 rounding differences between execution tiers after ~8 iterations (the first
    # 7 agree to 1e-10, tests/test_rotation.py pins them to the reference), so the run may stop a few iterations
    # earlier or later: per-node tier 650.0916 at iteration 31 (the doc's digits), fused-sweep tier 649.9677 at 24.
    np.testing.assert_allclose(L[-1], 650.09, rtol=5e-4)
    assert "Converged at iteration" in out


def test_user_guide_inference_example(backend, capsys):
    """doc/source/user_guide/inference.rst:5-28,40-46,60-63,96,154,185-233: the PCA model with four rows of the data
    missing, a random initialisation of X, then Q.update() / Q.update(C, X) / Q.update(C, X, C, tau) /
    Q.update(repeat=10) / Q.update(repeat=1000) / Q.update(repeat=10000, tol=1e-6) — every bound the guide prints
    ('Iteration 1: loglike=-9.305259e+02' ... 'Iteration 13: loglike=-1.405139e+02') and both convergence points
    ('Converged at iteration 488.', '... 847.')."""
    from bayespy_b200.nodes import GaussianARD, Gamma, Dot
    from bayespy_b200.inference import VB
    np.random.seed(1)
    D = 3
    X = GaussianARD(0, 1, shape=(D,), plates=(1, 100), name="X")
    alpha = Gamma(1e-3, 1e-3, plates=(D,), name="alpha")
    C = GaussianARD(0, alpha, shape=(D,), plates=(10, 1), name="C")
    F = Dot(C, X)
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau)
    c = np.random.randn(10, 2)
    x = np.random.randn(2, 100)
    data = np.dot(c, x) + 0.1 * np.random.randn(10, 100)
    Y.observe(data)
    Y.observe(data, mask=[[True], [False], [False], [True], [True], [False], [True], [True], [True], [False]])
    Q = VB(Y, C, X, alpha, tau)
    assert Q["X"] is X
    X.initialize_from_parameters(np.random.randn(1, 100, D), 10)
    Q.update()
    Q.update(C, X)
    Q.update(C, X, C, tau)
    Q.update(repeat=10)
    out = capsys.readouterr().out
    printed = ["-9.305259e+02", "-8.818976e+02", "-8.071222e+02", "-7.167588e+02", "-6.827873e+02", "-6.259477e+02",
               "-4.725400e+02", "-3.270816e+02", "-2.208865e+02", "-1.658761e+02", "-1.469468e+02", "-1.420311e+02",
               "-1.405139e+02"]
    for i, v in enumerate(printed):
        assert "Iteration %d: loglike=%s" % (i + 1, v) in out
    Q.update(repeat=1000)
    out = capsys.readouterr().out
    assert "Iteration 14: loglike=-1.396481e+02" in out
    assert Q.converged and abs(Q.iter - 488) <= 2
    assert "Converged at iteration %d." % Q.iter in out
    np.testing.assert_allclose(Q.L[Q.iter - 1], -1.224106e+02, rtol=2e-6)
    Q.update(repeat=10000, tol=1e-6)
    out = capsys.readouterr().out
    assert Q.converged and abs(Q.iter - 847) <= 5
    assert "Converged at iteration %d." % Q.iter in out
    np.testing.assert_allclose(Q.L[Q.iter - 1], -1.222506e+02, rtol=2e-6)


def test_hmm_doc_example_known_and_unknown_parameters(backend, capsys):
    """hmm.rst:8-11,35-96 ('Iteration 1: loglike=-1.095883e+02', the exact posterior of the weather chain) and
    :150-290 ('Iteration 1: loglike=-9.963054e+02 ... Iteration 8: loglike=-9.235053e+02, Converged at iteration 8'):
    a categorical Markov chain under a categorical and then a Gaussian mixture, with seeded draws from the chain and
    from the mixture (host RNG in the reference's order)."""
    from bayespy_b200.nodes import CategoricalMarkovChain, Categorical, Mixture, Dirichlet, Gaussian
    from bayespy_b200.inference import VB
    np.random.seed(1)
    a0 = [0.6, 0.4]
    A = [[0.7, 0.3], [0.4, 0.6]]
    N = 100
    Z = CategoricalMarkovChain(a0, A, states=N)
    P = [[0.1, 0.4, 0.5], [0.6, 0.3, 0.1]]
    Y = Mixture(Z, Categorical, P)
    weather = Z.random()
    activity = Mixture(weather, Categorical, P).random()
    Y.observe(activity)
    Q = VB(Y, Z)
    Q.update()
    L = _loglikes(capsys.readouterr().out)
    np.testing.assert_allclose(L, [-1.095883e+02], rtol=5e-7)
    pz = np.asarray(Z._to_categorical().get_moments()[0])
    assert pz.shape == (N, 2) and np.allclose(pz.sum(axis=-1), 1.0)
    # unknown parameters, Gaussian emissions
    mu = np.array([[0, 0], [3, 4], [6, 0]])
    std, K, N = 2.0, 3, 200
    p0 = np.ones(K) / K
    q = 0.9
    r = (1 - q) / (K - 1)
    P = q * np.identity(K) + r * (np.ones((3, 3)) - np.identity(3))
    y = np.zeros((N, 2))
    z = np.zeros(N)
    state = np.random.choice(K, p=p0)
    for n in range(N):
        z[n] = state
        y[n, :] = std * np.random.randn(2) + mu[state]
        state = np.random.choice(K, p=P[state])
    a0 = Dirichlet(1e-3 * np.ones(K))
    A = Dirichlet(1e-3 * np.ones((K, K)))
    Z = CategoricalMarkovChain(a0, A, states=N)
    Lambda = std ** (-2) * np.identity(2)
    Y = Mixture(Z, Gaussian, mu, Lambda)
    Y.observe(y)
    Q = VB(Y, Z, A, a0)
    Q.update(repeat=1000)
    out = capsys.readouterr().out
    L = _loglikes(out)
    assert len(L) == 8 and "Converged at iteration 8." in out
    np.testing.assert_allclose(L[0], -9.963054e+02, rtol=5e-7)
    np.testing.assert_allclose(L[7], -9.235053e+02, rtol=5e-7)
    # the chain recovers most of the simulated states
    zhat = np.argmax(np.asarray(Z._to_categorical().get_moments()[0]), axis=-1)
    assert np.mean(zhat == z) > 0.9


def test_bmm_doc_example(oracle_backend, capsys):
    """bmm.rst:8-11,28-125: a Bernoulli mixture with Beta-distributed success probabilities —
    'Iteration 1: loglike=-6.872145e+02 ... Iteration 17: loglike=-5.236921e+02, Converged at iteration 17'.
    (Host-logic test: the Beta / Bernoulli nodes are the two-category case of the Dirichlet / Multinomial kernels.)"""
    from bayespy_b200.nodes import Categorical, Dirichlet, Beta, Mixture, Bernoulli
    from bayespy_b200.inference import VB
    from bayespy_b200.utils import random
    np.random.seed(1)
    p0 = [0.1, 0.9, 0.1, 0.9, 0.1, 0.9, 0.1, 0.9, 0.1, 0.9]
    p1 = [0.1, 0.1, 0.1, 0.1, 0.1, 0.9, 0.9, 0.9, 0.9, 0.9]
    p2 = [0.9, 0.9, 0.9, 0.9, 0.9, 0.1, 0.1, 0.1, 0.1, 0.1]
    p = np.array([p0, p1, p2])
    z = random.categorical([1 / 3, 1 / 3, 1 / 3], size=100)
    x = random.bernoulli(p[z])
    N, D, K = 100, 10, 10
    R = Dirichlet(K * [1e-5], name="R")
    Z = Categorical(R, plates=(N, 1), name="Z")
    P = Beta([0.5, 0.5], plates=(D, K), name="P")
    X = Mixture(Z, Bernoulli, P)
    Q = VB(Z, R, X, P)
    P.initialize_from_random()
    X.observe(x)
    Q.update(repeat=1000)
    out = capsys.readouterr().out
    L = _loglikes(out)
    assert len(L) == 17 and "Converged at iteration 17." in out
    np.testing.assert_allclose(L[0], -6.872145e+02, rtol=5e-7)
    np.testing.assert_allclose(L[16], -5.236921e+02, rtol=5e-7)
