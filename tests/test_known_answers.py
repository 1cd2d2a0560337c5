"""Closed-form known answers, following the reference's own test strategy: conjugate posteriors,
exact marginal likelihoods and special-function identities that pin the node arithmetic independently of the
golden vectors.  Each test cites the reference test it mirrors (bayespy/inference/vmp/nodes/tests/...)."""
import numpy as np
import pytest
from scipy import special


def f(x):
    return float(np.asarray(x))


def test_gamma_observed_bound_and_student_t_marginal(backend):
    """test_gamma.py:37-88: log-density of an observed Gamma; with a single latent tau the VB bound of
    Y ~ N(x, 1/tau), tau ~ Gamma(a, b) is the exact Student-t marginal likelihood."""
    from bayespy_b200.nodes import Gamma, GaussianARD
    a, b, y = 15.0, 21.0, 4.0
    x = Gamma(a, b)
    x.observe(y)
    np.testing.assert_allclose(f(x.lower_bound_contribution()),
                               a * np.log(b) + (a - 1) * np.log(y) - b * y - special.gammaln(a), rtol=1e-12)
    a, b, xm, y = 2.3, 4.1, 1.9, 4.8
    tau = Gamma(a, b)
    Y = GaussianARD(xm, tau)
    Y.observe(y)
    tau.update()
    np.testing.assert_allclose([f(tau.phi[0]), f(tau.phi[1])], [-(b + 0.5 * (y - xm) ** 2), a + 0.5], rtol=1e-13)
    nu, s2 = 2 * a, b / a
    exact = (special.gammaln((nu + 1) / 2) - special.gammaln(nu / 2) - 0.5 * np.log(nu) - 0.5 * np.log(np.pi)
             - 0.5 * np.log(s2) - 0.5 * (nu + 1) * np.log(1 + (y - xm) ** 2 / (nu * s2)))
    np.testing.assert_allclose(f(Y.lower_bound_contribution()) + f(tau.lower_bound_contribution()), exact, rtol=1e-11)


def test_dirichlet_moments_and_constant(backend):
    """test_dirichlet.py:68-91."""
    from bayespy_b200.nodes import Dirichlet
    p = Dirichlet([2, 3, 4])
    np.testing.assert_allclose(np.asarray(p.get_moments()[0]), special.psi([2, 3, 4]) - special.psi(9), rtol=1e-12)
    p = Dirichlet([1, 1, 1])
    p.initialize_from_value([0.5, 0.4, 0.1])
    np.testing.assert_allclose(np.asarray(p.get_moments()[0]), np.log([0.5, 0.4, 0.1]), rtol=1e-13)


def test_wishart_moments(backend):
    """test_wishart.py:79-125: <Lambda> = n V^-1, <log|Lambda|> = psi_D(n/2) + D log 2 - log|V|."""
    from bayespy_b200.nodes import Wishart
    rs = np.random.RandomState(0)
    for Dm in (1, 3, 6):
        R = rs.randn(Dm, Dm)
        V = R @ R.T + Dm * np.identity(Dm)
        n = Dm + 2.5
        W = Wishart(n, V)
        u = W.get_moments()
        np.testing.assert_allclose(np.asarray(u[0]), n * np.linalg.inv(V), rtol=1e-10)
        mdg = sum(special.psi(n / 2 - 0.5 * i) for i in range(Dm))
        np.testing.assert_allclose(f(u[1]), mdg + Dm * np.log(2) - np.linalg.slogdet(V)[1], rtol=1e-11)


def test_gaussian_conjugate_posterior_and_exact_evidence(backend):
    """test_gaussian.py:299-387 / :692-779 pattern: mu ~ N(m0, L0^-1), y_n ~ N(mu, Lam^-1) with Lam fixed.
    One latent node -> after one update the posterior is exact and the VB bound is the exact log evidence."""
    from bayespy_b200.nodes import Gaussian
    rs = np.random.RandomState(1)
    Dm, N = 3, 7
    m0 = rs.randn(Dm)
    R = rs.randn(Dm, Dm); L0 = R @ R.T + np.identity(Dm)
    R = rs.randn(Dm, Dm); Lam = R @ R.T + np.identity(Dm)
    y = rs.randn(N, Dm)
    mu = Gaussian(m0, L0)
    Y = Gaussian(mu, Lam, plates=(N,))
    Y.observe(y)
    mu.update()
    Lp = L0 + N * Lam
    mp = np.linalg.solve(Lp, L0 @ m0 + Lam @ y.sum(0))
    np.testing.assert_allclose(np.asarray(mu.u[0]), mp, rtol=1e-10)
    np.testing.assert_allclose(np.asarray(mu.u[1]), np.linalg.inv(Lp) + np.outer(mp, mp), rtol=1e-10)
    # exact log p(y) = log N(vec(y) | 1 (x) m0, I (x) Lam^-1 + 11^T (x) L0^-1)
    C = np.kron(np.identity(N), np.linalg.inv(Lam)) + np.kron(np.ones((N, N)), np.linalg.inv(L0))
    r = (y - m0).ravel()
    exact = -0.5 * r @ np.linalg.solve(C, r) - 0.5 * np.linalg.slogdet(C)[1] - 0.5 * N * Dm * np.log(2 * np.pi)
    np.testing.assert_allclose(f(Y.lower_bound_contribution()) + f(mu.lower_bound_contribution()), exact, rtol=1e-10)


def test_gaussian_ard_missing_values_only_count_observed_plates(backend):
    """node.py:570-655 masks: plates masked out of an observation neither send messages nor enter the bound."""
    from bayespy_b200.nodes import GaussianARD, Gamma
    rs = np.random.RandomState(2)
    N = 11
    y = rs.randn(N)
    mask = rs.rand(N) < 0.6
    mu = GaussianARD(0.3, 2.0)
    tau = Gamma(1.5, 0.7)
    tau.initialize_from_value(1.3)
    Y = GaussianARD(mu, tau, plates=(N,))
    Y.observe(y, mask=mask)
    mu.update()
    n_obs, t = mask.sum(), 1.3
    prec = 2.0 + t * n_obs
    np.testing.assert_allclose(f(mu.u[0]), (2.0 * 0.3 + t * y[mask].sum()) / prec, rtol=1e-12)
    np.testing.assert_allclose(f(mu.u[1]) - f(mu.u[0]) ** 2, 1 / prec, rtol=1e-11)
    m, v = f(mu.u[0]), 1 / prec
    expect = np.sum(-0.5 * t * ((y[mask] - m) ** 2 + v)) + n_obs * (0.5 * np.log(t) - 0.5 * np.log(2 * np.pi))
    np.testing.assert_allclose(f(Y.lower_bound_contribution()), expect, rtol=1e-11)


def test_one_component_mixture_equals_the_gaussian(backend):
    """mixture.py:53-160 with K = 1: responsibilities are exactly one, the mixture's bound term equals the plain
    Gaussian log-density and the messages to mu are the Gaussian's."""
    from bayespy_b200.nodes import Categorical, Mixture, Gaussian
    rs = np.random.RandomState(3)
    Dm, N = 2, 9
    y = rs.randn(N, Dm)
    Lam = np.array([[2.0, 0.3], [0.3, 1.5]])
    mu = Gaussian(np.zeros(Dm), np.identity(Dm), plates=(1,))
    Z = Categorical([1.0], plates=(N,))
    Y = Mixture(Z, Gaussian, mu, Lam[None])
    Y.observe(y)
    Z.update()
    np.testing.assert_allclose(np.asarray(Z.u[0]), np.ones((N, 1)), rtol=1e-13)
    mu.update()
    Lp = np.identity(Dm) + N * Lam
    np.testing.assert_allclose(np.asarray(mu.u[0])[0], np.linalg.solve(Lp, Lam @ y.sum(0)), rtol=1e-10)


def test_categorical_one_hot_is_exact(backend):
    """categorical.py:30-47: integer labels -> one-hot moments, bit-exact (north star: index work is exact)."""
    from bayespy_b200.nodes import Categorical
    rs = np.random.RandomState(4)
    K, N = 7, 1000
    z = rs.randint(0, K, size=N)
    Z = Categorical(np.ones(K) / K, plates=(N,))
    Z.observe(z)
    u = np.asarray(Z.u[0])
    assert u.dtype == np.float64 and np.array_equal(u, np.eye(K)[z])
    np.testing.assert_allclose(f(Z.lower_bound_contribution()), N * np.log(1.0 / K), rtol=1e-12)
