"""Model-level parity against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py): posterior u / phi / g of every node and the per-sweep
lower bound Q.L, rtol 1e-6 (north_star: 1e-5).  backend='oracle' checks the host-side
graph logic on CPU; backend='cuda' (-m gpu) is the end-to-end parity test of the kernels."""
import numpy as np
import pytest

from conftest import golden

RTOL = 1e-6


def close(a, b, rtol=RTOL, atol=1e-9):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def check_node(g, name, node, skip_g=False):
    for i in range(len(node.u)):
        close(node.u[i], g["%s_u%d" % (name, i)])
        close(node.phi[i], g["%s_phi%d" % (name, i)])
    if not skip_g:
        close(node.g, g["%s_g" % name])


def test_quickstart(backend, capsys):
    """doc/source/user_guide/quickstart.rst:8-118 — loglike -6.020956e+01 ... converged at 4."""
    from bayespy_b200.nodes import GaussianARD, Gamma
    from bayespy_b200.inference import VB
    g = golden("quickstart")
    np.random.seed(1)
    data = np.random.normal(5, 10, size=(10,))
    assert np.array_equal(data, g["data"])
    mu = GaussianARD(0, 1e-6)
    tau = Gamma(1e-6, 1e-6)
    y = GaussianARD(mu, tau, plates=(10,))
    y.observe(data)
    Q = VB(mu, tau, y)
    Q.update(repeat=20)
    out = capsys.readouterr().out
    assert "Iteration 1: loglike=-6.020956e+01" in out
    assert "Iteration 4: loglike=-5.820288e+01" in out
    assert "Converged at iteration 4." in out
    assert Q.iter == int(g["iters"])
    close(Q.L[:Q.iter], g["L"], rtol=1e-9)
    check_node(g, "mu", mu)
    check_node(g, "tau", tau)


def build_pca(g, M, N, K, fused, seeded_init=True):
    from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
    from bayespy_b200.utils import random
    np.random.seed(1)
    w = np.random.randn(M, 4)
    x = np.random.randn(N, 4)
    y = w @ x.T + 0.1 * np.random.randn(M, N)
    assert np.array_equal(y, g["y"])
    X = GaussianARD(0, 1, plates=(1, N), shape=(K,), name="X")
    alpha = Gamma(1e-5, 1e-5, plates=(K,), name="alpha")
    C = GaussianARD(0, alpha, plates=(M, 1), shape=(K,), name="C")
    F = SumMultiply("d,d->", X, C)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    if "mask" in g.files:
        mask = random.mask(M, N, p=0.8)
        assert np.array_equal(mask, g["mask"])
        Y.observe(y, mask=mask)
    else:
        Y.observe(y)
    if seeded_init:
        C.initialize_from_random()          # host RNG parity with the reference stream
        close(C.u[0], g["C_init"], rtol=1e-12)
    else:
        C.initialize_from_value(g["C_init"])
    Q = VB(Y, X, C, alpha, tau, fused=fused)
    return Q, dict(X=X, C=C, alpha=alpha, tau=tau, Y=Y, F=F)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name,M,N,K", [("pca_small", 20, 100, 5), ("pca_64x16", 64, 96, 16)])
def test_pca(backend, name, M, N, K, fused):
    g = golden(name)
    Q, n = build_pca(g, M, N, K, fused)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    close(Q.L[:iters], g["L"], rtol=1e-8)
    for nm in ("Y", "X", "C", "alpha", "tau"):
        close(Q.l[n[nm]][:iters], g["l_" + nm], rtol=1e-7, atol=1e-7)
    for nm in ("X", "C", "alpha", "tau"):
        check_node(g, nm, n[nm])
    if fused:
        assert len(Q.plans) == 1 and Q.plans[0].fused_calls > 0
    # composite messages arriving at the stochastic parents
    mC = n["C"].message_from_children()
    close(mC[0], g["msgC0"]); close(mC[1], g["msgC1"])
    mt = n["tau"].message_from_children()
    close(mt[0], g["msgtau0"]); close(mt[1], g["msgtau1"])


def test_pca_masked_generic(backend):
    """Missing values (random.mask, p=0.8): per-plate covariances, masks and broadcasting."""
    g = golden("pca_masked")
    Q, n = build_pca(g, 12, 40, 4, fused=False)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    close(Q.L[:iters], g["L"], rtol=1e-8)
    for nm in ("X", "C", "alpha", "tau"):
        check_node(g, nm, n[nm])


@pytest.mark.parametrize("name,M,N,K", [("pca_masked", 12, 40, 4), ("pca_masked_64x16", 64, 300, 16)])
def test_pca_masked_fused(backend, name, M, N, K):
    """Missing values through the FUSED masked sweep (csrc/pca_masked.cu: masked tensor-pipe GEMMs + one thread per
    column for the K x K inverse; nothing of size (N,K,K) is stored) against the reference's trajectory."""
    g = golden(name)
    Q, n = build_pca(g, M, N, K, fused=True)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    assert Q.plans[0].masked() and Q.plans[0].fused_calls >= 3 * iters - 2
    close(Q.L[:iters], g["L"], rtol=1e-8)
    for nm in ("X", "C", "alpha", "tau"):
        check_node(g, nm, n[nm])
    # the latent (unobserved) entries of Y are produced on demand and equal the per-node tier's
    Q0, n0 = build_pca(g, M, N, K, fused=False)
    Q0.update(repeat=iters, verbose=False, tol=0)
    for i in range(2):
        close(np.asarray(n["Y"].u[i]), np.asarray(n0["Y"].u[i]), rtol=1e-7)


def test_pca_masked_fused_any_update_order(backend):
    """Stale masked statistics (somebody else changed X, or C is updated first) fall back to the generic messages."""
    g = golden("pca_masked")
    res = []
    for fused in (False, True):
        Q, n = build_pca(g, 12, 40, 4, fused)
        for _ in range(3):
            Q.update("C", "tau", "X", "alpha", "X", "C", verbose=False, tol=0)
        res.append((Q.L[:3].copy(), np.asarray(n["X"].u[0]), np.asarray(n["C"].u[1]), np.asarray(n["X"].u[1])))
    close(res[0][0], res[1][0], rtol=1e-9)
    close(res[0][1], res[1][1], rtol=1e-8)
    close(res[0][2], res[1][2], rtol=1e-8)
    close(res[0][3], res[1][3], rtol=1e-7)


def test_pca_update_order_independent_of_plan(backend):
    """Arbitrary user update orders give the same answer with and without the fused plan
    (stale statistics must be detected through the version tags)."""
    g = golden("pca_small")
    order = ("tau", "C", "X", "alpha", "C", "tau", "X")
    res = []
    for fused in (False, True):
        Q, n = build_pca(g, 20, 100, 5, fused)
        for _ in range(3):
            Q.update(*order, verbose=False, tol=0)
        res.append((Q.L[:3].copy(), np.asarray(n["X"].u[0]), np.asarray(n["C"].u[1])))
    close(res[0][0], res[1][0], rtol=1e-9)
    close(res[0][1], res[1][1], rtol=1e-8)
    close(res[0][2], res[1][2], rtol=1e-8)


def test_vb_api(backend, capsys):
    """VB bookkeeping: duplicate removal, lookup by name, tol / converged, histories."""
    from bayespy_b200.nodes import GaussianARD, Gamma
    from bayespy_b200.inference import VB
    np.random.seed(3)
    mu = GaussianARD(0, 1e-3, name="mu")
    tau = Gamma(1e-3, 1e-3, name="tau")
    y = GaussianARD(mu, tau, plates=(50,), name="y")
    y.observe(np.random.randn(50) + 2)
    Q = VB(y, mu, tau, mu)
    assert Q.model == [y, mu, tau]
    assert Q["mu"] is mu
    with pytest.raises(ValueError):
        VB(y, 3)
    Q.update(repeat=100, tol=1e-8, verbose=False)
    assert Q.has_converged() and Q.iter < 100
    assert len(Q.L) >= Q.iter and np.all(np.diff(Q.L[:Q.iter]) > -1e-6)
    assert abs(Q.compute_lowerbound() - Q.L[Q.iter - 1]) < 1e-8 * abs(Q.L[Q.iter - 1])
    terms = Q.compute_lowerbound_terms()
    assert abs(sum(terms.values()) - Q.L[Q.iter - 1]) < 1e-8 * abs(Q.L[Q.iter - 1])


def build_gmm(g, N, D, K, fused, y=None):
    from bayespy_b200.nodes import Gaussian, Wishart, Dirichlet, Categorical, Mixture
    from bayespy_b200.inference import VB
    if y is None:
        np.random.seed(1)
        means = 5 * np.random.randn(K, D)
        z = np.random.randint(0, K, size=N)
        y = means[z] + np.random.randn(N, D)
    assert np.array_equal(y, g["y"])
    alpha = Dirichlet(1e-5 * np.ones(K), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="Z")
    mu = Gaussian(np.zeros(D), 1e-5 * np.identity(D), plates=(K,), name="mu")
    Lambda = Wishart(D, 1e-5 * np.identity(D), plates=(K,), name="Lambda")
    Y = Mixture(Z, Gaussian, mu, Lambda, name="Y")
    Z.initialize_from_random()                       # host RNG parity (categorical one-hot: bit-exact)
    assert np.array_equal(np.asarray(Z.u[0]), g["Z_init"])
    return Y, Z, mu, Lambda, alpha, y


@pytest.mark.parametrize("fused", [False, True])
def test_gmm_small(backend, fused):
    """Mixture + Gaussian + Wishart + Dirichlet + Categorical (gmm.rst:71-98 on 5 blobs)."""
    from bayespy_b200.inference import VB
    g = golden("gmm_small")
    Y, Z, mu, Lambda, alpha, y = build_gmm(g, 300, 3, 5, fused)
    Y.observe(y)
    Q = VB(Y, mu, Lambda, Z, alpha, fused=fused)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    close(Q.L[:iters], g["L"], rtol=1e-8)
    for nm, node in (("Z", Z), ("mu", mu), ("Lambda", Lambda), ("alpha", alpha)):
        check_node(g, nm, node)
        close(Q.l[node][:iters], g["l_" + nm], rtol=1e-7, atol=1e-6)
    close(Q.l[Y][:iters], g["l_Y"], rtol=1e-8)
    if fused:
        assert len(Q.plans) == 1 and Q.plans[0].fused_calls > 0


@pytest.mark.parametrize("fused", [False, True])
def test_gmm_doc_example(backend, fused, capsys):
    """doc/source/examples/gmm.rst:8-118: 'Iteration 1: loglike=-1.402345e+03' ...
    'Iteration 61: loglike=-8.888464e+02', 'Converged at iteration 61.'"""
    from bayespy_b200.nodes import Gaussian, Wishart, Dirichlet, Categorical, Mixture
    from bayespy_b200.inference import VB
    g = golden("gmm_doc")
    np.random.seed(1)
    y0 = np.random.multivariate_normal([0, 0], [[2, 0], [0, 0.1]], size=50)
    y1 = np.random.multivariate_normal([0, 0], [[0.1, 0], [0, 2]], size=50)
    y2 = np.random.multivariate_normal([2, 2], [[2, -1.5], [-1.5, 2]], size=50)
    y3 = np.random.multivariate_normal([-2, -2], [[0.5, 0], [0, 0.5]], size=50)
    y = np.vstack([y0, y1, y2, y3])
    assert np.array_equal(y, g["y"])
    N, D, K = 200, 2, 10
    alpha = Dirichlet(1e-5 * np.ones(K), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="z")
    mu = Gaussian(np.zeros(D), 1e-5 * np.identity(D), plates=(K,), name="mu")
    Lambda = Wishart(D, 1e-5 * np.identity(D), plates=(K,), name="Lambda")
    Y = Mixture(Z, Gaussian, mu, Lambda, name="Y")
    Z.initialize_from_random()
    assert np.array_equal(np.asarray(Z.u[0]), g["Z_init"])
    Q = VB(Y, mu, Lambda, Z, alpha, fused=fused)
    Y.observe(y)
    Q.update(repeat=1000)
    out = capsys.readouterr().out
    assert "Iteration 1: loglike=-1.402345e+03" in out
    assert "Iteration 61: loglike=-8.888464e+02" in out
    assert "Converged at iteration 61." in out
    close(Q.L[:Q.iter], g["L"], rtol=1e-7)


def test_mixture_of_gaussian_ard_components(backend):
    """mixture.py:26-488 with a mixed class other than Gaussian (per-node path: no fused plan for it)."""
    from bayespy_b200.nodes import GaussianARD, Gamma, Dirichlet, Categorical, Mixture
    from bayespy_b200.inference import VB
    g = golden("mixture_ard")
    N, K = len(g["y"]), 3
    alpha = Dirichlet(1e-3 * np.ones(K), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="Z")
    mu = GaussianARD(0, 1e-3, plates=(K,), name="mu")
    tau = Gamma(1e-3, 1e-3, plates=(K,), name="tau")
    Y = Mixture(Z, GaussianARD, mu, tau, name="Y")
    Z.initialize_from_value(np.argmax(g["Z_init"], axis=-1))
    Y.observe(g["y"])
    Q = VB(Y, mu, tau, Z, alpha)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    np.testing.assert_allclose(np.asarray(Z.u[0]), g["Z_u0"], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(np.asarray(mu.u[0]), g["mu_u0"], rtol=1e-8)
    np.testing.assert_allclose(np.asarray(tau.u[0]), g["tau_u0"], rtol=1e-8)
