"""Differential tests: the same model script is built twice — with the unmodified reference (oracle/_ref) and with this
package — from the same seeds, updated for a few iterations, and every bound is compared.  The models sit at the edges
of the node set: chains driven by inputs (time-varying dynamics, masks, constant inputs), Gaussian-gamma variables
under SumMultiply / Take / Gate / Mixture / scalar plates, a categorical Markov chain with per-step transitions over
plates and as the selector of a switching state-space model, multinomial counts over plates.

Host-logic tests (plates, broadcasting, message routing): they run on the oracle backend; the kernels underneath are
the ones the ``-m gpu`` parity tests exercise."""
import numpy as np
import pytest


@pytest.fixture
def both(oracle_backend):
    from oracle import make_ref
    make_ref.build()
    if not make_ref.available():
        pytest.skip("oracle/_ref is not staged and /root/reference is absent")
    make_ref.import_reference()
    import bayespy.inference as ref_inference
    import bayespy.nodes as ref_nodes
    import bayespy_b200.inference as our_inference
    import bayespy_b200.nodes as our_nodes
    return (ref_nodes, ref_inference), (our_nodes, our_inference)


def chain_inputs_time_varying(N, rs):
    D_, K, T = 2, 2, 6
    U = rs.randn(T - 1, K)
    A = N.GaussianARD(0, 1.0, shape=(D_ + K,), plates=(T - 1, D_), name="A")
    A.initialize_from_value(0.3 * rs.randn(T - 1, D_, D_ + K))
    nu = N.Gamma(2.0, 2.0, plates=(T - 1, D_), name="nu")
    X = N.GaussianMarkovChain(np.zeros(D_), np.identity(D_), A, nu, inputs=U, name="X")
    Y = N.Gaussian(X, 3.0 * np.identity(D_), name="Y")
    Y.observe(rs.randn(T, D_), mask=rs.rand(T) < 0.8)
    return [X, A, nu, Y]


def gaussian_gamma_product(N, rs):
    Kd, P, M = 3, 4, 5
    b1 = N.Gamma(2.0, 2.0, plates=(P, 1), name="b1")
    W = N.GaussianGamma(np.zeros(Kd), np.identity(Kd), 3.0, b1, plates=(P, 1), name="W")
    b2 = N.Gamma(1.5, 1.0, plates=(1, M), name="b2")
    Z = N.GaussianGamma(np.ones(Kd), 2 * np.identity(Kd), 2.0, b2, plates=(1, M), name="Z")
    F = N.SumMultiply("k,k->", W, Z, name="F")
    Y = N.GaussianARD(F, 1.0, name="Y")
    Y.observe(rs.randn(P, M))
    return [Y, W, Z, b1, b2]


def gaussian_gamma_times_constant_and_gaussian(N, rs):
    Kd, P = 3, 6
    b1 = N.Gamma(2.0, 2.0, plates=(P,), name="b1")
    W = N.GaussianGamma(np.zeros(Kd), np.identity(Kd), 3.0, b1, plates=(P,), name="W")
    X = N.GaussianARD(0, 1, shape=(Kd,), plates=(P,), name="X")
    X.initialize_from_value(rs.randn(P, Kd))
    F = N.SumMultiply("k,k,k->", W, X, rs.randn(P, Kd), name="F")
    tau = N.Gamma(1e-2, 1e-2, name="tau")
    Y = N.GaussianARD(F, tau, name="Y")
    Y.observe(rs.randn(P))
    return [Y, W, X, b1, tau]


def gaussian_gamma_taken_by_index(N, rs):
    G, Nn, Dm = 3, 12, 2
    b = N.Gamma(2.0, 2.0, plates=(G,), name="b")
    W = N.GaussianGamma(np.zeros(Dm), np.identity(Dm), 3.0, b, plates=(G,), name="W")
    m = N.Take(W, rs.randint(0, G, size=Nn), name="m")
    Y = N.Gaussian(m, 2.0 * np.identity(Dm), name="Y")
    Y.observe(rs.randn(Nn, Dm))
    return [Y, W, b]


def gaussian_gamma_gated(N, rs):
    K, Nn, Dm = 3, 10, 2
    b = N.Gamma(2.0, 2.0, plates=(K,), name="b")
    W = N.GaussianGamma(np.zeros(Dm), np.identity(Dm), 3.0, b, plates=(K,), name="W")
    Z = N.Categorical(np.ones(K) / K, plates=(Nn,), name="Z")
    Y = N.Gaussian(N.Gate(Z, W, name="m"), 2.0 * np.identity(Dm), name="Y")
    Y.observe(rs.randn(Nn, Dm))
    return [Y, W, b, Z]


def mixture_with_gaussian_gamma_means(N, rs):
    K, Nn, Dm = 3, 15, 2
    mu = N.GaussianGamma(np.zeros(Dm), 0.1 * np.identity(Dm), 2.0, 2.0, plates=(K,), name="mu")
    L = N.Wishart(Dm + 1.0, np.identity(Dm), plates=(K,), name="L")
    Z = N.Categorical(N.Dirichlet(np.ones(K), name="al"), plates=(Nn,), name="Z")
    Z.initialize_from_value(rs.randint(0, K, size=Nn))
    Y = N.Mixture(Z, N.Gaussian, mu, L, name="Y")
    Y.observe(rs.randn(Nn, Dm))
    return [Y, mu, L, Z]


def scalar_gaussian_gamma_over_plates(N, rs):
    P, M = 3, 4
    lam = N.Gamma(2.0, 1.0, plates=(P, 1), name="lam")
    b = N.Gamma(1.0, 1.0, plates=(1, M), name="b")
    mc = N.GaussianGamma(rs.randn(P, M), lam.as_wishart(ndim=0), 1.5, b, ndim=0, name="mc")
    Y = N.GaussianARD(mc, 2.0, plates=(5, P, M), name="Y")
    Y.observe(rs.randn(5, P, M), mask=rs.rand(5, P, M) < 0.8)
    return [Y, mc, lam, b]


def hidden_markov_chains_with_per_step_transitions(N, rs):
    Kc, T, P = 3, 9, 2
    a0 = N.Dirichlet(np.ones(Kc), plates=(P,), name="a0")
    A = N.Dirichlet(np.ones(Kc), plates=(P, T - 1, Kc), name="A")
    Zc = N.CategoricalMarkovChain(a0, A, name="Zc")
    mu = N.GaussianARD(0, 1e-1, shape=(2,), plates=(Kc,), name="mu")
    mu.initialize_from_value(rs.randn(Kc, 2))
    Y = N.Mixture(Zc, N.Gaussian, mu, np.identity(2), name="Y")
    Y.observe(rs.randn(P, T, 2))
    return [Y, Zc, mu, A, a0]


def switching_state_space_model_selected_by_a_markov_chain(N, rs):
    Dm, T, K = 2, 10, 2
    a0 = N.Dirichlet(np.ones(K), name="a0")
    A = N.Dirichlet(np.ones(K), plates=(K,), name="A")
    Zc = N.CategoricalMarkovChain(a0, A, states=T - 1, name="Zc")
    B = N.GaussianARD(0, 1.0, shape=(Dm,), plates=(K, Dm), name="B")
    B.initialize_from_value(0.4 * rs.randn(K, Dm, Dm))
    X = N.SwitchingGaussianMarkovChain(np.zeros(Dm), np.identity(Dm), B, Zc, np.ones(Dm), n=T, name="X")
    Y = N.Gaussian(X, 4.0 * np.identity(Dm), name="Y")
    Y.observe(rs.randn(T, Dm))
    return [X, B, Zc, A, a0, Y]


def multinomial_counts_over_plates(N, rs):
    Kc, P = 4, 5
    p = N.Dirichlet(np.ones(Kc), plates=(P,), name="p")
    n = rs.randint(3, 9, size=(6, P))
    X = N.Multinomial(n, p, name="Xm")
    X.observe(np.array([[rs.multinomial(n[i, j], np.ones(Kc) / Kc) for j in range(P)] for i in range(6)]))
    return [X, p]


def poisson_counts_with_gamma_rates(N, rs):
    lam = N.Gamma(2.0, 1.0, plates=(4,), name="lam")
    X = N.Poisson(lam, plates=(10, 4), name="X")
    X.observe(rs.poisson(3.0, size=(10, 4)), mask=rs.rand(10, 4) < 0.9)
    return [X, lam]


def bernoulli_mixture(N, rs):
    K, Nn, Dd = 3, 20, 4
    Z = N.Categorical(N.Dirichlet(np.ones(K), name="R"), plates=(Nn, 1), name="Z")
    Z.initialize_from_value(rs.randint(0, K, size=(Nn, 1)))
    P = N.Beta([0.5, 0.5], plates=(Dd, K), name="P")
    X = N.Mixture(Z, N.Bernoulli, P, name="X")
    X.observe(rs.rand(Nn, Dd) < 0.4)
    return [X, P, Z]


def binomial_counts_over_plates(N, rs):
    p = N.Beta([1.0, 2.0], plates=(3,), name="p")
    n = rs.randint(2, 9, size=(5, 3))
    X = N.Binomial(n, p, name="X")
    X.observe(rs.binomial(n, 0.3))
    return [X, p]


def sum_of_independent_gaussians(N, rs):
    X = N.Gaussian(np.zeros(2), np.identity(2), plates=(3,), name="X")
    Y = N.GaussianARD(1.0, 2.0, shape=(2,), name="Y2")
    Z = N.Add(X, Y, rs.randn(3, 2), name="Z")
    tau = N.Gamma(1e-2, 1e-2, name="tau")
    W = N.GaussianARD(Z, tau, ndim=1, plates=(4, 3), name="W")
    W.observe(rs.randn(4, 3, 2), mask=rs.rand(4, 3) < 0.8)
    return [W, X, Y, tau]


def concatenated_groups(N, rs):
    a = N.GaussianARD(0, 1.0, shape=(2,), plates=(3,), name="a")
    b = N.GaussianARD(1, 2.0, shape=(2,), plates=(4,), name="b")
    Z = N.Concatenate(a, b, name="Z")
    tau = N.Gamma(1e-2, 1e-2, name="tau")
    Y = N.GaussianARD(Z, tau, ndim=1, plates=(5, 7), name="Y")
    Y.observe(rs.randn(5, 7, 2))          # (a mask through Concatenate raises in the reference: its mask split is off)
    return [Y, a, b, tau]


def mixture_over_two_cluster_axes(N, rs):
    K1, K2, Nn = 2, 3, 12
    z1 = N.Categorical(N.Dirichlet(np.ones(K1), name="r1"), plates=(Nn,), name="z1")
    z2 = N.Categorical(N.Dirichlet(np.ones(K2), name="r2"), plates=(Nn,), name="z2")
    z1.initialize_from_value(rs.randint(0, K1, size=Nn))
    z2.initialize_from_value(rs.randint(0, K2, size=Nn))
    mu = N.GaussianARD(0, 1e-1, plates=(K1, K2), name="mu")
    mu.initialize_from_value(rs.randn(K1, K2))
    X = N.MultiMixture([z1, z2], N.GaussianARD, mu, 2.0, name="X")
    X.observe(rs.randn(Nn))
    return [X, mu, z1, z2]


def chosen_and_complemented(N, rs):
    x0, x1, x2 = (N.GaussianARD(m, 1.0, name="x%d" % i) for i, m in enumerate((0.0, 10.0, 20.0)))
    z = N.Categorical(np.ones(3) / 3, plates=(6,), name="z")
    Y = N.GaussianARD(N.Choose(z, x0, x1, x2), 2.0, name="Y")
    Y.observe(rs.randn(6) * 5 + 10)
    p = N.Beta([2.0, 3.0], name="p")
    B = N.Bernoulli(p.complement(), plates=(8,), name="B")
    B.observe(rs.rand(8) < 0.3)
    return [Y, z, x0, x1, x2, B, p]


def gamma_shape_point_estimate(N, rs):
    a = N.GammaShape(name="a")
    b = N.Gamma(1e-5, 1e-5, name="b")
    tau = N.Gamma(a, b, plates=(200,), name="tau")
    tau.observe(rs.gamma(10.0, 1 / 20.0, size=200))
    return [tau, a, b]


def dirichlet_concentration_point_estimate(N, rs):
    conc = N.DirichletConcentration(4, name="conc")
    p = N.Dirichlet(conc, plates=(25,), name="p")
    X = N.Multinomial(40, p, name="X")
    X.observe(np.array([rs.multinomial(40, pk) for pk in rs.dirichlet([2.0, 5.0, 1.0, 3.0], size=25)]))
    return [X, p, conc]


def masked_hidden_markov_chains_over_plates(N, rs):
    Kc, T, P = 3, 8, 4
    a0 = N.Dirichlet(np.ones(Kc), name="a0")
    A = N.Dirichlet(np.ones(Kc), plates=(Kc,), name="A")
    Zc = N.CategoricalMarkovChain(a0, A, states=T, plates=(P,), name="Zc")
    mu = N.GaussianARD(0, 1e-1, plates=(Kc,), name="mu")
    mu.initialize_from_value(rs.randn(Kc))
    tau = N.Gamma(1e-2, 1e-2, plates=(Kc,), name="tau")
    Y = N.Mixture(Zc, N.GaussianARD, mu, tau, name="Y")
    Y.observe(rs.randn(P, T), mask=rs.rand(P, T) < 0.7)
    return [Y, Zc, mu, tau, A, a0]


def gaussians_gated_by_a_markov_chain(N, rs):
    Kc, T = 3, 7
    Zc = N.CategoricalMarkovChain([0.2, 0.3, 0.5], N.Dirichlet(np.ones(Kc), plates=(Kc,), name="A"), states=T, name="Zc")
    B = N.GaussianARD(0, 1.0, shape=(2,), plates=(Kc,), name="B")
    B.initialize_from_value(rs.randn(Kc, 2))
    Y = N.GaussianARD(N.Gate(Zc, B, name="G"), 2.0, ndim=1, name="Y")
    Y.observe(rs.randn(T, 2))
    return [Y, Zc, B]


MODELS = [chain_inputs_time_varying, gaussian_gamma_product, gaussian_gamma_times_constant_and_gaussian,
          gaussian_gamma_taken_by_index, gaussian_gamma_gated, mixture_with_gaussian_gamma_means,
          scalar_gaussian_gamma_over_plates, hidden_markov_chains_with_per_step_transitions,
          switching_state_space_model_selected_by_a_markov_chain, multinomial_counts_over_plates,
          poisson_counts_with_gamma_rates, sum_of_independent_gaussians, concatenated_groups, mixture_over_two_cluster_axes,
          chosen_and_complemented, gamma_shape_point_estimate, dirichlet_concentration_point_estimate,
          masked_hidden_markov_chains_over_plates, gaussians_gated_by_a_markov_chain, bernoulli_mixture, binomial_counts_over_plates]


@pytest.mark.parametrize("model", MODELS, ids=[m.__name__ for m in MODELS])
def test_same_script_same_bounds(both, model):
    bounds = []
    for nodes, inference in both:
        np.random.seed(7)
        Q = inference.VB(*model(nodes, np.random.RandomState(11)))
        Q.update(repeat=4, verbose=False, tol=0)
        bounds.append(np.array(Q.L[:4]))
    assert np.all(np.isfinite(bounds[0]))
    np.testing.assert_allclose(bounds[1], bounds[0], rtol=1e-9)


def test_constant_inputs_equal_inputs_repeated_over_time(both):
    """An input signal with a unit time plate drives every step (the reference's own message code does not take that
    form): same bounds as the reference gets with the rows written out over chains and time."""
    def model(N, tiled):
        rs = np.random.RandomState(5)
        D_, K, T, P = 2, 2, 8, 3
        z = rs.randn(1, K)
        if tiled:
            z = np.broadcast_to(np.repeat(z, T - 1, axis=0), (P, T - 1, K)).copy()
        A = N.GaussianARD(0, 1.0, shape=(D_ + K,), plates=(P, 1, D_), name="A")
        A.initialize_from_value(0.3 * rs.randn(P, 1, D_, D_ + K))
        nu = N.Gamma(2.0, 2.0, plates=(P, 1, D_), name="nu")
        X = N.GaussianMarkovChain(np.zeros(D_), np.identity(D_), A, nu, inputs=z, n=T, name="X")
        Y = N.Gaussian(X, 3.0 * np.identity(D_), name="Y")
        Y.observe(rs.randn(P, T, D_))
        return [X, A, nu, Y]
    (ref_nodes, ref_inf), (our_nodes, our_inf) = both
    Qr = ref_inf.VB(*model(ref_nodes, True))
    Qr.update(repeat=4, verbose=False, tol=0)
    Qo = our_inf.VB(*model(our_nodes, False))
    Qo.update(repeat=4, verbose=False, tol=0)
    np.testing.assert_allclose(Qo.L[:4], Qr.L[:4], rtol=1e-9)


def test_gaussian_rotate_and_vb_conveniences(both, tmp_path, caplog):
    """``Gaussian.rotate`` leaves a consistent q (moments and log-normaliser recomputed from the rotated natural
    parameters agree) and matches the reference's; ``VB.set_autosave`` / ``use_logging`` / ``node.lowerbound``."""
    import logging
    (ref_nodes, ref_inf), (our_nodes, our_inf) = both
    rs = np.random.RandomState(2)
    data, R = rs.randn(7, 3), rs.randn(3, 3)
    states = []
    for N in (ref_nodes, our_nodes):
        X = N.Gaussian(np.ones(3), 2.0 * np.identity(3), plates=(7,), name="X")
        Y = N.Gaussian(X, np.identity(3), name="Y")
        Y.observe(data)
        X.update()
        X.rotate(R)
        states.append(([np.asarray(v) for v in X.u], [np.asarray(v) for v in X.phi], np.asarray(X.g)))
        last = X
    for a, b in zip(states[0][0] + states[0][1], states[1][0] + states[1][1]):
        np.testing.assert_allclose(b, a, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(states[1][2], states[0][2], rtol=1e-10)
    u, g = last._distribution.compute_moments_and_cgf(last.phi)
    np.testing.assert_allclose(np.asarray(u[0]), states[1][0][0], rtol=1e-9)
    np.testing.assert_allclose(np.asarray(g), states[1][2], rtol=1e-9)
    assert last.lowerbound() == float(last.lower_bound_contribution())
    Q = our_inf.VB(Y, last)
    fn = str(tmp_path / "auto.ckpt")
    Q.set_autosave(fn, iterations=1)
    Q.use_logging(True)
    with caplog.at_level(logging.INFO):
        Q.update(repeat=2, tol=0)
    assert any("loglike" in r.getMessage() for r in caplog.records) and any("Auto-saved" in r.getMessage() for r in caplog.records)
    import os
    assert os.path.exists(fn) or os.path.exists(fn + ".npz")


def test_exponential_is_gamma_with_unit_shape(both):
    """``Exponential(l)`` (a stub that raises in the reference, exponential.py:61) is ``Gamma(1, l)``."""
    _, (N, I) = both
    Ls = []
    for make in (lambda b: N.Exponential(b, plates=(6,), name="E"), lambda b: N.Gamma(1, b, plates=(6,), name="E")):
        rs = np.random.RandomState(4)
        b = N.Gamma(2.0, 1.0, name="b")
        E = make(b)
        Y = N.GaussianARD(0, E, plates=(5, 6), name="Y")
        Y.observe(rs.randn(5, 6))
        Q = I.VB(Y, E, b)
        Q.update(repeat=3, verbose=False, tol=0)
        Ls.append(Q.L[:3].copy())
    np.testing.assert_array_equal(Ls[0], Ls[1])


def test_density_queries_match_the_reference(both):
    """``node.logpdf`` / ``pdf`` (expfamily.py:483-504) and ``Mixture.integrated_logpdf_from_parents`` (mixture.py:491-545)."""
    rs = np.random.RandomState(8)
    data, grid = rs.randn(30, 2), rs.randn(11, 2)
    got = []
    for N, I in both:
        np.random.seed(5)
        K = 3
        al = N.Dirichlet(np.ones(K), name="al")
        Z = N.Categorical(al, plates=(30,), name="Z")
        Z.initialize_from_value(np.arange(30) % K)
        mu = N.Gaussian(np.zeros(2), 1e-2 * np.identity(2), plates=(K,), name="mu")
        L = N.Wishart(3.0, np.identity(2), plates=(K,), name="L")
        Y = N.Mixture(Z, N.Gaussian, mu, L, name="Y")
        Y.observe(data)
        Q = I.VB(Y, mu, L, Z, al)
        Q.update(repeat=3, verbose=False, tol=0)
        zh = N.Categorical(al, name="zh")
        Yh = N.Mixture(zh, N.Gaussian, mu, L, name="Yh")
        got.append((np.asarray(Yh.integrated_logpdf_from_parents(grid, 0)),
                    np.asarray(mu.logpdf(np.ones((K, 2)))),
                    np.asarray(mu.pdf(np.zeros((K, 2))))))
    for a, b in zip(got[0], got[1]):
        np.testing.assert_allclose(b, a, rtol=1e-9)
