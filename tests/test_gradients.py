"""Gradient-based learning: natural / Euclidean gradients of the nodes' natural parameters,
``VB.gradient_step``, ``VB.optimize`` (Riemannian conjugate gradient with collapsed nodes, plain CG, gradient ascent) and
``VB.pattern_search`` — vmp.py:402-662, expfamily.py:260-340, gaussian.py:489-555 / :824-890, gamma.py:183-211 — against
the unmodified reference (tests/golden/pca_gradients.npz) and against finite differences of the bound."""
import numpy as np
import pytest

from conftest import golden


def _pca(g):
    from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
    y = g["y"]
    M, N = y.shape
    Dm = g["C_init"].shape[-1]
    X = GaussianARD(0, 1, shape=(Dm,), plates=(1, N), name="X")
    alpha = Gamma(1e-3, 1e-3, plates=(Dm,), name="alpha")
    C = GaussianARD(0, alpha, shape=(Dm,), plates=(M, 1), name="C")
    F = SumMultiply("d,d->", X, C)
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    C.initialize_from_value(g["C_init"])
    Q = VB(Y, X, C, alpha, tau)
    return Q, dict(X=X, C=C, alpha=alpha, tau=tau)


def _close(a, ref, rtol, name, scale=0.0):
    ref = np.asarray(ref)
    np.testing.assert_allclose(np.broadcast_to(np.asarray(a), ref.shape), ref, rtol=rtol,
                               atol=rtol * max(np.max(np.abs(ref)), scale, 1e-300), err_msg=name)


def test_gradients_steps_and_optimizers_match_reference(backend):
    g = golden("pca_gradients")
    Q, n = _pca(g)
    Q.update(repeat=2, verbose=False, tol=0)
    rg, gr = Q.get_gradients(n["C"], n["tau"], n["X"], euclidian=True)
    for nm, r_, g_ in zip(("C", "tau", "X"), rg, gr):
        for i in range(2):
            # a node that has just been updated has a zero natural gradient: differences of two equal parameters
            scale = float(np.max(np.abs(np.asarray(n[nm].phi[i]))))
            _close(r_[i], g["rg_%s_%d" % (nm, i)], 1e-8, "Riemannian gradient %s[%d]" % (nm, i), scale)
            _close(g_[i], g["g_%s_%d" % (nm, i)], 1e-8, "gradient %s[%d]" % (nm, i), scale)
    np.testing.assert_allclose(Q.dot(rg, gr), float(g["dot"]), rtol=1e-9)
    Q.gradient_step(n["C"], n["tau"], scale=0.4)
    np.testing.assert_allclose(Q.compute_lowerbound(), float(g["L_after_step"]), rtol=1e-9)
    for nm in ("C", "tau"):
        for i in range(2):
            _close(n[nm].u[i], g["step_%s_u%d" % (nm, i)], 1e-8, "%s.u[%d] after the step" % (nm, i))
            _close(n[nm].phi[i], g["step_%s_phi%d" % (nm, i)], 1e-8, "%s.phi[%d] after the step" % (nm, i))
    Q.optimize(n["C"], n["tau"], maxiter=6, collapsed=[n["X"], n["alpha"]], verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:Q.iter], g["L_opt1"], rtol=1e-7)
    Q.optimize(n["C"], n["X"], maxiter=4, riemannian=False, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:Q.iter], g["L_opt2"], rtol=1e-7)
    Q.optimize(n["C"], n["tau"], maxiter=3, method="gradient", verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:Q.iter], g["L_opt3"], rtol=1e-7)
    Q.pattern_search(n["C"], n["tau"], collapsed=[n["X"], n["alpha"]])
    Q.pattern_search(n["C"], n["X"])
    np.testing.assert_allclose(Q.L[:Q.iter], g["L"], rtol=1e-6)
    for nm in ("X", "C", "alpha", "tau"):
        for i in range(2):
            _close(n[nm].u[i], g["%s_u%d" % (nm, i)], 1e-5, "%s.u[%d] at the end" % (nm, i))
    with pytest.raises(Exception):
        Q.optimize(n["C"], method="newton")


def test_gradient_is_the_derivative_of_the_bound_also_with_annealing(backend):
    """vmp/tests/test_annealing.py:36-96 restated: d bound / d phi_i by central differences equals ``get_gradient`` of
    the Riemannian gradient, for a scalar GaussianARD, a vector GaussianARD and a Gamma node, at annealing 1 and 0.1."""
    from bayespy_b200.nodes import GaussianARD, Gamma
    from bayespy_b200.inference import VB
    rs = np.random.RandomState(4)
    for annealing in (1.0, 0.1):
        X = GaussianARD(3, 4, name="x")
        X.initialize_from_parameters(-1, 6)
        V = GaussianARD(rs.randn(3), [1.0, 2.0, 0.5], shape=(3,), plates=(2,), name="v")
        V.initialize_from_parameters(rs.randn(2, 3), 3 + rs.rand(2, 3))
        T = Gamma(2.0, 3.0, plates=(2,), name="t")
        T.initialize_from_value(np.array([0.7, 1.9]))
        T.update()
        T.set_parameters([-np.array([2.5, 1.5]), np.array([3.0, 4.5])])
        W = GaussianARD(0.5, T, plates=(2,), name="w")
        W.observe(np.array([0.3, -1.2]))
        Q = VB(X, V, T, W)
        Q.set_annealing(annealing)
        for node in (X, V, T):
            rg = node.get_riemannian_gradient()
            gr = [np.asarray(a) for a in node.get_gradient(rg)]
            p0 = [np.asarray(a).copy() for a in node.get_parameters()]
            for i in range(len(p0)):
                flat = p0[i].reshape(-1)
                for j in rs.choice(flat.size, size=min(3, flat.size), replace=False):
                    eps = 1e-6 * max(1.0, abs(flat[j]))
                    vals = []
                    for sgn in (+1, -1):
                        p = [a.copy() for a in p0]
                        p[i].reshape(-1)[j] += sgn * eps
                        if node is not T and i == 1 and p[i].ndim >= 2:
                            # keep the precision parameter symmetric: perturb the mirrored entry too
                            K = p[i].shape[-1]
                            r, c = (j // K) % K, j % K
                            if r != c:
                                p[i][..., c, r].reshape(-1)[j // (K * K)] += sgn * eps
                        node.set_parameters(p)
                        vals.append(Q.compute_lowerbound(ignore_masked=False))
                    node.set_parameters(p0)
                    num = (vals[0] - vals[1]) / (2 * eps)
                    ana = gr[i].reshape(-1)[j]
                    if node is not T and i == 1 and p0[i].ndim >= 2:
                        K = p0[i].shape[-1]
                        r, c = (j // K) % K, j % K
                        if r != c:
                            ana = ana + gr[i][..., c, r].reshape(-1)[j // (K * K)]
                    np.testing.assert_allclose(ana, num, rtol=2e-4, atol=1e-6,
                                               err_msg="%s phi[%d][%d], annealing %g" % (node.name, i, j, annealing))


def test_stochastic_variational_inference_with_plate_multipliers(backend):
    """demos/stochastic_inference.py:93-141 in small: the mini-batch model's Categorical stands for N / N_batch times its
    plates (``plates_multiplier``); every step observes a seeded mini-batch, updates the local Z and takes a
    natural-gradient step of the global nodes.  Means, class weights and bound after every step against the reference."""
    from bayespy_b200.nodes import Gaussian, Dirichlet, Categorical, Mixture
    from bayespy_b200.inference import VB
    g = golden("svi_mixture")
    data, subsets = g["data"], g["subsets"]
    N, Dm = data.shape
    N_batch, K = subsets.shape[1], g["mu_init"].shape[0]
    mu = Gaussian(np.zeros(Dm), np.identity(Dm), plates=(K,), name="means")
    alpha = Dirichlet(np.ones(K), name="class probabilities")
    Z = Categorical(alpha, plates=(N_batch,), plates_multiplier=(N / N_batch,), name="classes")
    Y = Mixture(Z, Gaussian, mu, np.identity(Dm), name="observations")
    assert tuple(Y.plates_multiplier) == (N / N_batch,) and tuple(mu.plates_multiplier) == ()
    mu.initialize_from_value(g["mu_init"])
    Q = VB(Y, Z, mu, alpha)
    Q.ignore_bound_checks = True
    for n in range(len(subsets)):
        Y.observe(data[subsets[n], :])
        Q.update(Z, verbose=False)
        Q.gradient_step(mu, alpha, scale=(n + 1) ** (-0.7))
        np.testing.assert_allclose(np.asarray(mu.u[0]), g["mus"][n], rtol=1e-7, atol=1e-9, err_msg="means, step %d" % n)
        np.testing.assert_allclose(np.asarray(alpha.u[0]), g["alphas"][n], rtol=1e-7, atol=1e-9,
                                   err_msg="class weights, step %d" % n)
        np.testing.assert_allclose(Q.compute_lowerbound(), g["Ls"][n], rtol=1e-8, err_msg="bound, step %d" % n)
    np.testing.assert_allclose(Q.L[:Q.iter], g["L"], rtol=1e-8)
    # incompatible multipliers are refused
    with pytest.raises(ValueError):
        Mixture(Categorical(alpha, plates=(N_batch,), plates_multiplier=(3.0,)), Gaussian,
                Gaussian(np.zeros(Dm), np.identity(Dm), plates=(N_batch, K), plates_multiplier=(2.0, 1)), np.identity(Dm))


def _guide_pca(data, nx, mult=None):
    from bayespy_b200.nodes import GaussianARD, Gamma, Dot
    Dm = 3
    kw = {} if mult is None else dict(plates_multiplier=mult)
    X = GaussianARD(0, 1, shape=(Dm,), plates=(1, nx), name="X", **kw)
    alpha = Gamma(1e-3, 1e-3, plates=(Dm,), name="alpha")
    C = GaussianARD(0, alpha, shape=(Dm,), plates=(10, 1), name="C")
    F = Dot(C, X)
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    return X, alpha, C, tau, Y


def test_advanced_guide_annealing_and_stochastic_pca(backend):
    """doc/source/user_guide/advanced.rst: deterministic annealing (:219-224: the annealing is raised by 1.5x after every
    convergence at tol=1e-4; same stopping iterations and bounds as the reference) and stochastic VI on the PCA model
    with ``plates_multiplier=(1, 20)`` on X, inherited by Dot and Y (:276-313: C, alpha, tau after every step)."""
    from bayespy_b200.inference import VB
    g = golden("advanced_guide")
    data = g["data"]
    X, alpha, C, tau, Y = _guide_pca(data, 100)
    Y.observe(data)
    Q = VB(Y, C, X, alpha, tau)
    X.initialize_from_parameters(g["X_init"], 10)
    beta, sched = 0.1, []
    while beta < 1.0:
        beta = min(beta * 1.5, 1.0)
        Q.set_annealing(beta)
        Q.update(repeat=100, tol=1e-4, verbose=False)
        sched.append((beta, Q.iter))
    np.testing.assert_allclose(np.array(sched), g["anneal_schedule"], rtol=1e-12)
    np.testing.assert_allclose(Q.L[:Q.iter], g["anneal_L"], rtol=1e-7)
    for nm, node in (("C", C), ("tau", tau), ("alpha", alpha)):
        for i in range(2):
            np.testing.assert_allclose(np.asarray(node.u[i]), g["anneal_%s_u%d" % (nm, i)], rtol=1e-6, atol=1e-9,
                                       err_msg="%s.u[%d] after annealing" % (nm, i))
    # stochastic variational inference
    X, alpha, C, tau, Y = _guide_pca(data, 5, mult=(1, 20))
    assert tuple(Y.plates_multiplier) == (20,) or tuple(Y.plates_multiplier) == (1, 20)
    Q = VB(Y, C, X, alpha, tau)
    C.initialize_from_value(g["C_init"])
    Q.ignore_bound_checks = True
    for n, subset in enumerate(g["subsets"]):
        Y.observe(data[:, subset])
        Q.update(X, verbose=False)
        Q.gradient_step(C, alpha, tau, scale=(n + 2.0) ** (-0.7))
        np.testing.assert_allclose(np.asarray(C.u[0]), g["svi_C"][n], rtol=1e-7, atol=1e-9, err_msg="C, step %d" % n)
        np.testing.assert_allclose(np.asarray(tau.u[0]), g["svi_tau"][n], rtol=1e-7, err_msg="tau, step %d" % n)
        np.testing.assert_allclose(np.asarray(alpha.u[0]), g["svi_alpha"][n], rtol=1e-7, err_msg="alpha, step %d" % n)
    np.testing.assert_allclose(Q.L[:Q.iter], g["svi_L"], rtol=1e-8)
