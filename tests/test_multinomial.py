"""``Multinomial`` (nodes/multinomial.py:60-319) and mixtures of multinomials — against the reference's own results
(tests/golden/multinomial.npz from make_golden.py: multinomial_models) and closed forms."""
import numpy as np
import pytest

from conftest import golden


def _same(node, g, prefix, rtol=1e-9):
    for i in range(len(node.u)):
        ref = g["%s_u%d" % (prefix, i)]
        a, ref = np.broadcast_arrays(np.asarray(node.u[i]), ref)
        np.testing.assert_allclose(a, ref, rtol=rtol, atol=1e-12, err_msg="%s.u[%d]" % (prefix, i))
        ref = g["%s_phi%d" % (prefix, i)]
        a, ref = np.broadcast_arrays(np.asarray(node.phi[i]), ref)
        np.testing.assert_allclose(a, ref, rtol=rtol, atol=1e-12, err_msg="%s.phi[%d]" % (prefix, i))
    ref = g[prefix + "_g"]
    a, ref = np.broadcast_arrays(np.asarray(node.g), ref)
    np.testing.assert_allclose(a, ref, rtol=rtol, atol=1e-12, err_msg=prefix + ".g")


def test_counts_with_a_dirichlet_prior(backend):
    from bayespy_b200.nodes import Multinomial, Dirichlet
    from bayespy_b200.inference import VB
    g = golden("multinomial")
    prior = np.array([1.0, 2.0, 0.5, 1.5])
    p = Dirichlet(prior, name="p")
    X = Multinomial(12, p, plates=(6,), name="X")
    X.observe(g["a_x"])
    Q = VB(X, p)
    Q.update(repeat=2, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:2], g["a_L"], rtol=1e-10)
    _same(p, g, "a_p")
    np.testing.assert_allclose(np.asarray(p.phi[0]), prior + g["a_x"].sum(axis=0), rtol=1e-13)     # conjugate update
    for bad in (g["a_x"] + 1, -g["a_x"], g["a_x"].astype(float)):
        with pytest.raises(ValueError):
            Multinomial(12, p, plates=(6,)).observe(bad)
    with pytest.raises(ValueError):
        Multinomial(2.5, [0.5, 0.5])
    with pytest.raises(ValueError):
        Multinomial(-1, [0.5, 0.5])


def test_mixture_of_multinomials_matches_reference(backend):
    from bayespy_b200.nodes import Multinomial, Dirichlet, Categorical, Mixture
    from bayespy_b200.inference import VB
    g = golden("multinomial")
    counts, n = g["b_counts"], g["b_n"]
    N, K = counts.shape
    C = 3
    alpha = Dirichlet(np.ones(C), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="Z")
    Z.initialize_from_value(g["b_zinit"])
    P = Dirichlet(np.ones(K), plates=(C,), name="P")
    Xm = Mixture(Z, Multinomial, n, P, name="Xm")
    assert tuple(Xm.plates) == (N,)
    Xm.observe(counts)
    Q = VB(Xm, P, Z, alpha)
    Q.update(repeat=6, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:6], g["b_L"], rtol=1e-9)
    for nm, nd in (("b_P", P), ("b_Z", Z), ("b_alpha", alpha)):
        _same(nd, g, nm, rtol=1e-8)
    with pytest.raises(ValueError):
        Mixture(Z, Multinomial, np.ones((N, C), dtype=int), P)       # trials must not vary over the cluster axis


def test_unobserved_multinomial_moments(backend):
    from bayespy_b200.nodes import Multinomial
    g = golden("multinomial")
    Xf = Multinomial(np.array([[3], [7]]), np.array([[0.2, 0.8], [0.5, 0.5], [0.9, 0.1]]), name="Xf")
    assert tuple(Xf.plates) == (2, 3)
    _same(Xf, g, "c_Xf", rtol=1e-12)
    np.random.seed(4)
    x = Xf.random()
    assert x.shape == (2, 3, 2) and np.array_equal(x.sum(axis=-1), np.broadcast_to([[3], [7]], (2, 3)))
