"""``Take`` (nodes/take.py): index gathers along a plate axis, bit-exact moments, scatter-added messages — against
the reference's own results (tests/golden/take.npz from make_golden.py: take_models) and its argument checks."""
import numpy as np
import pytest

from conftest import golden


def test_take_docstring_example_and_plates(backend):
    from bayespy_b200.nodes import Gamma, Take, GaussianARD
    alpha = Gamma([1, 2, 3], [1, 1, 1])
    x = Take(alpha, [1, 1, 2, 2, 1, 0])
    assert np.array_equal(np.asarray(x.get_moments()[0]), [2., 2., 3., 3., 2., 1.])      # take.py:37-41, bit-exact
    X = GaussianARD(1, 1, plates=(2,), shape=())
    assert Take(X, 1).plates == () and Take(X, [1, 1, 0, 1]).plates == (4,)
    assert Take(X, [[1, 1, 0], [1, 0, 1]]).plates == (2, 3)
    X2 = GaussianARD(1, 1, plates=(3, 2), shape=(4,))
    assert Take(X2, [1, 1, 0], plate_axis=-2).plates == (3, 2)
    assert tuple(Take(X2, [0, 1, 1, 0]).dims) == ((4,), (4, 4))
    for bad in (dict(plate_axis=0), dict(plate_axis=-3), dict(indices=[0.5]), dict(indices=[2]), dict(indices=[-3])):
        kw = dict(indices=[0], plate_axis=-1)
        kw.update(bad)
        with pytest.raises(ValueError):
            Take(X, kw["indices"], plate_axis=kw["plate_axis"])


def test_take_model_matches_the_reference(backend):
    from bayespy_b200.nodes import GaussianARD, Gamma, Take
    from bayespy_b200.inference import VB
    g = golden("take")
    G = 3
    mu = GaussianARD(0, 1e-3, plates=(G,), name="mu")
    m = Take(mu, g["idx"], name="m")
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(m, tau, name="Y")
    Y.observe(g["y"], mask=g["mask"])
    Q = VB(mu, tau, Y)
    Q.update(repeat=5, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:5], g["L"], rtol=1e-9)
    for i in range(2):
        np.testing.assert_allclose(np.asarray(mu.u[i]), g["mu_u%d" % i], rtol=1e-9)
        np.testing.assert_allclose(np.asarray(tau.u[i]), g["tau_u%d" % i], rtol=1e-9)
        np.testing.assert_allclose(np.asarray(m.get_moments()[i]), g["m_u%d" % i], rtol=1e-9)
    # the gather itself moves bits: the taken moments equal the parent's rows exactly
    idx = np.where(g["idx"] < 0, g["idx"] + G, g["idx"])
    assert np.array_equal(np.asarray(m.get_moments()[0]), np.asarray(mu.u[0])[idx])


def test_take_along_an_inner_plate_axis_with_a_matrix_of_indices(backend):
    from bayespy_b200.nodes import GaussianARD, Take
    from bayespy_b200.inference import VB
    g = golden("take")
    rs = np.random.RandomState(31)
    # replay the generator's stream up to the second model (same draws as make_golden.take_models)
    N, G = 40, 3
    rs.randint(0, G, size=N); rs.randn(N); rs.rand(N)
    X = GaussianARD(0, 1, plates=(3, 4), shape=(2,), name="X")
    X.initialize_from_value(rs.randn(3, 4, 2))
    Z = Take(X, g["idx2"], plate_axis=-2, name="Z")
    assert Z.plates == (2, 2, 4)
    W = GaussianARD(Z, 2.0, name="W")
    W.observe(g["y2"])
    Q = VB(X, W)
    Q.update(repeat=2, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:2], g["L2"], rtol=1e-9)
    np.testing.assert_allclose(np.asarray(Z.get_moments()[0]), g["Z_u0"], rtol=1e-9)
    for i in range(2):
        np.testing.assert_allclose(np.asarray(X.u[i]), g["X_u%d" % i], rtol=1e-9)
