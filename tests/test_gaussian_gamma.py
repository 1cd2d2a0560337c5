"""``GaussianGamma`` (nodes/gaussian.py:892-1136, :1777-2142), its consumers (SumMultiply with a Gaussian-gamma parent,
GaussianARD / Gaussian with a Gaussian-gamma mean) and the diagonal Wishart made of gamma scalars — against the
reference's own results (tests/golden/gaussian_gamma.npz from make_golden.py: gaussian_gamma_models) and closed forms."""
import numpy as np
import pytest
import scipy.special as sp

from conftest import golden

RTOL = 1e-8


def _check_state(g, prefix, node, rtol=RTOL):
    for i in range(len(node.dims)):
        want = g["%s_u%d" % (prefix, i)]
        np.testing.assert_allclose(np.asarray(node.u[i]) * np.ones(want.shape), want, rtol=rtol, atol=1e-11,
                                   err_msg="%s u%d" % (prefix, i))
        wantp = g["%s_phi%d" % (prefix, i)]
        np.testing.assert_allclose(np.asarray(node.phi[i]) * np.ones(wantp.shape), wantp, rtol=rtol, atol=1e-11,
                                   err_msg="%s phi%d" % (prefix, i))
    wg = g["%s_g" % prefix]
    np.testing.assert_allclose(np.asarray(node.g) * np.ones(wg.shape), wg, rtol=rtol, atol=1e-11, err_msg=prefix + " g")


def test_closed_form_moments_and_shapes(backend):
    """gaussian.py:1051-1083 for fixed parents: tau ~ Gamma(a, b), x | tau ~ N(mu, (tau Lambda)^-1)."""
    from bayespy_b200.nodes import GaussianGamma, Gaussian, Wishart, Gamma
    mu, a, b = np.array([1.0, 2.0, 3.0]), 2.0, 10.0
    X = GaussianGamma(mu, np.identity(3), a, b)
    assert X.plates == () and X.dims == ((3,), (3, 3), (), ())
    u = [np.asarray(v) for v in X.get_moments()]
    np.testing.assert_allclose(u[0], a / b * mu, rtol=1e-13)
    np.testing.assert_allclose(u[1], np.identity(3) + a / b * np.outer(mu, mu), rtol=1e-13)
    np.testing.assert_allclose(u[2], a / b, rtol=1e-13)
    np.testing.assert_allclose(u[3], sp.digamma(a) - np.log(b), rtol=1e-13)
    # plates come from any parent
    assert GaussianGamma(np.ones((4, 3)), np.identity(3), 2, 10).plates == (4,)
    assert GaussianGamma(np.ones(3), np.identity(3), np.ones(4), 10).plates == (4,)
    assert GaussianGamma(np.ones(3), np.identity(3), 2, np.ones(4)).plates == (4,)
    assert GaussianGamma([1, 2], [0.1, 0.2], [0.02, 0.03], [0.03, 0.04], ndim=0).plates == (2,)
    with pytest.raises(ValueError):
        GaussianGamma(np.ones((4, 3)), np.identity(3), 2, 10, plates=(5,))
    with pytest.raises(ValueError):
        GaussianGamma(np.ones(3), np.identity(4), 2, 10)
    # node parents, and a Gaussian-gamma mean
    Xn = GaussianGamma(Gaussian(np.zeros(3), np.identity(3)), Wishart(10, np.identity(3)), 2, Gamma(1, 1))
    assert Xn.dims == ((3,), (3, 3), (), ())
    assert GaussianGamma(GaussianGamma(np.ones(3), np.identity(3), 5, 5), np.identity(3), 5, 5).plates == ()
    # the prior's bound term vanishes (q = p)
    assert abs(float(Xn.lower_bound_contribution())) < 1e-10


def test_factor_model_with_scaled_loadings_matches_the_reference(backend):
    """(a): GaussianGamma -> SumMultiply (Gaussian-gamma output) -> observed, masked GaussianARD."""
    from bayespy_b200.nodes import GaussianGamma, GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
    g = golden("gaussian_gamma")
    M, N, K = 5, 30, 3
    b = Gamma(2.0, 2.0, plates=(M, 1), name="b")
    W = GaussianGamma(np.zeros(K), np.identity(K), 3.0, b, plates=(M, 1), name="W")
    X = GaussianARD(0, 1, shape=(K,), plates=(1, N), name="X")
    X.initialize_from_value(g["a_Xinit"])
    F = SumMultiply("k,k->", W, X, name="F")
    assert F.dims == ((), (), (), ()) and F.moment_kind == "gaussian_gamma"
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(g["a_y"], mask=g["a_mask"])
    Q = VB(Y, W, X, tau, b)
    Q.update(repeat=6, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:6], g["a_L"], rtol=RTOL)
    for nm, nd in (("a_W", W), ("a_X", X), ("a_tau", tau), ("a_b", b)):
        _check_state(g, nm, nd)
    for i, u in enumerate(F.get_moments()):
        want = g["a_F_u%d" % i]
        np.testing.assert_allclose(np.asarray(u) * np.ones(want.shape), want, rtol=RTOL)
    for nm, nd in (("a_lW", W), ("a_lX", X), ("a_ltau", tau), ("a_lb", b), ("a_lY", Y)):
        np.testing.assert_allclose(float(nd.lower_bound_contribution()), g[nm], rtol=1e-7, atol=1e-8)


def test_gaussian_gamma_mean_of_a_gaussian_matches_the_reference(backend):
    """(b) unknown mean, precision and rate above the Gaussian-gamma, a Wishart precision next to it; then (e) rotate and
    translate of the posterior."""
    from bayespy_b200.nodes import GaussianGamma, Gaussian, Wishart, Gamma
    from bayespy_b200.inference import VB
    g = golden("gaussian_gamma")
    z = g["b_z"]
    Nb, Dm = z.shape
    mu0 = Gaussian(np.zeros(Dm), 1e-2 * np.identity(Dm), name="mu0")
    L0 = Wishart(Dm + 1.0, np.identity(Dm), name="L0")
    bb = Gamma(1.5, 1.0, name="bb")
    m = GaussianGamma(mu0, L0, 2.5, bb, name="m")
    L1 = Wishart(Dm + 2.0, np.identity(Dm), name="L1")
    Z = Gaussian(m, L1, plates=(Nb,), name="Z")
    Z.observe(z)
    Q = VB(Z, m, L1, mu0, L0, bb)
    Q.update(repeat=5, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:5], g["b_L"], rtol=RTOL)
    for nm, nd in (("b_m", m), ("b_L1", L1), ("b_mu0", mu0), ("b_L0", L0), ("b_bb", bb)):
        _check_state(g, nm, nd)
    L_before = Q.compute_lowerbound()
    m.rotate(g["e_R"])
    _check_state(g, "e_rot", m)
    m.translate(g["e_b"])
    _check_state(g, "e_tra", m)
    np.testing.assert_allclose(m.get_gaussian_location(), g["e_loc"], rtol=RTOL)
    # the Student-t marginal: mean = location, variance = nu / (nu - 2) * b / a * diag(Cov)
    mean, var = m.get_gaussian_mean_and_variance()
    a = float(np.asarray(m.phi[3]))
    tau = float(np.asarray(m.u[2]))
    Cov = np.linalg.inv(-2 * np.asarray(m.phi[1]))
    np.testing.assert_allclose(mean, g["e_loc"], rtol=RTOL)
    np.testing.assert_allclose(var, 2 * a / (2 * a - 2) * np.diag(Cov) / tau, rtol=1e-7)
    assert np.isfinite(L_before)


def test_scalar_gaussian_gamma_as_the_mean_of_a_gaussian_ard(backend):
    """(c): ndim=0 with a gamma-like precision; GaussianARD's joint message is split between the scaled mean and alpha."""
    from bayespy_b200.nodes import GaussianGamma, GaussianARD, Gamma
    from bayespy_b200.inference import VB
    g = golden("gaussian_gamma")
    P = 4
    lam = Gamma(2.0, 1.0, plates=(P,), name="lam")
    bc = Gamma(1.0, 1.0, plates=(P,), name="bc")
    mc = GaussianGamma(np.zeros(P), lam.as_wishart(ndim=0), 1.5 * np.ones(P), bc, ndim=0, name="mc")
    assert mc.plates == (P,) and mc.dims == ((), (), (), ())
    al = Gamma(1e-2, 1e-2, plates=(6, 1), name="al")
    Yc = GaussianARD(mc, al, name="Yc")
    assert Yc.plates == (6, P)
    Yc.observe(g["c_y"])
    Q = VB(Yc, mc, lam, bc, al)
    Q.update(repeat=5, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:5], g["c_L"], rtol=RTOL)
    for nm, nd in (("c_mc", mc), ("c_lam", lam), ("c_bc", bc), ("c_al", al)):
        _check_state(g, nm, nd)
    with pytest.raises(NotImplementedError):
        GaussianARD(GaussianGamma(np.ones(3), np.identity(3), 5, 5), 1.0)      # vector Gaussian-gamma as a scalar mean


def test_diagonal_wishart_of_gamma_scalars(backend):
    """(d): gamma.py:337-397."""
    from bayespy_b200.nodes import Gaussian, Gamma
    from bayespy_b200.inference import VB
    g = golden("gaussian_gamma")
    z = g["b_z"]
    Nb, Dm = z.shape
    gam = Gamma(1e-2, 1e-2, plates=(Dm,), name="g")
    Wd = gam.diag()
    assert Wd.plates == () and Wd.dims == ((Dm, Dm), ())
    Zd = Gaussian(np.zeros(Dm), Wd, plates=(Nb,), name="Zd")
    Zd.observe(z)
    Q = VB(Zd, gam)
    Q.update(repeat=2, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:2], g["d_L"], rtol=RTOL)
    _check_state(g, "d_g", gam)
    for i, u in enumerate(Wd.get_moments()):
        np.testing.assert_allclose(np.asarray(u), g["d_W_u%d" % i], rtol=RTOL)
    with pytest.raises(Exception):
        Gamma(1, 1).diag()
