"""Rotation parameter expansion against the reference's own run of the pca.rst example
with ``Q.set_callback(R.rotate)``: cost function and gradient of both blocks at a fixed R, the lower-bound
trajectory with the callback, and the rotated posteriors."""
import numpy as np
import pytest

from conftest import golden


def _model(g, M, N, Dm):
    from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
    X = GaussianARD(0, 1, shape=(Dm,), plates=(1, N), name="X")
    alpha = Gamma(1e-5, 1e-5, plates=(Dm,), name="alpha")
    C = GaussianARD(0, alpha, shape=(Dm,), plates=(M, 1), name="C")
    F = SumMultiply("d,d->", X, C)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(g["y"])
    C.initialize_from_value(g["C_init"])
    Q = VB(Y, X, C, alpha, tau)
    return Q, dict(X=X, C=C, alpha=alpha, tau=tau)


@pytest.mark.parametrize("fused", [True, False])
def test_pca_with_rotations_matches_reference(backend, fused):
    from bayespy_b200.inference.vmp.transformations import RotateGaussianARD, RotationOptimizer
    g = golden("pca_rotated")
    M, N, Dm = 10, 60, 4
    Q, n = _model(g, M, N, Dm)
    if not fused:
        Q.plans = []
        for node in (n["X"], n["C"]):
            pass
    rot_X = RotateGaussianARD(n["X"])
    rot_C = RotateGaussianARD(n["C"], n["alpha"])
    R = RotationOptimizer(rot_X, rot_C, Dm)
    Q.update(repeat=2, verbose=False, tol=0)
    rot_X.setup()
    rot_C.setup()
    bX, dbX = rot_X.bound(g["Rtest"])
    bC, dbC = rot_C.bound(np.linalg.inv(g["Rtest"]).T)
    np.testing.assert_allclose(bX, g["bX"], rtol=1e-9)
    np.testing.assert_allclose(dbX, g["dbX"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(bC, g["bC"], rtol=1e-9)
    np.testing.assert_allclose(dbC, g["dbC"], rtol=1e-8, atol=1e-8)
    Q.set_callback(R.rotate)
    iters = len(g["L"]) - 2
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters + 2], g["L"], rtol=1e-6)
    for nm, node in n.items():
        for i in range(2):
            np.testing.assert_allclose(np.asarray(node.u[i]), g["%s_u%d" % (nm, i)], rtol=1e-4, atol=1e-6,
                                       err_msg="%s.u[%d]" % (nm, i))


def test_rotation_does_not_change_the_bound_at_identity_and_improves_it(backend):
    from bayespy_b200.inference.vmp.transformations import RotateGaussianARD, RotationOptimizer
    g = golden("pca_rotated")
    Q, n = _model(g, 10, 60, 4)
    Q.update(repeat=3, verbose=False, tol=0)
    L0 = Q.compute_lowerbound()
    R = RotationOptimizer(RotateGaussianARD(n["X"]), RotateGaussianARD(n["C"], n["alpha"]), 4)
    R.rotate(maxiter=0)                      # R = I: moments, phi and g must be unchanged
    np.testing.assert_allclose(Q.compute_lowerbound(), L0, rtol=1e-10)
    R.rotate(check_bound=True)
    assert Q.compute_lowerbound() >= L0 - 1e-8 * abs(L0)
