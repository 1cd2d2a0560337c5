"""Pin the model-level oracle (oracle/pca_ref.py) against the reference's golden vectors."""
import numpy as np
import pytest

from conftest import golden


@pytest.mark.parametrize("name,K", [("pca_small", 5), ("pca_64x16", 16), ("pca_masked", 4)])
def test_pca_oracle_matches_reference(name, K):
    from oracle.pca_ref import PcaOracle
    g = golden(name)
    mask = g["mask"] if "mask" in g.files else None
    o = PcaOracle(g["y"], K, g["C_init"], mask=mask)
    iters = len(g["L"])
    for _ in range(iters):
        o.sweep()
    np.testing.assert_allclose(o.L, g["L"], rtol=1e-9)
    for nm in ("Y", "X", "C", "alpha", "tau"):
        np.testing.assert_allclose([t[nm] for t in o.l], g["l_" + nm], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(o.X_u0, g["X_u0"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(o.X_u1, g["X_u1"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(o.C_u0, g["C_u0"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(o.alpha_u0, g["alpha_u0"], rtol=1e-8)
    np.testing.assert_allclose(o.tau_u0, g["tau_u0"], rtol=1e-8)
    np.testing.assert_allclose(o.X_phi[1], g["X_phi1"], rtol=1e-8, atol=1e-10)


def test_make_data_matches_golden():
    from oracle.pca_ref import make_data
    g = golden("pca_small")
    assert np.array_equal(make_data(20, 100), g["y"])
