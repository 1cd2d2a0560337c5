"""``node[...]``: basic slicing of plates (node.py:761-763, Slice :868-1160) — strided views of the parent's moments,
messages placed back through the same view — against the unmodified reference (tests/golden/slice.npz) and against
NumPy's own indexing."""
import numpy as np
import pytest

from conftest import golden


def test_sliced_parents_match_reference(backend):
    from bayespy_b200.nodes import GaussianARD
    from bayespy_b200.inference import VB
    g = golden("slice")
    mu = GaussianARD(0.5, 1e-2, plates=(6, 4), name="mu")
    vec = GaussianARD(0, 1e-1, shape=(2,), plates=(5,), name="vec")
    children = [("a", mu[1:4, ::2], 2.0), ("b", mu[0], 1.5), ("c", mu[None, 5, 1:3], 3.0), ("d", mu[..., -1], 1.0),
                ("e", mu[::2, 1], 0.7)]
    nodes = []
    for nm, parent, prec in children:
        Y = GaussianARD(parent, prec, name="y_" + nm)
        assert tuple(Y.plates) == tuple(g["plates_" + nm])
        Y.observe(g["y_" + nm])
        nodes.append(Y)
    Yv = GaussianARD(vec[1:4], [2.0, 0.5], shape=(2,), name="y_v")
    Yv.observe(g["y_v"])
    Q = VB(mu, vec, Yv, *nodes)
    Q.update(repeat=2, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:2], g["L"], rtol=1e-10)
    for nm, node in (("mu", mu), ("vec", vec)):
        for i in range(2):
            np.testing.assert_allclose(np.asarray(node.u[i]), g["%s_u%d" % (nm, i)], rtol=1e-10, atol=1e-12,
                                       err_msg="%s.u[%d]" % (nm, i))
            ref = g["%s_phi%d" % (nm, i)]
            np.testing.assert_allclose(np.broadcast_to(np.asarray(node.phi[i]), ref.shape), ref, rtol=1e-10, atol=1e-12,
                                       err_msg="%s.phi[%d]" % (nm, i))


@pytest.mark.parametrize("index", [0, (slice(None), 2), (slice(1, 5, 2), slice(None, None, -1)), (None, 3), (Ellipsis, 1),
                                   (2, None, slice(0, 3)), (slice(None), slice(None)), -1, (Ellipsis,),
                                   (slice(4, None, -3), Ellipsis)])
def test_slice_moments_and_plates_follow_numpy(backend, index):
    """Moments of ``X[index]`` are NumPy's ``u[index]`` bit for bit (pure data movement), plates included; the doc's
    examples ``y[0]``, ``y[:, ::2]``, ``y[:5, 10:20:5]`` (modelconstruct.rst:205-220) are of this kind."""
    from bayespy_b200.nodes import GaussianARD
    rs = np.random.RandomState(1)
    X = GaussianARD(0, 1, shape=(3,), plates=(6, 4), name="X")
    X.initialize_from_parameters(rs.randn(6, 4, 3), 1 + rs.rand(6, 4, 3))
    S = X[index]
    u = [np.asarray(a) for a in X.u]
    idx = index if isinstance(index, tuple) else (index,)
    for i, ui in enumerate(S.get_moments()):
        flat = np.asarray(u[i]).reshape(6, 4, -1)                 # the index applies to the PLATES only
        ref = np.stack([flat[..., k][idx] for k in range(flat.shape[-1])], axis=-1)
        assert tuple(S.plates) == ref.shape[:-1]
        np.testing.assert_array_equal(np.asarray(ui).reshape(ref.shape), ref)


def test_slice_errors_like_the_reference(backend):
    from bayespy_b200.nodes import GaussianARD
    X = GaussianARD(0, 1, plates=(6, 4))
    with pytest.raises(IndexError):
        X[0, 1, 2]
    with pytest.raises(IndexError):
        X[7]
    with pytest.raises(IndexError):
        X[3:3]
    with pytest.raises(TypeError):
        X[[0, 2]]
    assert X[:5, 1:4:2].plates == (5, 2)
