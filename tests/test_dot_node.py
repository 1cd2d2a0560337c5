"""SumMultiply node (dot.py:19-633) against the reference for key patterns beyond the PCA inner product:
matrix-vector, outer product, three operands, kept axes — moments <f>, <f f> and the messages to every parent
(the cases of nodes/tests/test_dot.py, with the reference's own outputs as goldens)."""
import numpy as np
import pytest

from conftest import golden

CASES = {
    "dot": ("i,i", [((4,), (3, 1)), ((4,), (1, 5))]),
    "matvec": ("ij,j->i", [((3, 4), (2,)), ((4,), (2,))]),
    "outer": ("i,j->ij", [((3,), (5,)), ((2,), (1,))]),
    "triple": ("i,i,i->", [((3,), (2, 1)), ((3,), (1, 4)), ((3,), ())]),
    "keepdim": ("ij,ik->jk", [((2, 3), ()), ((2, 4), (3,))]),
}


@pytest.mark.parametrize("tag", sorted(CASES))
def test_summultiply_moments_and_messages(backend, tag):
    from bayespy_b200.nodes import GaussianARD, SumMultiply
    g = golden("summultiply_nodes")
    spec, ins = CASES[tag]
    nodes = []
    for i, (shape, plates) in enumerate(ins):
        x = GaussianARD(0, 1, shape=shape, plates=plates, name="x%d" % i)
        nd = len(shape)
        u0, u1 = g["%s_in%d_mu" % (tag, i)], g["%s_in%d_u1" % (tag, i)]
        cov = u1 - u0.reshape(u0.shape + (1,) * nd) * u0.reshape(u0.shape[:u0.ndim - nd] + (1,) * nd + u0.shape[u0.ndim - nd:])
        x.initialize_from_mean_and_covariance(u0, cov)
        nodes.append(x)
    F = SumMultiply(spec, *nodes)
    assert tuple(F.plates) == tuple(int(v) for v in g[tag + "_plates"])
    u = F.get_moments()
    np.testing.assert_allclose(np.asarray(u[0]), g[tag + "_u0"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.asarray(u[1]), g[tag + "_u1"], rtol=1e-10, atol=1e-12)
    Y = GaussianARD(F, float(g[tag + "_tau"]), ndim=len(F.dims[0]))
    Y.observe(g[tag + "_y"])
    for i in range(len(nodes)):
        m = F.message_to_parent(i)
        for j in range(2):
            ref = g["%s_m%d_%d" % (tag, i, j)]
            got = np.asarray(m[j].materialize() if hasattr(m[j], "materialize") else m[j])
            while got.ndim > ref.ndim and got.shape[0] == 1:      # extra unit plate axes are only a convention
                got = got[0]
            np.testing.assert_allclose(np.broadcast_to(got, ref.shape), ref, rtol=1e-10, atol=1e-11,
                                       err_msg="%s: message %d to parent %d" % (tag, j, i))
