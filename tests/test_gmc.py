"""Gaussian Markov chain: bpk_block_banded_solve against the reference's
utils.linalg.block_banded_solve, and a small linear state-space model (lssm.rst:45-181 scaled down)
against the reference's own VB run — lower-bound trajectory and every node's moments."""
import numpy as np
import pytest

from conftest import golden
from bayespy_b200.darray import DArray


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_block_banded_solve_golden(backend, tag):
    g = golden("block_banded")
    A, B, y = g["A_" + tag], g["B_" + tag], g["y_" + tag]
    T, Dm = y.shape
    Ad, yd = DArray.from_numpy(A), DArray.from_numpy(y)
    Bd = DArray.from_numpy(B) if T > 1 else DArray.empty((1,))
    V, C, x, ld = DArray.empty((T, Dm, Dm)), DArray.empty((max(T - 1, 1), Dm, Dm)), DArray.empty((T, Dm)), DArray.empty(())
    backend.block_banded_solve(Ad.ptr, Bd.ptr, yd.ptr, 1, T, Dm, V.ptr, C.ptr, x.ptr, ld.ptr, True)
    np.testing.assert_allclose(V.numpy(), g["V_" + tag], rtol=1e-9, atol=1e-12)
    if T > 1:
        np.testing.assert_allclose(C.numpy()[:T - 1], g["C_" + tag], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(x.numpy(), g["x_" + tag], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(float(ld.numpy()), float(g["ldet_" + tag]), rtol=1e-12)


def test_block_banded_solve_batched_and_dense_inverse(backend):
    """Several chains in one launch, checked against the dense inverse of the assembled matrix."""
    rs = np.random.RandomState(2)
    batch, T, Dm = 3, 9, 4
    A = np.empty((batch, T, Dm, Dm)); B = np.empty((batch, T - 1, Dm, Dm)); y = rs.randn(batch, T, Dm)
    dense = []
    for b in range(batch):
        R = rs.randn(T * Dm, T * Dm)
        full = R @ R.T + T * Dm * np.identity(T * Dm)
        for n in range(T):
            A[b, n] = full[n * Dm:(n + 1) * Dm, n * Dm:(n + 1) * Dm]
            if n < T - 1:
                B[b, n] = full[n * Dm:(n + 1) * Dm, (n + 1) * Dm:(n + 2) * Dm]
        P = np.zeros_like(full)
        for n in range(T):
            P[n * Dm:(n + 1) * Dm, n * Dm:(n + 1) * Dm] = A[b, n]
            if n < T - 1:
                P[n * Dm:(n + 1) * Dm, (n + 1) * Dm:(n + 2) * Dm] = B[b, n]
                P[(n + 1) * Dm:(n + 2) * Dm, n * Dm:(n + 1) * Dm] = B[b, n].T
        dense.append(P)
    d = [DArray.from_numpy(a) for a in (A, B, y)]
    V, C, x, ld = DArray.empty((batch, T, Dm, Dm)), DArray.empty((batch, T - 1, Dm, Dm)), DArray.empty((batch, T, Dm)), \
        DArray.empty((batch,))
    backend.block_banded_solve(d[0].ptr, d[1].ptr, d[2].ptr, batch, T, Dm, V.ptr, C.ptr, x.ptr, ld.ptr, True)
    Vn, Cn, xn, ldn = V.numpy(), C.numpy(), x.numpy(), ld.numpy()
    for b in range(batch):
        Pinv = np.linalg.inv(dense[b])
        for n in range(T):
            np.testing.assert_allclose(Vn[b, n], Pinv[n * Dm:(n + 1) * Dm, n * Dm:(n + 1) * Dm], rtol=1e-8, atol=1e-12)
            if n < T - 1:
                np.testing.assert_allclose(Cn[b, n], Pinv[n * Dm:(n + 1) * Dm, (n + 1) * Dm:(n + 2) * Dm], rtol=1e-8,
                                           atol=1e-12)
        np.testing.assert_allclose(xn[b].ravel(), Pinv @ y[b].ravel(), rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(ldn[b], np.linalg.slogdet(dense[b])[1], rtol=1e-11)


def test_block_banded_not_spd(backend):
    from bayespy_b200 import _bpk
    from oracle import bpk_ref
    T, Dm = 4, 3
    A = np.tile(np.identity(Dm), (T, 1, 1))
    A[2] = -np.identity(Dm)
    B = np.zeros((T - 1, Dm, Dm)); y = np.ones((T, Dm))
    d = [DArray.from_numpy(a) for a in (A, B, y)]
    V, C, x, ld = DArray.empty((T, Dm, Dm)), DArray.empty((T - 1, Dm, Dm)), DArray.empty((T, Dm)), DArray.empty(())
    with pytest.raises((_bpk.NotPositiveDefinite, bpk_ref.NotPositiveDefinite)):
        backend.block_banded_solve(d[0].ptr, d[1].ptr, d[2].ptr, 1, T, Dm, V.ptr, C.ptr, x.ptr, ld.ptr, True)


def _lssm(g, M, N, Dm, masked):
    from bayespy_b200.nodes import GaussianARD, GaussianMarkovChain, Gamma, Dot
    from bayespy_b200.inference import VB
    alpha = Gamma(1e-5, 1e-5, plates=(Dm,), name="alpha")
    A = GaussianARD(0, alpha, shape=(Dm,), plates=(Dm,), name="A")
    X = GaussianMarkovChain(np.zeros(Dm), 1e-3 * np.identity(Dm), A, np.ones(Dm), n=N, name="X")
    gamma = Gamma(1e-5, 1e-5, plates=(Dm,), name="gamma")
    C = GaussianARD(0, gamma, shape=(Dm,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    C.initialize_from_value(g["C_init"])
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    if masked:
        Y.observe(g["y"], mask=g["mask"])
    else:
        Y.observe(g["y"])
    Q = VB(X, C, gamma, A, alpha, tau, Y)
    return Q, dict(X=X, C=C, gamma=gamma, A=A, alpha=alpha, tau=tau)


@pytest.mark.parametrize("name,masked", [("lssm_small", False), ("lssm_masked", True)])
def test_lssm_matches_reference(backend, name, masked):
    g = golden(name)
    M, N, Dm = 6, 40, 3
    Q, nodes = _lssm(g, M, N, Dm, masked)
    assert F_plates(Q) == (M, N)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    for node in Q.model:
        np.testing.assert_allclose(Q.l[node][:iters], g["l_" + node.name], rtol=1e-6, atol=1e-6, err_msg=node.name)
    for nm, node in nodes.items():
        for i in range(len(node.u)):
            np.testing.assert_allclose(np.asarray(node.u[i]), g["%s_u%d" % (nm, i)], rtol=1e-6, atol=1e-9,
                                       err_msg="%s.u[%d]" % (nm, i))
        for i in range(len(node.phi)):
            ref = g["%s_phi%d" % (nm, i)]
            np.testing.assert_allclose(np.broadcast_to(np.asarray(node.phi[i]), ref.shape), ref, rtol=1e-6, atol=1e-9,
                                       err_msg="%s.phi[%d]" % (nm, i))
        np.testing.assert_allclose(np.broadcast_to(np.asarray(node.g), np.shape(g[nm + "_g"])), g[nm + "_g"], rtol=1e-6,
                                   atol=1e-8, err_msg=nm + ".g")


def F_plates(Q):
    return tuple(Q["Y"].plates)


@pytest.mark.parametrize("T,Dm", [(777, 6), (64, 32), (1025, 3), (9, 2), (300, 12), (130, 16), (50, 40), (2049, 32)])
def test_block_banded_long_chain_vs_oracle(backend, T, Dm):
    """Chains long enough for the parallel-in-time (block cyclic reduction) path, any T (not only powers of two),
    against the oracle's sequential restatement of linalg.block_banded_solve."""
    from oracle.bpk_ref import RefBackend
    rs = np.random.RandomState(T + Dm)
    a = 0.9 * np.linalg.qr(rs.randn(Dm, Dm))[0]
    A = np.tile(2.5 * np.identity(Dm) + a.T @ a, (T, 1, 1)) + 0.05 * np.eye(Dm) * rs.rand(T, 1, 1)
    B = np.tile(-a.T, (T - 1, 1, 1)) * (1 + 0.1 * rs.randn(T - 1, 1, 1))
    y = rs.randn(T, Dm)
    ref = RefBackend()
    hV, hC, hx, hl = np.empty((T, Dm, Dm)), np.empty((T - 1, Dm, Dm)), np.empty((T, Dm)), np.empty(1)
    ref.block_banded_solve(A.ctypes.data, B.ctypes.data, y.ctypes.data, 1, T, Dm, hV.ctypes.data, hC.ctypes.data,
                           hx.ctypes.data, hl.ctypes.data)
    d = [DArray.from_numpy(v) for v in (A, B, y)]
    V, C, x, ld = DArray.empty((T, Dm, Dm)), DArray.empty((T - 1, Dm, Dm)), DArray.empty((T, Dm)), DArray.empty(())
    backend.block_banded_solve(d[0].ptr, d[1].ptr, d[2].ptr, 1, T, Dm, V.ptr, C.ptr, x.ptr, ld.ptr, True)
    np.testing.assert_allclose(V.numpy(), hV, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(C.numpy(), hC, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(x.numpy(), hx, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(float(ld.numpy()), hl[0], rtol=1e-12)


def test_utils_linalg_block_banded_solve_api(backend):
    """The drop-in function of seam 1 (utils.linalg.block_banded_solve): NumPy in -> NumPy out, plates broadcast."""
    from bayespy_b200.utils import linalg
    g = golden("block_banded")
    V, C, x, ld = linalg.block_banded_solve(g["A_b"], g["B_b"], g["y_b"])
    np.testing.assert_allclose(V, g["V_b"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(C, g["C_b"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(x, g["x_b"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(ld, g["ldet_b"], rtol=1e-12)
    # two right-hand sides against one matrix: plates (2,) from y only
    y2 = np.stack([g["y_b"], 2 * g["y_b"]])
    V2, C2, x2, ld2 = linalg.block_banded_solve(g["A_b"], g["B_b"], y2)
    assert V2.shape == (2,) + g["V_b"].shape and ld2.shape == (2,)
    np.testing.assert_allclose(x2[1], 2 * g["x_b"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(V2[1], g["V_b"], rtol=1e-9, atol=1e-12)
    with pytest.raises(ValueError):
        linalg.block_banded_solve(g["A_b"][:-1], g["B_b"], g["y_b"])


def test_plated_chains_match_reference(backend):
    """Two independent chains (plates (2,)) sharing A and C (gaussian_markov_chain.py with plates): bound trajectory
    and moments against the reference (oracle on CPU, libbpk under -m gpu)."""
    from bayespy_b200.nodes import GaussianARD, GaussianMarkovChain, Gamma, Dot
    from bayespy_b200.inference import VB
    g = golden("lssm_plated")
    M, P, N = g["y"].shape
    Dm = g["mu0"].shape[-1]
    alpha = Gamma(1e-5, 1e-5, plates=(Dm,), name="alpha")
    A = GaussianARD(0, alpha, shape=(Dm,), plates=(Dm,), name="A")
    X = GaussianMarkovChain(g["mu0"], 1e-3 * np.identity(Dm), A, np.ones(Dm), n=N, name="X")
    assert tuple(X.plates) == (P,)
    gamma = Gamma(1e-5, 1e-5, plates=(Dm,), name="gamma")
    C = GaussianARD(0, gamma, shape=(Dm,), plates=(M, 1, 1), name="C")
    F = Dot(C, X, name="F")
    assert tuple(F.plates) == (M, P, N)
    C.initialize_from_value(g["C_init"])
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(g["y"])
    Q = VB(X, C, gamma, A, alpha, tau, Y)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    for nm, node in (("X", X), ("C", C), ("A", A), ("alpha", alpha), ("tau", tau)):
        for i in range(len(node.u)):
            np.testing.assert_allclose(np.asarray(node.u[i]), g["%s_u%d" % (nm, i)], rtol=1e-6, atol=1e-9,
                                       err_msg="%s.u[%d]" % (nm, i))


def _check_nodes(g, nodes, rtol=1e-6):
    for nm, node in nodes:
        for i in range(len(node.u)):
            np.testing.assert_allclose(np.asarray(node.u[i]), g["%s_u%d" % (nm, i)], rtol=rtol, atol=1e-9,
                                       err_msg="%s.u[%d]" % (nm, i))


def test_time_varying_dynamics_match_reference(backend):
    """A and nu with plates (N-1, D): one transition matrix / innovation precision per step, chain length inferred
    from the parents (gaussian_markov_chain.py:660-706, 840-880) — the reference's own plate layout."""
    from bayespy_b200.nodes import GaussianARD, GaussianMarkovChain, Gamma, Dot
    from bayespy_b200.inference import VB
    g = golden("lssm_varying")
    M, N = g["y"].shape
    Dm = g["C_init"].shape[-1]
    alpha = Gamma(1e-5, 1e-5, plates=(Dm,), name="alpha")
    A = GaussianARD(0, alpha, shape=(Dm,), plates=(N - 1, Dm), name="A")
    A.initialize_from_value(g["A_init"])
    nu = Gamma(1e-3, 1e-3, plates=(N - 1, Dm), name="nu")
    X = GaussianMarkovChain(np.zeros(Dm), 1e-3 * np.identity(Dm), A, nu, name="X")
    assert tuple(X.plates) == () and tuple(X.dims[0]) == (N, Dm)
    gamma = Gamma(1e-5, 1e-5, plates=(Dm,), name="gamma")
    C = GaussianARD(0, gamma, shape=(Dm,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    C.initialize_from_value(g["C_init"])
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(g["y"])
    Q = VB(X, C, gamma, A, alpha, nu, tau, Y)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    _check_nodes(g, (("X", X), ("C", C), ("A", A), ("alpha", alpha), ("nu", nu), ("tau", tau)))


def test_plated_dynamics_match_reference(backend):
    """A with plates (P, 1, D): P independent chains, each with its own time-invariant transition matrix."""
    from bayespy_b200.nodes import GaussianARD, GaussianMarkovChain, Gamma, Dot
    from bayespy_b200.inference import VB
    g = golden("lssm_plated_dynamics")
    M, P, N = g["y"].shape
    Dm = g["C_init"].shape[-1]
    alpha = Gamma(1e-5, 1e-5, plates=(Dm,), name="alpha")
    A = GaussianARD(0, alpha, shape=(Dm,), plates=(P, 1, Dm), name="A")
    A.initialize_from_value(g["A_init"])
    X = GaussianMarkovChain(np.zeros(Dm), 1e-3 * np.identity(Dm), A, np.ones(Dm), n=N, name="X")
    assert tuple(X.plates) == (P,)
    gamma = Gamma(1e-5, 1e-5, plates=(Dm,), name="gamma")
    C = GaussianARD(0, gamma, shape=(Dm,), plates=(M, 1, 1), name="C")
    F = Dot(C, X, name="F")
    C.initialize_from_value(g["C_init"])
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(g["y"])
    Q = VB(X, C, gamma, A, alpha, tau, Y)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    _check_nodes(g, (("X", X), ("C", C), ("A", A), ("alpha", alpha), ("tau", tau)))


def test_dynamics_plates_follow_the_reference_layout(oracle_backend):
    """(N-1, D) is ONE time-varying chain, not N-1 chains; a wrong time extent is rejected."""
    from bayespy_b200.nodes import GaussianARD, GaussianMarkovChain
    Dm, N = 2, 6
    A = GaussianARD(0, 1, shape=(Dm,), plates=(N - 1, Dm))
    X = GaussianMarkovChain(np.zeros(Dm), np.identity(Dm), A, np.ones(Dm), n=N)
    assert tuple(X.plates) == () and np.asarray(X.u[0]).shape == (N, Dm)
    with pytest.raises(ValueError):
        GaussianMarkovChain(np.zeros(Dm), np.identity(Dm), A, np.ones(Dm), n=N + 2)
    A3 = GaussianARD(0, 1, shape=(Dm,), plates=(4, 1, Dm))
    X3 = GaussianMarkovChain(np.zeros(Dm), np.identity(Dm), A3, np.ones(Dm), n=N)
    assert tuple(X3.plates) == (4,)
    with pytest.raises(ValueError):
        GaussianMarkovChain(np.zeros(Dm), np.identity(Dm), GaussianARD(0, 1, shape=(Dm,), plates=(Dm,)), np.ones(Dm))


def test_varying_chain_mixing_matrices_matches_reference(backend):
    """VaryingGaussianMarkovChain (gaussian_markov_chain.py:930-1452): A_n = sum_k s_nk B_k; the chain, the mixing
    matrices B (shape (D, K), plates (D,)), their ARD prior, the weights S and the observation model against the
    unmodified reference: bound trajectory, every bound term, every node's moments and natural parameters."""
    from bayespy_b200.nodes import GaussianARD, VaryingGaussianMarkovChain, Gamma, Dot
    from bayespy_b200.inference import VB
    g = golden("lssm_mixing")
    y = g["y"]
    M, N = y.shape
    Dm, K = g["B_init"].shape[1:]
    beta = Gamma(1e-3, 1e-3, plates=(K,), name="beta")
    B = GaussianARD(0, beta, shape=(Dm, K), plates=(Dm,), name="B")
    B.initialize_from_value(g["B_init"])
    S = GaussianARD(0, 1, shape=(K,), plates=(N - 1,), name="S")
    S.initialize_from_value(g["S_init"])
    X = VaryingGaussianMarkovChain(np.zeros(Dm), 1e-3 * np.identity(Dm), B, S, g["nu"], name="X")
    assert tuple(X.plates) == () and tuple(X.dims[0]) == (N, Dm)
    gamma = Gamma(1e-5, 1e-5, plates=(Dm,), name="gamma")
    C = GaussianARD(0, gamma, shape=(Dm,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    C.initialize_from_value(g["C_init"])
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(X, C, gamma, B, beta, S, tau, Y)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    for node in Q.model:
        np.testing.assert_allclose(Q.l[node][:iters], g["l_" + node.name], rtol=1e-7, atol=1e-6, err_msg=node.name)
    for nm, node in (("X", X), ("C", C), ("B", B), ("beta", beta), ("S", S), ("tau", tau)):
        for i, u in enumerate(node.u):
            ref = g["%s_u%d" % (nm, i)]
            np.testing.assert_allclose(np.asarray(u).reshape(ref.shape), ref, rtol=1e-7, atol=1e-9 * np.max(np.abs(ref)),
                                       err_msg="%s.u[%d]" % (nm, i))
        for i, ph in enumerate(node.phi):
            ref = g["%s_phi%d" % (nm, i)]
            np.testing.assert_allclose(np.broadcast_to(np.asarray(ph), ref.shape), ref, rtol=1e-7,
                                       atol=1e-9 * np.max(np.abs(ref)), err_msg="%s.phi[%d]" % (nm, i))
    # wrong shapes are refused like in the reference
    with pytest.raises(ValueError):
        VaryingGaussianMarkovChain(np.zeros(Dm), np.identity(Dm), GaussianARD(0, 1, shape=(Dm, K), plates=(Dm + 1,)), S,
                                   np.ones(Dm))
    with pytest.raises(ValueError):
        VaryingGaussianMarkovChain(np.zeros(Dm), np.identity(Dm), B, GaussianARD(0, 1, shape=(K + 1,), plates=(N - 1,)),
                                   np.ones(Dm))


def test_switching_chain_matches_reference(backend):
    """SwitchingGaussianMarkovChain (gaussian_markov_chain.py:1454-1985): A_n = B_{z_n}, z_n ~ Categorical(pi); chain,
    transition matrices, selector and its Dirichlet prior against the unmodified reference."""
    from bayespy_b200.nodes import (GaussianARD, SwitchingGaussianMarkovChain, Gamma, Dot, Dirichlet, Categorical)
    from bayespy_b200.inference import VB
    g = golden("lssm_switching")
    y = g["y"]
    M, N = y.shape
    K, Dm = g["B_init"].shape[:2]
    beta = Gamma(1e-3, 1e-3, plates=(K, 1, 1), name="beta")
    B = GaussianARD(0, beta, shape=(Dm,), plates=(K, Dm), name="B")
    B.initialize_from_value(g["B_init"])
    pi = Dirichlet(np.ones(K), name="pi")
    Z = Categorical(pi, plates=(N - 1,), name="Z")
    Z.initialize_from_value(g["Z_init"])
    X = SwitchingGaussianMarkovChain(np.zeros(Dm), 1e-3 * np.identity(Dm), B, Z, g["nu"], name="X")
    assert tuple(X.plates) == () and tuple(X.dims[0]) == (N, Dm)
    gamma = Gamma(1e-5, 1e-5, plates=(Dm,), name="gamma")
    C = GaussianARD(0, gamma, shape=(Dm,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    C.initialize_from_value(g["C_init"])
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(X, C, gamma, B, beta, Z, pi, tau, Y)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    for node in Q.model:
        np.testing.assert_allclose(Q.l[node][:iters], g["l_" + node.name], rtol=1e-7, atol=1e-6, err_msg=node.name)
    for nm, node in (("X", X), ("C", C), ("B", B), ("beta", beta), ("Z", Z), ("pi", pi), ("tau", tau)):
        for i, u in enumerate(node.u):
            ref = g["%s_u%d" % (nm, i)]
            np.testing.assert_allclose(np.asarray(u).reshape(ref.shape), ref, rtol=1e-7, atol=1e-9 * np.max(np.abs(ref)),
                                       err_msg="%s.u[%d]" % (nm, i))
    with pytest.raises(ValueError):
        SwitchingGaussianMarkovChain(np.zeros(Dm), np.identity(Dm), B, Categorical(np.ones(K + 1) / (K + 1), plates=(N - 1,)),
                                     np.ones(Dm))


def _check_state(g, pairs, rtol=1e-6):
    def same(a, ref, msg):
        a, ref = np.broadcast_arrays(np.asarray(a), ref)       # either side may hold an axis in broadcast (unit) form
        np.testing.assert_allclose(a, ref, rtol=rtol, atol=1e-8, err_msg=msg)
    for nm, node in pairs:
        for i in range(len(node.u)):
            same(node.u[i], g["%s_u%d" % (nm, i)], "%s.u[%d]" % (nm, i))
        for i in range(len(node.phi)):
            same(node.phi[i], g["%s_phi%d" % (nm, i)], "%s.phi[%d]" % (nm, i))
        same(node.g, g[nm + "_g"], nm + ".g")


def test_chain_driven_by_known_input_signals_matches_reference(backend):
    """gaussian_markov_chain.py:485-540, :608-616, :638-655: x_n = [A B] [x_{n-1}; z_{n-1}] + noise, z an array."""
    from bayespy_b200.nodes import GaussianMarkovChain, GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
    g = golden("lssm_inputs")
    M, N, Dm, K = 4, 25, 2, 2
    alpha = Gamma(1e-3, 1e-3, plates=(Dm + K,), name="alpha")
    A = GaussianARD(0, alpha, shape=(Dm + K,), plates=(Dm,), name="A")
    X = GaussianMarkovChain(np.zeros(Dm), 1e-2 * np.identity(Dm), A, np.ones(Dm), inputs=g["a_z"], n=N, name="X")
    C = GaussianARD(0, 1e-2, shape=(Dm,), plates=(M, 1), name="C")
    C.initialize_from_value(g["a_Cinit"])
    F = SumMultiply("i,i", C, X, name="F")
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(g["a_y"])
    Q = VB(X, C, A, alpha, tau, Y)
    iters = len(g["a_L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["a_L"], rtol=1e-8)
    _check_state(g, (("a_X", X), ("a_C", C), ("a_A", A), ("a_alpha", alpha), ("a_tau", tau)))
    with pytest.raises(ValueError):
        GaussianMarkovChain(np.zeros(Dm), np.identity(Dm), GaussianARD(0, 1, shape=(Dm,), plates=(Dm,)), np.ones(Dm),
                            inputs=g["a_z"], n=N)                      # rows of A must have length D + K
    with pytest.raises(ValueError):
        GaussianMarkovChain(np.zeros(Dm), np.identity(Dm), A, np.ones(Dm), inputs=g["a_z"], n=N + 3)


def test_plated_chains_with_uncertain_inputs_match_reference(backend):
    """Inputs as a Gaussian node (it receives the chain's message, :504-527), chain plates, a masked observation."""
    from bayespy_b200.nodes import GaussianMarkovChain, GaussianARD, Gaussian, Gamma
    from bayespy_b200.inference import VB
    g = golden("lssm_inputs")
    N, Dm, K, P = 25, 2, 2, 3
    U = GaussianARD(g["b_Umean"], 4.0, shape=(K,), plates=(P, N - 1), name="U")
    A2 = GaussianARD(0, 1.0, shape=(Dm + K,), plates=(P, 1, Dm), name="A2")
    A2.initialize_from_value(g["b_A2init"])
    nu2 = Gamma(2.0, 2.0, plates=(P, 1, Dm), name="nu2")
    X2 = GaussianMarkovChain(np.zeros(Dm), np.identity(Dm), A2, nu2, inputs=U, name="X2")
    assert X2.plates == (P,) and X2.N == N
    Y2 = Gaussian(X2, 5.0 * np.identity(Dm), name="Y2")
    Y2.observe(g["b_y"], mask=g["b_mask"])
    Q = VB(X2, A2, nu2, U, Y2)
    iters = len(g["b_L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["b_L"], rtol=1e-8)
    _check_state(g, (("b_X2", X2), ("b_A2", A2), ("b_nu2", nu2), ("b_U", U)))


def test_chain_with_gaussian_gamma_dynamics_matches_reference(backend):
    """A GaussianGamma node as the dynamics (gaussian_markov_chain.py:817 joins its scale with nu)."""
    from bayespy_b200.nodes import GaussianMarkovChain, GaussianGamma, Gaussian, Gamma
    from bayespy_b200.inference import VB
    g = golden("lssm_inputs")
    N, Dm, K = 25, 2, 2
    b3 = Gamma(2.0, 1.0, plates=(Dm,), name="b3")
    A3 = GaussianGamma(np.zeros(Dm + K), np.identity(Dm + K), 2.0, b3, plates=(Dm,), name="A3")
    X3 = GaussianMarkovChain(np.zeros(Dm), np.identity(Dm), A3, np.ones(Dm), inputs=g["a_z"], n=N, name="X3")
    Y3 = Gaussian(X3, 10.0 * np.identity(Dm), name="Y3")
    Y3.observe(g["c_y"])
    Q = VB(X3, A3, b3, Y3)
    iters = len(g["c_L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["c_L"], rtol=1e-8)
    _check_state(g, (("c_X3", X3), ("c_A3", A3), ("c_b3", b3)))


def test_plated_varying_chains_match_reference(backend):
    """VaryingGaussianMarkovChain over chain plates (mixing matrices (P, D), weights (P, N-1), innovation precision
    (P, 1, D)) against the reference's dedicated node, masked observation."""
    from bayespy_b200.nodes import GaussianARD, Gaussian, VaryingGaussianMarkovChain
    from bayespy_b200.inference import VB
    g = golden("lssm_mixing_plated")
    P, N, Dm = g["y"].shape
    K = g["S_init"].shape[-1]
    B = GaussianARD(0, 0.5, shape=(Dm, K), plates=(P, Dm), name="B")
    B.initialize_from_value(g["B_init"])
    S = GaussianARD(0, 1, shape=(K,), plates=(P, N - 1), name="S")
    S.initialize_from_value(g["S_init"])
    X = VaryingGaussianMarkovChain(np.zeros(Dm), np.identity(Dm), B, S, g["nu"], name="X")
    assert tuple(X.plates) == (P,) and tuple(X.dims[0]) == (N, Dm)
    Y = Gaussian(X, 4.0 * np.identity(Dm), name="Y")
    Y.observe(g["y"], mask=g["mask"])
    Q = VB(X, B, S, Y)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    _check_state(g, (("X", X), ("B", B), ("S", S)))
