"""The device-resident VB loop (bpk_pca_vb_run, csrc/pca_vb.cu) against the per-node path.

Both paths run the same model on the same backend; the per-node path is the one pinned to the
reference's goldens in test_models.py, so agreement here extends that pin to the resident loop.
Covers: every update order, non-zero prior means, constant alpha / tau, the device-side
convergence stop (must end at the same iteration), chunked updates, and the generic-shape path."""
import ctypes
import itertools
import os

import numpy as np
import pytest

from conftest import ROOT


def _pca(y, K, order="XCat", mux=0.0, ax=1.0, muc=0.0, alpha_const=None, tau_const=None, resident=True,
         seed=3):
    from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
    M, N = y.shape
    X = GaussianARD(mux, ax, plates=(1, N), shape=(K,), name="X")
    alpha = Gamma(1e-3, 1e-3, plates=(K,), name="alpha") if alpha_const is None else alpha_const
    C = GaussianARD(muc, alpha, plates=(M, 1), shape=(K,), name="C")
    F = SumMultiply("d,d->", X, C)
    tau = Gamma(1e-3, 1e-3, name="tau") if tau_const is None else tau_const
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    C.initialize_from_value(np.random.RandomState(seed).randn(M, 1, K))
    nodes = dict(X=X, C=C, a=alpha, t=tau)
    model = [Y] + [nodes[c] for c in order if not isinstance(nodes[c], (int, float, np.ndarray))]
    Q = VB(*model, resident=resident)
    return Q, dict(X=X, C=C, alpha=alpha, tau=tau, Y=Y)


def _data(M, N, K, seed=0):
    rs = np.random.RandomState(seed)
    return rs.randn(M, 3) @ rs.randn(3, N) + 0.3 * rs.randn(M, N)


def _compare(Qa, na, Qb, nb, iters, rtol=1e-9):
    np.testing.assert_allclose(Qa.L[:iters], Qb.L[:iters], rtol=rtol)
    for node in Qa.model:
        np.testing.assert_allclose(Qa.l[node][:iters], Qb.l[Qb[node.name]][:iters], rtol=1e-7, atol=1e-6)
    for name in ("X", "C", "alpha", "tau"):
        a, b = na[name], nb[name]
        if not hasattr(a, "u"):
            continue
        for i in range(2):
            np.testing.assert_allclose(np.asarray(a.u[i]), np.asarray(b.u[i]), rtol=1e-7, atol=1e-9,
                                       err_msg="%s.u[%d]" % (name, i))
            np.testing.assert_allclose(np.broadcast_to(np.asarray(a.phi[i]), np.shape(np.asarray(b.phi[i]))),
                                       np.asarray(b.phi[i]), rtol=1e-7, atol=1e-9, err_msg="%s.phi[%d]" % (name, i))
        np.testing.assert_allclose(np.broadcast_to(np.asarray(a.g), np.shape(np.asarray(b.g))), np.asarray(b.g),
                                   rtol=1e-7, atol=1e-8, err_msg="%s.g" % name)


@pytest.mark.parametrize("order", ["".join(p) for p in itertools.permutations("XCat")][::3])
def test_resident_matches_per_node_every_order(backend, order):
    y = _data(12, 150, 4)
    Qa, na = _pca(y, 4, order, resident=True)
    Qb, nb = _pca(y, 4, order, resident=False)
    Qa.update(repeat=6, verbose=False, tol=0)
    Qb.update(repeat=6, verbose=False, tol=0)
    assert Qa.plans[0].fused_calls >= 6
    _compare(Qa, na, Qb, nb, 6)


@pytest.mark.parametrize("M,N,K", [(64, 700, 16), (20, 129, 5), (80, 90, 3), (10, 60, 20)])
def test_resident_shapes_and_priors(backend, M, N, K):
    y = _data(M, N, K, seed=M)
    rs = np.random.RandomState(1)
    kw = dict(mux=0.1 * rs.randn(K), ax=0.5 + rs.rand(K), muc=0.2 * rs.randn(K))
    Qa, na = _pca(y, K, resident=True, **kw)
    Qb, nb = _pca(y, K, resident=False, **kw)
    for _ in range(2):                       # two calls: re-entry from a published state
        Qa.update(repeat=3, verbose=False, tol=0)
        Qb.update(repeat=3, verbose=False, tol=0)
    _compare(Qa, na, Qb, nb, 6)


def test_resident_constant_alpha_tau(backend):
    y = _data(16, 200, 4)
    Qa, na = _pca(y, 4, alpha_const=0.7, tau_const=2.5, resident=True)
    Qb, nb = _pca(y, 4, alpha_const=0.7, tau_const=2.5, resident=False)
    Qa.update(repeat=4, verbose=False, tol=0)
    Qb.update(repeat=4, verbose=False, tol=0)
    assert Qa.plans[0].fused_calls >= 4
    np.testing.assert_allclose(Qa.L[:4], Qb.L[:4], rtol=1e-9)
    np.testing.assert_allclose(np.asarray(na["C"].u[0]), np.asarray(nb["C"].u[0]), rtol=1e-8, atol=1e-10)


def test_resident_converges_at_the_same_iteration(backend, capsys):
    y = _data(12, 300, 3, seed=4)
    Qa, na = _pca(y, 3, resident=True)
    Qb, nb = _pca(y, 3, resident=False)
    Qa.update(repeat=400, verbose=False, tol=1e-6)
    Qb.update(repeat=400, verbose=False, tol=1e-6)
    assert Qa.converged and Qb.converged
    assert Qa.iter == Qb.iter and Qa.iter < 400
    _compare(Qa, na, Qb, nb, Qa.iter)
    # verbose output keeps the reference's format (vmp.py:725,745)
    Qc, _ = _pca(y, 3, resident=True)
    Qc.update(repeat=400, tol=1e-6)
    out = capsys.readouterr().out.strip().splitlines()
    assert out[0].startswith("Iteration 1: loglike=") and out[-1] == "Converged at iteration %d." % Qa.iter


def test_resident_mixed_with_single_node_updates(backend):
    """A resident run followed by hand-driven node updates (and back) stays consistent."""
    y = _data(12, 150, 4, seed=7)
    Qa, na = _pca(y, 4, resident=True)
    Qb, nb = _pca(y, 4, resident=False)
    for Q, n in ((Qa, na), (Qb, nb)):
        Q.update(repeat=2, verbose=False, tol=0)
        Q.update(n["C"], n["tau"], repeat=1, verbose=False, tol=0)      # not a full sweep: per-node path
        Q.update(repeat=2, verbose=False, tol=0)
    _compare(Qa, na, Qb, nb, 5)


def test_layout_of_the_library_matches_the_oracle():
    """Host-only entry point: callable without a GPU."""
    from bayespy_b200 import _bpk
    from oracle.bpk_ref import RefBackend
    if not os.path.exists(_bpk.LIB_PATH):
        pytest.skip("libbpk.so not built")
    lib = ctypes.CDLL(_bpk.LIB_PATH)
    lib.bpk_pca_vb_field_name.restype = ctypes.c_char_p
    for M, K in ((64, 16), (7, 3), (100, 20)):
        n = ctypes.c_int()
        assert lib.bpk_pca_vb_layout(M, K, None, ctypes.byref(n)) == 0
        off = (ctypes.c_int64 * (n.value + 1))()
        assert lib.bpk_pca_vb_layout(M, K, off, ctypes.byref(n)) == 0
        lay, total = RefBackend().pca_vb_layout(M, K)
        assert total == off[n.value]
        for i in range(n.value):
            name = lib.bpk_pca_vb_field_name(i).decode()
            assert lay[name] == (off[i], off[i + 1] - off[i]), name


# ---- plan preconditions are re-validated on every call (round-1 advisor findings) -------------------------------
def test_observing_a_hyper_node_after_a_resident_run_leaves_the_resident_path(backend):
    """tau.observe(...) between two updates: the resident loop must not keep updating tau as latent."""
    y = _data(12, 150, 4, seed=9)
    out = []
    for resident, fused in ((True, True), (False, False)):
        from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
        from bayespy_b200.inference import VB
        M, N, K = 12, 150, 4
        X = GaussianARD(0, 1, plates=(1, N), shape=(K,), name="X")
        alpha = Gamma(1e-3, 1e-3, plates=(K,), name="alpha")
        C = GaussianARD(0, alpha, plates=(M, 1), shape=(K,), name="C")
        F = SumMultiply("d,d->", X, C)
        tau = Gamma(1e-3, 1e-3, name="tau")
        Y = GaussianARD(F, tau, name="Y")
        Y.observe(y)
        C.initialize_from_value(np.random.RandomState(3).randn(M, 1, K))
        Q = VB(Y, X, C, alpha, tau, resident=resident, fused=fused)
        Q.update(repeat=2, verbose=False, tol=0)
        tau.observe(7.0)
        Q.update(repeat=2, verbose=False, tol=0)
        out.append((Q.L[:4].copy(), np.asarray(tau.u[0]).copy(), np.asarray(C.u[0]).copy()))
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-9)
    np.testing.assert_allclose(out[0][1], 7.0)
    np.testing.assert_allclose(out[0][2], out[1][2], rtol=1e-8, atol=1e-10)


def test_a_child_added_after_vb_is_not_dropped(backend):
    """A second child hung on X after VB(...) was built: the fused X update would drop its message."""
    from bayespy_b200.nodes import GaussianARD
    y = _data(12, 150, 4, seed=10)
    res = []
    for fused in (True, False):
        Q, n = _pca(y, 4, resident=fused)
        if not fused:
            for plan in Q.plans:          # take the plans out entirely: the per-node path pinned to the reference
                for k, fn in (("col.update", "update"), ("col.lb", "lower_bound_contribution")):
                    setattr(plan.col, fn, plan._orig[k])
                plan.F.message_to_parent = plan._orig["F.msg"]
                plan.Y.message_to_parent = plan._orig["Y.msg"]
                plan.Y.lower_bound_contribution = plan._orig["Y.lb"]
            Q.plans = []
        Q.update(repeat=2, verbose=False, tol=0)
        W = GaussianARD(n["X"], 5.0, name="W")
        W.observe(0.3 * np.ones((1, 150, 4)))
        Q.update(repeat=2, verbose=False, tol=0)
        res.append(np.asarray(n["X"].u[0]).copy())
    np.testing.assert_allclose(res[0], res[1], rtol=1e-8, atol=1e-10)
    # and the extra child really mattered (the result differs from the model without it)
    Q0, n0 = _pca(y, 4, resident=True)
    Q0.update(repeat=4, verbose=False, tol=0)
    assert np.max(np.abs(np.asarray(n0["X"].u[0]) - res[0])) > 1e-3


def test_callback_keeps_the_node_updates_on_the_fused_path(backend):
    """With a callback (here: the rotation of pca.rst:86-112) the host is needed between the node updates and the
    bound (vmp.py:702-713): every sweep is still ONE fused launch (sweep_resident), and the trajectory equals the
    per-node tier's for as long as rounding differences stay below the optimiser's sensitivity."""
    from bayespy_b200.inference.vmp.transformations import RotateGaussianARD, RotationOptimizer
    y = _data(20, 100, 6, seed=11)
    res = []
    for resident in (True, False):
        Q, n = _pca(y, 6, resident=resident)
        R = RotationOptimizer(RotateGaussianARD(n["X"]), RotateGaussianARD(n["C"], n["alpha"]), 6)
        Q.set_callback(R.rotate)
        before = Q.plans[0].fused_calls
        Q.update(repeat=5, verbose=False, tol=0)
        res.append((Q.L[:5].copy(), np.asarray(n["C"].u[0]).copy(), Q.plans[0].fused_calls - before))
    assert res[0][2] >= 5
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-7)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-5, atol=1e-7)


# ---- the mixture model's resident loop (bpk_gmm_vb_run, csrc/gmm_vb.cu) ---------------------------------------------
def _gmm(N, Dm, K, order="mLZa", resident=True, seed=5, a0=1e-5, m0=None, l0=1e-5, n0=None, v0=1e-5):
    from bayespy_b200.nodes import Gaussian, Wishart, Dirichlet, Categorical, Mixture
    from bayespy_b200.inference import VB
    rs = np.random.RandomState(seed)
    means = 4 * rs.randn(K, Dm)
    y = means[rs.randint(0, K, size=N)] + rs.randn(N, Dm)
    alpha = Dirichlet(a0 * np.ones(K), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="Z")
    mu = Gaussian(np.zeros(Dm) if m0 is None else m0, l0 * np.identity(Dm), plates=(K,), name="mu")
    Lam = Wishart(Dm if n0 is None else n0, v0 * np.identity(Dm), plates=(K,), name="Lambda")
    Y = Mixture(Z, Gaussian, mu, Lam, name="Y")
    Z.initialize_from_value(rs.randint(0, K, size=N))
    Y.observe(y)
    nodes = dict(m=mu, L=Lam, Z=Z, a=alpha)
    Q = VB(Y, *[nodes[c] for c in order], resident=resident)
    return Q, dict(mu=mu, Lambda=Lam, Z=Z, alpha=alpha, Y=Y)


def _compare_gmm(Qa, na, Qb, nb, iters, rtol=1e-9):
    # When Z is swept before mu and Lambda have seen any data, the FIRST bound is a difference of terms of size
    # ~1e13 (the vague initial q(mu), q(Lambda): tr(<mu mu^T><Lambda>) R_k ~ 1e5 * 1e5 * N), so its last bits are
    # summation-order noise (2^-9 observed); every later bound, and the state itself, is compared tightly.
    np.testing.assert_allclose(Qa.L[:1], Qb.L[:1], rtol=1e-5)
    np.testing.assert_allclose(Qa.L[1:iters], Qb.L[1:iters], rtol=rtol)
    for node in Qa.model:
        np.testing.assert_allclose(Qa.l[node][1:iters], Qb.l[Qb[node.name]][1:iters], rtol=1e-7, atol=1e-6,
                                   err_msg="bound term " + node.name)
        np.testing.assert_allclose(Qa.l[node][:1], Qb.l[Qb[node.name]][:1], rtol=1e-5, atol=1e-2,
                                   err_msg="first bound term " + node.name)
    # entries of a dying cluster (R_k -> 0 exponentially fast) are tiny and amplify rounding differences between
    # the two summation orders: tolerances are relative to the largest entry of each array
    for name in ("mu", "Lambda", "alpha", "Z"):
        a, b = na[name], nb[name]
        for i in range(len(b.u)):
            ub, pb = np.asarray(b.u[i]), np.asarray(b.phi[i])
            np.testing.assert_allclose(np.asarray(a.u[i]), ub, rtol=1e-7, atol=1e-7 * np.max(np.abs(ub)),
                                       err_msg="%s.u[%d]" % (name, i))
            np.testing.assert_allclose(np.broadcast_to(np.asarray(a.phi[i]), pb.shape), pb, rtol=1e-7,
                                       atol=1e-7 * np.max(np.abs(pb)), err_msg="%s.phi[%d]" % (name, i))
        gb = np.asarray(b.g)
        np.testing.assert_allclose(np.broadcast_to(np.asarray(a.g), gb.shape), gb, rtol=1e-7,
                                   atol=1e-7 * np.max(np.abs(gb)), err_msg="%s.g" % name)


@pytest.mark.parametrize("order", ["".join(p) for p in itertools.permutations("mLZa")][::3])
def test_gmm_resident_matches_per_node_every_order(backend, order):
    Qa, na = _gmm(240, 3, 5, order=order, resident=True)
    Qb, nb = _gmm(240, 3, 5, order=order, resident=False)
    Qa.update(repeat=6, verbose=False, tol=0)
    Qb.update(repeat=6, verbose=False, tol=0)
    assert Qa.plans[0]._res_cache is not None and getattr(Qb.plans[0], "_res_cache", None) is None
    _compare_gmm(Qa, na, Qb, nb, 6)
    # the bound recomputed node by node from the published state is the loop's own
    np.testing.assert_allclose(Qa.compute_lowerbound(), Qa.L[5], rtol=1e-10)


@pytest.mark.parametrize("N,Dm,K", [(500, 8, 16), (131, 2, 3), (300, 5, 40), (64, 1, 2)])
def test_gmm_resident_shapes_and_priors(backend, N, Dm, K):
    kw = dict(a0=0.7, m0=0.3 * np.arange(Dm), l0=0.02, n0=Dm + 2.5, v0=0.4)
    Qa, na = _gmm(N, Dm, K, resident=True, **kw)
    Qb, nb = _gmm(N, Dm, K, resident=False, **kw)
    Qa.update(repeat=4, verbose=False, tol=0)
    Qb.update(repeat=4, verbose=False, tol=0)
    _compare_gmm(Qa, na, Qb, nb, 4)


def test_gmm_resident_converges_at_the_same_iteration(backend, capsys):
    Qa, na = _gmm(200, 2, 4, resident=True)
    Qb, nb = _gmm(200, 2, 4, resident=False)
    Qa.update(repeat=400, tol=1e-6)
    Qb.update(repeat=400, tol=1e-6)
    out = capsys.readouterr().out
    assert out.count("Converged at iteration %d." % Qb.iter) == 2
    assert Qa.iter == Qb.iter and Qa.iter < 400
    _compare_gmm(Qa, na, Qb, nb, Qa.iter, rtol=1e-8)
    # silent run: chunks of sweeps between two host reads, same stopping point
    Qc, nc = _gmm(200, 2, 4, resident=True)
    Qc.update(repeat=400, tol=1e-6, verbose=False)
    assert Qc.iter == Qb.iter and Qc.converged
    _compare_gmm(Qc, nc, Qb, nb, Qc.iter, rtol=1e-8)


def test_gmm_resident_mixed_with_single_node_updates(backend):
    Qa, na = _gmm(150, 3, 4, resident=True)
    Qb, nb = _gmm(150, 3, 4, resident=False)
    for Q, n in ((Qa, na), (Qb, nb)):
        Q.update(repeat=2, verbose=False, tol=0)
        Q.update(n["mu"], n["alpha"], repeat=1, verbose=False, tol=0)      # not a full sweep: per-node path
        Q.update(repeat=2, verbose=False, tol=0)                           # state rebuilt from the nodes
        Q.update(repeat=1, verbose=False, tol=0)                           # state re-used
    _compare_gmm(Qa, na, Qb, nb, 6)


def test_gmm_callback_keeps_the_node_updates_on_the_device(backend):
    seen = []
    res = []
    for resident in (True, False):
        Q, n = _gmm(150, 3, 4, resident=resident)
        Q.set_callback(lambda: seen.append(float(np.asarray(n["alpha"].u[0])[0])))
        Q.update(repeat=4, verbose=False, tol=0)
        res.append((Q.L[:4].copy(), np.asarray(n["mu"].u[0]).copy()))
    assert len(seen) == 8
    np.testing.assert_allclose(seen[:4], seen[4:], rtol=1e-8)
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-9)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-7, atol=1e-9)


def test_gmm_observing_a_hyper_node_leaves_the_resident_path(backend):
    Q, n = _gmm(120, 2, 3, resident=True)
    Q.update(repeat=2, verbose=False, tol=0)
    assert Q.plans[0].resident_program(Q, Q.model) is not None
    n["alpha"].observe(np.full(3, 1.0 / 3))
    assert Q.plans[0].resident_program(Q, Q.model) is None
    Q.update(repeat=2, verbose=False, tol=0)            # generic path, no error
    assert np.isfinite(Q.L[3])


def test_gmm_layout_of_the_library_matches_the_oracle():
    from bayespy_b200 import _bpk
    from oracle.bpk_ref import RefBackend
    if not os.path.exists(_bpk.LIB_PATH):
        pytest.skip("libbpk.so not built")
    lib = ctypes.CDLL(_bpk.LIB_PATH)
    lib.bpk_gmm_vb_field_name.restype = ctypes.c_char_p
    for Dm, K in ((8, 64), (3, 5), (1, 2), (16, 128)):
        n = ctypes.c_int()
        assert lib.bpk_gmm_vb_layout(Dm, K, None, ctypes.byref(n)) == 0
        off = (ctypes.c_int64 * (n.value + 1))()
        assert lib.bpk_gmm_vb_layout(Dm, K, off, ctypes.byref(n)) == 0
        lay, total = RefBackend().gmm_vb_layout(Dm, K)
        assert total == off[n.value] and len(lay) == n.value
        for i in range(n.value):
            name = lib.bpk_gmm_vb_field_name(i).decode()
            assert lay[name][0] == off[i] and lay[name][1] <= off[i + 1] - off[i], name
