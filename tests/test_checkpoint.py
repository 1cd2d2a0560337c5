"""VB.save / VB.load / autosave (vmp.py:237-356, 750-758): a run resumed from a checkpoint continues exactly like the
uninterrupted one, the on-disk hierarchy is the reference's, and device-resident / virtual state is materialised."""
import numpy as np
import pytest


def _model(seed=5):
    from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
    from bayespy_b200.inference import VB
    rs = np.random.RandomState(seed)
    M, N, K = 10, 60, 3
    y = rs.randn(M, 2) @ rs.randn(2, N) + 0.2 * rs.randn(M, N)
    X = GaussianARD(0, 1, plates=(1, N), shape=(K,), name="X")
    alpha = Gamma(1e-3, 1e-3, plates=(K,), name="alpha")
    C = GaussianARD(0, alpha, plates=(M, 1), shape=(K,), name="C")
    F = SumMultiply("d,d->", X, C, name="F")
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    C.initialize_from_value(rs.randn(M, 1, K))
    return VB(Y, X, C, alpha, tau), dict(X=X, C=C, alpha=alpha, tau=tau, Y=Y)


def test_save_load_resumes_exactly(backend, tmp_path):
    fn = str(tmp_path / "run.ckpt")
    Qa, na = _model()
    Qa.update(repeat=7, verbose=False, tol=0)
    Qb, nb = _model()
    Qb.update(repeat=4, verbose=False, tol=0)
    Qb.save(filename=fn)
    Qc, nc = _model()                       # a fresh graph, e.g. another process
    Qc.load(filename=fn)
    assert Qc.iter == 4
    np.testing.assert_allclose(Qc.L[:4], Qa.L[:4], rtol=1e-12)
    Qc.update(repeat=3, verbose=False, tol=0)
    np.testing.assert_allclose(Qc.L[:7], Qa.L[:7], rtol=1e-10)
    for k in ("X", "C", "alpha", "tau"):
        for i in range(2):
            np.testing.assert_allclose(np.asarray(nc[k].u[i]), np.asarray(na[k].u[i]), rtol=1e-8, atol=1e-10)


def test_checkpoint_has_the_reference_hierarchy(backend, tmp_path):
    fn = str(tmp_path / "run.ckpt")
    Q, n = _model()
    Q.user_data = {"note": np.arange(3)}
    Q.update(repeat=2, verbose=False, tol=0)
    Q.save(filename=fn)
    from bayespy_b200.inference import checkpoint
    r = checkpoint._Reader(fn)
    for path in ("nodes/X/u0", "nodes/X/u1", "nodes/X/phi0", "nodes/X/phi1", "nodes/X/g", "nodes/X/f", "nodes/X/observed",
                 "nodes/tau/u0", "nodes/Y/observed", "L", "cputime", "iter", "converged", "boundterms/C", "user_data/note"):
        assert r.has(path), path
    assert r.get("nodes/X/u1").shape == (1, 60, 3, 3)          # the factored second moment is materialised
    assert int(r.get("iter")) == 2
    r.close()


def test_autosave_and_nodes_only(backend, tmp_path):
    from bayespy_b200.inference import VB
    fn = str(tmp_path / "auto.ckpt")
    Q, n = _model()
    Qs = VB(*Q.model, autosave_filename=fn, autosave_iterations=2)
    Qs.update(repeat=4, verbose=False, tol=0)
    Q2, n2 = _model()
    Q2.load("C", "tau", filename=fn, nodes_only=True)
    assert Q2.iter == 0
    np.testing.assert_allclose(np.asarray(n2["C"].u[0]), np.asarray(Qs["C"].u[0]), rtol=1e-12)
    # without any file name the checkpoint goes to a dated file in the temporary directory, like the reference's
    # (vmp.py:86-97: VB() picks the name, nothing is written before the first save)
    import os
    assert "vb_autosave_" in os.path.basename(Q2.autosave_filename) and not os.path.exists(Q2.autosave_filename)
    Q2.save()
    Q3, n3 = _model()
    Q3.load(filename=Q2.autosave_filename)
    np.testing.assert_allclose(np.asarray(n3["C"].u[0]), np.asarray(n2["C"].u[0]), rtol=1e-12)
    os.remove(Q2.autosave_filename)
