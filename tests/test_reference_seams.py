"""The reference's OWN unit tests for seam 1 ("implication (1)"), pointed at the drop-ins:

    bayespy/utils/tests/test_misc.py:157-340    TestSumMultiply   -> bayespy_b200.utils.misc.sum_multiply
    bayespy/utils/tests/test_linalg.py:112-182  TestBandedSolve   -> bayespy_b200.utils.linalg.block_banded_solve

The test classes are imported from the staged, unmodified reference package (oracle/_ref; it travels to the GPU box),
and only the function under test is swapped inside the reference's module for the duration of the test — the
assertions, shapes and random inputs are the reference's.  Oracle backend on CPU (host logic), libbpk under -m gpu.
"""
import numpy as np
import pytest


def _reference_tests():
    from oracle import make_ref
    make_ref.build()
    if not make_ref.available():
        pytest.skip("oracle/_ref is not staged and /root/reference is absent")
    make_ref.import_reference()
    from bayespy.utils.tests import test_misc, test_linalg
    return test_misc, test_linalg


def test_reference_TestSumMultiply_against_the_drop_in(backend, monkeypatch):
    test_misc, _ = _reference_tests()
    from bayespy_b200.utils import misc as ours
    calls = []

    def sum_multiply(*args, **kwargs):
        calls.append(len(args))
        return ours.sum_multiply(*args, **kwargs)
    monkeypatch.setattr(test_misc.misc, "sum_multiply", sum_multiply)
    np.random.seed(0)
    case = test_misc.TestSumMultiply("test_sum_multiply")
    case.debug()                       # raises on the first failing assertion of the reference's test
    assert len(calls) > 20             # the reference's checks really went through the drop-in


def test_reference_TestBandedSolve_against_the_drop_in(backend, monkeypatch):
    _, test_linalg = _reference_tests()
    from bayespy_b200.utils import linalg as ours
    calls = []

    def block_banded_solve(A, B, y):
        calls.append(np.shape(A))
        return ours.block_banded_solve(A, B, y)
    monkeypatch.setattr(test_linalg.linalg, "block_banded_solve", block_banded_solve)
    np.random.seed(0)
    test_linalg.TestBandedSolve("test_block_banded_solve").debug()
    assert calls == [(40, 5, 5)]


def test_reference_chol_family_against_the_drop_in(backend):
    """chol* has no unit test of its own in the reference: A/B against the reference's functions
    on the batched shapes of the path, including the ndim=0 scalar fast paths (linalg.py:40-41,71-72,176-177)."""
    _reference_tests()
    from bayespy.utils import linalg as ref
    from bayespy_b200.utils import linalg as ours
    rs = np.random.RandomState(5)
    for batch, Dm in (((), 1), ((3,), 4), ((2, 3), 8), ((5,), 16), ((2,), 33)):
        W = rs.randn(*batch, Dm, 2 * Dm)
        C = W @ np.swapaxes(W, -1, -2) + 0.1 * np.identity(Dm)
        b = rs.randn(*batch, Dm)
        U_ref, U = ref.chol(C), ours.chol(C)
        np.testing.assert_allclose(np.triu(U), np.triu(U_ref), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(ours.chol_solve(U, b), ref.chol_solve(U_ref, b), rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(ours.chol_inv(U), ref.chol_inv(U_ref), rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(ours.chol_logdet(U), ref.chol_logdet(U_ref), rtol=1e-12)
    c = np.abs(rs.randn(7)) + 0.5
    np.testing.assert_allclose(ours.chol(c, ndim=0), ref.chol(c, ndim=0), rtol=1e-14)
    np.testing.assert_allclose(ours.chol_logdet(ours.chol(c, ndim=0), ndim=0), ref.chol_logdet(ref.chol(c, ndim=0), ndim=0),
                               rtol=1e-13)
