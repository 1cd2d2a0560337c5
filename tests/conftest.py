"""Test configuration.

Two backends run the same host-side code:
  * ``oracle``  — oracle/bpk_ref.py (NumPy restatement of every kernel); used by the
    ``-m "not gpu"`` tests to exercise the graph/plan/scheduler logic on a CPU box and to
    pin the oracle against the reference's golden vectors.  Test infrastructure only.
  * ``cuda``    — libbpk.so on cuda:0; the ``-m gpu`` tests are the parity tests proper.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


_cuda_backend = None


def _get_cuda():
    global _cuda_backend
    from bayespy_b200 import _bpk
    if _cuda_backend is None:
        _cuda_backend = _bpk.CudaBackend(0)      # raises without a GPU / library
    return _cuda_backend


@pytest.fixture
def oracle_backend():
    from bayespy_b200 import _bpk
    from oracle.bpk_ref import RefBackend
    be = RefBackend()
    old = _bpk._set_backend_for_testing(be)
    yield be
    _bpk._set_backend_for_testing(old)


@pytest.fixture
def cuda_backend():
    from bayespy_b200 import _bpk
    be = _get_cuda()
    old = _bpk._set_backend_for_testing(be)
    yield be
    be.sync()
    _bpk._set_backend_for_testing(old)


BACKENDS = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def backend(request):
    """Parametrised backend: 'oracle' on CPU, 'cuda' under -m gpu."""
    from bayespy_b200 import _bpk
    if request.param == "oracle":
        from oracle.bpk_ref import RefBackend
        be = RefBackend()
    else:
        be = _get_cuda()
    old = _bpk._set_backend_for_testing(be)
    yield be
    be.sync()
    _bpk._set_backend_for_testing(old)
