"""The per-thread 16 x 16 SPD solve / inverse / log-det of the masked factor-model sweep (csrc/spd16.cuh: block
Cholesky on 8 x 8 register tiles), compiled for the HOST and checked against NumPy — the arithmetic the reference does
per column with scipy cho_factor / cho_solve (linalg.py:31-223)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(ROOT, "tests", "spd16_host.cpp")
    out = os.path.join(ROOT, "tests", "_build", "libspd16_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    hdr = os.path.join(ROOT, "bayespy_b200", "csrc", "spd16.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    L = ctypes.CDLL(out)
    L.spd16_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return L


def _pack(A):
    return np.array([A[i, j] for i in range(16) for j in range(i, 16)])


def _unpack(p):
    A = np.zeros((16, 16))
    k = 0
    for i in range(16):
        for j in range(i, 16):
            A[i, j] = A[j, i] = p[k]
            k += 1
    return A


@pytest.mark.parametrize("seed,cond", [(0, 1.0), (1, 1e-3), (2, 1e3), (3, 1e-6)])
def test_spd16_matches_numpy(lib, seed, cond):
    rs = np.random.RandomState(seed)
    W = rs.randn(16, 40)
    A = W @ W.T * cond + np.diag(np.abs(rs.randn(16)) + 1e-3 * cond)
    phi = rs.randn(16)
    rows = np.zeros(lib.spd16_rows())
    rows[:136] = _pack(A)
    rows[136:152] = phi
    q, ld = ctypes.c_double(), ctypes.c_double()
    assert lib.spd16_host(rows.ctypes.data, ctypes.byref(q), ctypes.byref(ld)) == 0
    cov = np.linalg.inv(A)
    x = cov @ phi
    np.testing.assert_allclose(rows[136:152], x, rtol=1e-9, atol=1e-12 * np.abs(x).max())
    np.testing.assert_allclose(_unpack(rows[:136]), cov + np.outer(x, x), rtol=1e-8, atol=1e-10 * np.abs(cov).max())
    np.testing.assert_allclose(q.value, phi @ x, rtol=1e-10)
    np.testing.assert_allclose(ld.value, np.linalg.slogdet(A)[1], rtol=1e-11, atol=1e-10)


def test_spd16_flags_a_matrix_that_is_not_positive_definite(lib):
    A = np.identity(16)
    A[5, 5] = -1.0
    rows = np.zeros(lib.spd16_rows())
    rows[:136] = _pack(A)
    q, ld = ctypes.c_double(), ctypes.c_double()
    assert lib.spd16_host(rows.ctypes.data, ctypes.byref(q), ctypes.byref(ld)) == 1
    A = np.identity(16)
    A[12, 12] = 0.0
    rows[:136] = _pack(A)
    assert lib.spd16_host(rows.ctypes.data, ctypes.byref(q), ctypes.byref(ld)) == 1
