"""Timing of bpk_block_banded_solve at config 4 of BASELINE.json (T=1e5, D=32), next to the oracle
(NumPy/SciPy restatement of linalg.block_banded_solve) on a short chain.   python tools/bench_gmc.py [T] [D]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                      # noqa: E402
from bayespy_b200 import _bpk           # noqa: E402
from bayespy_b200.darray import DArray  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
Dm = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rs = np.random.RandomState(0)
# a stable LSSM-like precision: A_n = I + nu (I + a a'), B_n = -nu a'
a = 0.9 * np.linalg.qr(rs.randn(Dm, Dm))[0]
A = np.tile(2.5 * np.identity(Dm) + a.T @ a, (T, 1, 1))
B = np.tile(-a.T, (T - 1, 1, 1))
y = rs.randn(T, Dm)
be = _bpk.get()
Ad, Bd, yd = DArray.from_numpy(A), DArray.from_numpy(B), DArray.from_numpy(y)
V, C, x, ld = DArray.empty((T, Dm, Dm)), DArray.empty((T - 1, Dm, Dm)), DArray.empty((T, Dm)), DArray.empty(())
be.block_banded_solve(Ad.ptr, Bd.ptr, yd.ptr, 1, T, Dm, V.ptr, C.ptr, x.ptr, ld.ptr, True)
ts = []
for _ in range(3):
    t = be.timer_create()
    be.timer_record(t, 0)
    be.block_banded_solve(Ad.ptr, Bd.ptr, yd.ptr, 1, T, Dm, V.ptr, C.ptr, x.ptr, ld.ptr, False)
    be.timer_record(t, 1)
    ts.append(be.timer_elapsed_ms(t))
ms = float(np.median(ts))
# residual check: P x = y on a few rows
xs = x.numpy()
r = A[1] @ xs[1] + B[0].T @ xs[0] + B[1] @ xs[2] - y[1]
print("block_banded_solve T=%d D=%d: %.2f ms (%.2f us/step), residual %.2e, logdet %.6e"
      % (T, Dm, ms, 1e3 * ms / T, float(np.abs(r).max()), float(ld.numpy())))
# CPU: the oracle restatement (same SciPy-call-per-step structure as the reference) on a short chain
from oracle.bpk_ref import RefBackend    # noqa: E402
rb = RefBackend()
Ts = 2000
hA, hB, hy = np.ascontiguousarray(A[:Ts]), np.ascontiguousarray(B[:Ts - 1]), np.ascontiguousarray(y[:Ts])
hV, hC, hx, hl = np.empty((Ts, Dm, Dm)), np.empty((Ts - 1, Dm, Dm)), np.empty((Ts, Dm)), np.empty(1)
t0 = time.perf_counter()
rb.block_banded_solve(hA.ctypes.data, hB.ctypes.data, hy.ctypes.data, 1, Ts, Dm, hV.ctypes.data, hC.ctypes.data,
                      hx.ctypes.data, hl.ctypes.data)
dt = time.perf_counter() - t0
print("oracle (NumPy/SciPy per-step loop) T=%d: %.1f us/step -> %.1f s at T=%d (extrapolated); GPU speed-up %.0fx"
      % (Ts, 1e6 * dt / Ts, dt / Ts * T, T, (dt / Ts * T) / (ms * 1e-3)))
