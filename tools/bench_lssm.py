"""Config 3 of BASELINE.json: linear state-space model (lssm.rst:45-181 scaled up) — seconds per VB iteration.
    python tools/bench_lssm.py [T] [D] [M] [iters]
X = GaussianMarkovChain(0, 1e-3 I, A, 1, n=T), A = GaussianARD(0, alpha) (D x D), C = GaussianARD(0, gamma) plates (M,1),
F = Dot(C, X), Y = GaussianARD(F, tau) observed."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                                            # noqa: E402
from bayespy_b200 import _bpk                                                 # noqa: E402
from bayespy_b200.nodes import GaussianARD, GaussianMarkovChain, Gamma, Dot   # noqa: E402
from bayespy_b200.inference import VB                                         # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
Dm = int(sys.argv[2]) if len(sys.argv) > 2 else 32
M = int(sys.argv[3]) if len(sys.argv) > 3 else 256
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
rs = np.random.RandomState(0)
# data: 4 latent signals (two noisy rotators), lssm.rst:153-175 pattern
w = 0.05
a = np.array([[np.cos(w), -np.sin(w), 0, 0], [np.sin(w), np.cos(w), 0, 0], [0, 0, np.cos(3 * w), -np.sin(3 * w)],
              [0, 0, np.sin(3 * w), np.cos(3 * w)]])
x = np.empty((T, 4))
x[0] = rs.randn(4)
noise = 0.1 * rs.randn(T, 4)
for n in range(T - 1):
    x[n + 1] = a @ x[n] + noise[n]
c = rs.randn(M, 4)
y = c @ x.T + 3.0 * rs.randn(M, T)

alpha = Gamma(1e-5, 1e-5, plates=(Dm,), name="alpha")
A = GaussianARD(0, alpha, shape=(Dm,), plates=(Dm,), name="A")
X = GaussianMarkovChain(np.zeros(Dm), 1e-3 * np.identity(Dm), A, np.ones(Dm), n=T, name="X")
gamma = Gamma(1e-5, 1e-5, plates=(Dm,), name="gamma")
C = GaussianARD(0, gamma, shape=(Dm,), plates=(M, 1), name="C")
F = Dot(C, X, name="F")
C.initialize_from_value(np.random.RandomState(1).randn(M, 1, Dm))
tau = Gamma(1e-5, 1e-5, name="tau")
Y = GaussianARD(F, tau, name="Y")
Y.observe(y)
Q = VB(X, C, gamma, A, alpha, tau, Y)
be = _bpk.get()
Q.update(repeat=3, verbose=False, tol=0)     # warm-up: lets the stream-ordered memory pool reach its steady size
be.sync()
l0 = be.launch_count()
t0 = time.perf_counter()
Q.update(repeat=iters, verbose=False, tol=0)
be.sync()
dt = (time.perf_counter() - t0) / iters
print("LSSM T=%d D=%d M=%d: %.3f s per VB iteration (%d kernel launches per iteration), L = %s"
      % (T, Dm, M, dt, (be.launch_count() - l0) // iters, np.array2string(Q.L[:iters + 1], precision=6)))
