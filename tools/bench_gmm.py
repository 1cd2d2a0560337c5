"""Timing of the fused GMM sweep kernel (bpk_gmm_sweep) through the C-ABI: config 2 of BASELINE.json
(N=1e7, D=8, K=64).   python tools/bench_gmm.py [N]     (BPK_GMM_V0=1 selects the scalar v0 kernel)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                     # noqa: E402
from bayespy_b200 import _bpk          # noqa: E402
from bayespy_b200.darray import DArray  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
D, K = 8, 64
be = _bpk.get()
rng = np.random.default_rng(0)
means = 5 * rng.standard_normal((K, D))
Y = DArray.empty((N, D))
chunk = 2_000_000
for n0 in range(0, N, chunk):
    n1 = min(N, n0 + chunk)
    z = rng.integers(0, K, size=n1 - n0)
    blk = means[z] + rng.standard_normal((n1 - n0, D))
    be.h2d(Y.ptr + 8 * n0 * D, blk)
R = rng.standard_normal((K, D, D)) * 0.1
Lam = R @ np.swapaxes(R, -1, -2) + np.identity(D)
h = np.einsum("kij,kj->ki", Lam, means)
c = -0.5 * np.einsum("ki,ki->k", h, means) + 0.5 * np.linalg.slogdet(Lam)[1]
logpi = np.log(np.full(K, 1.0 / K))
d = {k: DArray.from_numpy(v) for k, v in dict(c=c, h=h, Lam=Lam, logpi=logpi).items()}
P, G = DArray.empty((N, K)), DArray.empty((N,))
st = DArray.zeros((K + K * D + K * D * D + 1,))


def run():
    be.gmm_sweep(Y.ptr, N, D, K, d["c"].ptr, d["h"].ptr, d["Lam"].ptr, d["logpi"].ptr, P.ptr, G.ptr, st.ptr)


for _ in range(2):
    run()
ts = []
for _ in range(5):
    t = be.timer_create()
    be.timer_record(t, 0)
    run()
    be.timer_record(t, 1)
    ts.append(be.timer_elapsed_ms(t))
ms = float(np.median(ts))
s = st.numpy()
print("gmm_sweep %s N=%d D=%d K=%d: %.3f ms  | %.0f GB/s of 576 B/row | %.1f TFLOP/s algorithmic (18.5 kflop/row), "
      "%.1f TFLOP/s executed (24 DMMA + 64 exp per row) | sum R = %.6e (expect %.6e per sweep)"
      % ("v0" if os.environ.get("BPK_GMM_V0") else "dmma", N, D, K, ms, 576 * N / ms / 1e6, 18.5e3 * N / ms / 1e9,
         12.3e3 * N / ms / 1e9, float(np.sum(s[:K])) / 7, float(N)))
