"""Timing of the fused masked factor-model sweep (bpk_pca_xsweep_masked_fused) through the C-ABI:
    python tools/bench_masked.py [N]      (BPK_PMASK_CHUNK_TILES sets the chunk: tiles of 128 columns per SM)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                      # noqa: E402
from bayespy_b200 import _bpk           # noqa: E402
from bayespy_b200.darray import DArray  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
M, K = 64, 16
be = _bpk.get()
rng = np.random.default_rng(0)
Y = DArray.empty((M, N))
mask = DArray.empty((M, N), "u1")
chunk = 1_000_000
for n0 in range(0, N, chunk):
    n1 = min(N, n0 + chunk)
    yb = rng.standard_normal((M, n1 - n0))
    mb = rng.random((M, n1 - n0)) < 0.8
    for m in range(M):
        be.h2d(Y.ptr + 8 * (m * N + n0), yb[m])
        be.h2d(mask.ptr + (m * N + n0), mb[m].astype(np.uint8))
W = rng.standard_normal((M, K))
Cw = 0.1 * rng.standard_normal((M, K, K))
WW = W[:, :, None] * W[:, None, :] + Cw @ np.swapaxes(Cw, -1, -2)
d = {k: DArray.from_numpy(v) for k, v in dict(W=W, WW=WW, alpha=np.ones(K), amu=np.zeros(K)).items()}
X, G = DArray.empty((N, K)), DArray.empty((N,))
st = DArray.zeros((M * K + M * K * K + K * K + K + 2,))


def run():
    be.pca_xsweep_masked_fused(Y.ptr, mask.ptr, M, N, K, d["W"].ptr, d["WW"].ptr, 1.3, d["alpha"].ptr, d["amu"].ptr,
                               X.ptr, G.ptr, st.ptr, False)


for _ in range(2):
    run()
ts = []
for _ in range(3):
    t = be.timer_create()
    be.timer_record(t, 0)
    l0 = be.launch_count()
    run()
    be.timer_record(t, 1)
    ts.append(be.timer_elapsed_ms(t))
    nl = be.launch_count() - l0
ms = float(np.median(ts))
print("pca_xsweep_masked_fused N=%d M=%d K=%d chunk_tiles=%s: %.3f ms (%d launches) | %.0f GB/s of 704 B/col | "
      "%.1f TFLOP/s on 76 DMMA/col executed (%.0f %% of the 36.9 TFLOP/s pipe; + ~3 kflop/col of DFMA) | %.2f us per 1000 columns"
      % (N, M, K, os.environ.get("BPK_PMASK_CHUNK_TILES", "8"), ms, nl, 704 * N / ms / 1e6, 76 * 512 * N / ms / 1e9,
         100 * 76 * 512 * N / ms / 1e9 / 36.9, ms * 1e3 / (N / 1000)))
sv = st.numpy()
print("  checksums: sum(stats) = %.12e, sum|stats| = %.12e, X[::100003] . 1 = %.12e"
      % (float(np.sum(sv)), float(np.sum(np.abs(sv))), float(np.sum(X.numpy()[::100003]))))
