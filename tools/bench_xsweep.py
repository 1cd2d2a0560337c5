"""A/B timing of the PCA sweep kernel variants (BPK_PCA_VARIANT) through the C-ABI.
    python tools/bench_xsweep.py [N]        # one process per variant (the choice is read once)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(N, reps=10):
    import numpy as np
    from bayespy_b200 import _bpk
    from bayespy_b200.darray import DArray
    be = _bpk.get()
    M, K = 64, 16
    rng = np.random.default_rng(0)
    Y = DArray.empty((M, N))
    chunk = 1_000_000
    for n0 in range(0, N, chunk):                      # fill on device-sized chunks
        n1 = min(N, n0 + chunk)
        blk = rng.standard_normal((M, n1 - n0))
        for m in range(M):
            be.h2d(Y.ptr + 8 * (m * N + n0), blk[m])
    A = DArray.from_numpy(rng.standard_normal((K, M)) / 8)
    b = DArray.from_numpy(rng.standard_normal(K))
    X = DArray.empty((N, K))
    st = DArray.zeros((M * K + K * K + K,))
    for _ in range(3):
        be.pca_xsweep(Y.ptr, M, N, K, A.ptr, b.ptr, X.ptr, st.ptr)
    ts = []
    for _ in range(reps):
        t = be.timer_create()
        be.timer_record(t, 0)
        be.pca_xsweep(Y.ptr, M, N, K, A.ptr, b.ptr, X.ptr, st.ptr)
        be.timer_record(t, 1)
        ts.append(be.timer_elapsed_ms(t))
    ms = float(np.median(ts))
    s = st.numpy()
    print("variant %s  N=%d  %.3f ms (incl. %d-thread final reduce)  %.0f GB/s  %.1f TFLOP/s  checksum %.6e"
          % (os.environ.get("BPK_PCA_VARIANT", "0"), N, ms, 128, 640 * N / ms / 1e6, 4864 * N / ms / 1e9,
             float(np.sum(s)) / (reps + 3)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "--one":
        one(int(sys.argv[1]))
    else:
        N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
        for v in ("-1", "0", "1", "2", "3"):
            env = dict(os.environ, BPK_PCA_VARIANT=v)
            subprocess.run([sys.executable, os.path.abspath(__file__), str(N), "--one"], env=env)
