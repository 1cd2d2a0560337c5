"""Throughput of the GEMM path of bpk_sum_multiply (dgemm_dmma_kernel) on the contraction shapes of the
state-space model (dot.py:403,581): python tools/bench_gemm.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                      # noqa: E402
from bayespy_b200 import _bpk           # noqa: E402
from bayespy_b200 import darray as D    # noqa: E402
from bayespy_b200.darray import DArray  # noqa: E402

be = _bpk.get()
rng = np.random.default_rng(0)
cases = [("<ff> = cc[m,ij] xx[n,ij]", (256, 1024), ["m", "k"], (100_000, 1024), ["n", "k"], ["m", "n"]),
         ("msg  = y[m,n] x[n,k]", (256, 100_000), ["m", "k"], (100_000, 1024), ["k", "n"], ["m", "n"]),
         ("square 4096^3", (4096, 4096), ["m", "k"], (4096, 4096), ["k", "n"], ["m", "n"]),
         ("tall  = a[n,i] b[i,j]", (1_000_000, 64), ["m", "k"], (64, 64), ["k", "n"], ["m", "n"])]
for name, sa, ka, sb, kb, ko in cases:
    A = DArray.from_numpy(rng.standard_normal(sa))
    B = DArray.from_numpy(rng.standard_normal(sb))
    ext = dict(zip(ka, sa)); ext.update(dict(zip(kb, sb)))
    out = DArray.empty(tuple(ext[k] for k in ko))
    for _ in range(2):
        D.sum_product([A, B], [ka, kb], ko, out=out)
    ts = []
    for _ in range(5):
        t = be.timer_create()
        be.timer_record(t, 0)
        D.sum_product([A, B], [ka, kb], ko, out=out)
        be.timer_record(t, 1)
        ts.append(be.timer_elapsed_ms(t))
    ms = float(np.median(ts))
    flops = 2.0 * ext["m"] * ext["n"] * ext["k"]
    # spot check against NumPy on a corner
    ref = np.einsum("%s,%s->%s" % ("".join(ka), "".join(kb), "".join(ko)), A.numpy()[:8] if ka[0] == "m" else A.numpy(),
                    B.numpy())[:8, :8] if ext["m"] * ext["n"] * ext["k"] < 3e11 else None
    err = float(np.max(np.abs(out.numpy()[:8, :8] - ref))) if ref is not None else float("nan")
    print("%-28s M=%d N=%d K=%d: %8.3f ms  %6.2f TFLOP/s (%.1f %% of the 36.9 TFLOP/s DMMA peak)  max|err| %.2e"
          % (name, ext["m"], ext["n"], ext["k"], ms, flops / ms / 1e9, 100 * flops / ms / 1e9 / 36.9, err))
