"""Where does one VB iteration of the LSSM workload (T = 1e5, D = 32, M = 256) spend its wall time?
Per node: update() + stream synchronisation, launches; then the lower bound.   python tools/lssm_node_timing.py [T]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np                      # noqa: E402
import bench_workloads as bw            # noqa: E402
from bayespy_b200 import _bpk           # noqa: E402
from bayespy_b200.nodes import GaussianARD, GaussianMarkovChain, Gamma, Dot   # noqa: E402
from bayespy_b200.inference import VB   # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
be = _bpk.get()
y = bw.lssm_data(T, 256)
Q, Y = bw.build_lssm(y, 32, (GaussianARD, GaussianMarkovChain, Gamma, Dot, VB))
Q.update(repeat=3, verbose=False)
be.sync()
for it in range(3):
    line = []
    t_it = time.perf_counter()
    for name in ("X", "C", "gamma", "A", "alpha", "tau"):
        node = Q[name]
        l0 = be.launch_count()
        t0 = time.perf_counter()
        node.update()
        t1 = time.perf_counter()
        be.sync()
        t2 = time.perf_counter()
        line.append("%s %.1f+%.1f ms (%d)" % (name, 1e3 * (t1 - t0), 1e3 * (t2 - t1), be.launch_count() - l0))
    l0 = be.launch_count()
    t0 = time.perf_counter()
    L = Q.compute_lowerbound()
    be.sync()
    t1 = time.perf_counter()
    line.append("bound %.1f ms (%d)" % (1e3 * (t1 - t0), be.launch_count() - l0))
    print("iteration %d: %.1f ms | " % (it, 1e3 * (time.perf_counter() - t_it)) + " | ".join(line), flush=True)
t0 = time.perf_counter()
Q.update(repeat=3, verbose=False)
be.sync()
print("Q.update(repeat=3): %.1f ms per iteration" % (1e3 * (time.perf_counter() - t0) / 3))
