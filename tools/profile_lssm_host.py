"""Host-side cProfile of one LSSM VB iteration on the GPU backend (where does the wall time go?)."""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["x"] + (sys.argv[1:] or ["100000", "32", "256", "1"])
pr = cProfile.Profile()
src = open(os.path.join(ROOT, "tools", "bench_lssm.py")).read().replace(
    "Q.update(repeat=iters, verbose=False, tol=0)\nbe.sync()\ndt",
    "pr.enable(); Q.update(repeat=iters, verbose=False, tol=0); be.sync(); pr.disable()\ndt")
exec(src)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
print(s.getvalue()[-3200:])
