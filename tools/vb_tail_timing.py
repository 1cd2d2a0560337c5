"""Where does the tail of the fused sweep launch spend its time?  (BPK_VB_DEBUG stamps)
    BPK_VB_DEBUG=1 python tools/vb_tail_timing.py [N]
"""
import os
import sys

os.environ["BPK_VB_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np          # noqa: E402
import bench                # noqa: E402
from bayespy_b200 import _bpk   # noqa: E402

from bayespy_b200 import parallel  # noqa: E402
world, rank = parallel.init_from_env()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000      # columns PER RANK
y = bench.synth_shard(64, rank * N, (rank + 1) * N, 1)
Q, nodes = bench.build_model(y)
Q.update(repeat=int(os.environ.get("TAIL_SWEEPS", "5")), verbose=False)
be = _bpk.get()
s = be.debug_stamps(192)
parallel.barrier()
if rank != 0:
    sys.exit(0)
t0 = s[0]
names = {1: "data pass done (CTA 0)", 2: "after grid barrier 1", 3: "CTA 0 reduction share / push done",
         4: "tail starts (after grid barrier 2, if any)", 5: "tail done"}
print("N=%d: stamps relative to kernel start (us)" % N)
for i in (1, 2, 3, 4):
    print("  %-32s %10.2f" % (names[i], (s[i] - t0) / 1e3))
opn = {2: "STATS", 3: "SXXT", 4: "XPRE", 5: "ROW", 6: "ALPHA", 7: "TAU", 8: "BOUND"}
plan = Q.plans[0]
ops = [o for o in plan.resident_program(Q, Q.model)[0] if o != 1]
# tail of one sweep = ops after XSWEEP of iteration i + ops before XSWEEP of iteration i+1
prog = plan.resident_program(Q, Q.model)[0]
k = prog.index(1)
tail = prog[k + 1:] + prog[:k]
prev = s[4]
for i, op in enumerate(tail):
    st = s[8 + i]
    nxt = s[8 + i + 1] if i + 1 < len(tail) else s[5]
    print("  op %-6s %10.2f us" % (opn[op], (nxt - st) / 1e3))
if s[6] > s[4]:
    print("  dry run (cold) before the real ops %8.2f us ; real ops (warm) %8.2f us" % ((s[6] - s[4]) / 1e3, (s[5] - s[6]) / 1e3))
if s[40] and s[42]:
    print("  p2p windows open: %s ; exchange inside STATS: push (single-CTA path only) %.2f us, gather %.2f us"
          % (parallel._state.get("p2p"), (s[41] - s[40]) / 1e3, (s[42] - s[41]) / 1e3))
if s[48] and s[56]:
    # probe CTA (last of the grid), second-to-last sweep of the launch
    b = s[48]
    lab = [(49, "A fragments loaded (X-warp 0)"), (50, "first tile landed (X-warp 0)"), (51, "last tile done (X-warp 0)"),
           (52, "last tile done (S-warp 0)"), (53, "CTA partial written"), (54, "after grid barrier 1"),
           (55, "slice reduced + pushed"), (56, "after grid barrier 3 (next sweep may start)"), (57, "next sweep starts")]
    print("  probe CTA %d, one steady-state sweep (us from its start):" % (-1))
    for i, nm in lab:
        if s[i]:
            print("    %-44s %10.2f" % (nm, (s[i] - b) / 1e3))
print("  %-32s %10.2f" % (names[5], (s[5] - t0) / 1e3))
print("  tail total (after data pass)     %10.2f" % ((s[5] - s[1]) / 1e3))

sw = [v for v in s[64:192] if v]
if len(sw) > 2:
    d = [(sw[0] - t0) / 1e3] + [(b - a) / 1e3 for a, b in zip(sw, sw[1:])]
    print("  sweeps of the last launch, service CTA passing grid barrier 1 (us since the previous one; first: since kernel start):")
    print("    " + " ".join("%.1f" % v for v in d))
