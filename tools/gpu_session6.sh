#!/bin/bash
# GPU session 6 (1 GPU): checks before the 8-GPU run: full suite, probe, bench at the per-rank size.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s6; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $O/pytest.txt
BPK_VB_DEBUG=1 TAIL_SWEEPS=20 timeout 300 python tools/vb_tail_timing.py 1250000 > $O/probe_1250k.txt 2>&1
timeout 300 python bench.py --columns 1250000 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_1250k.json 2> $O/bench_1250k.err
timeout 300 python bench.py --steps 20 --warmup 5 --e2e-steps 3 > $O/bench_default.json 2> $O/bench_default.err
echo finished > $O/done.txt
