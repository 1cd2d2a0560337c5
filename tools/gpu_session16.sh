#!/bin/bash
# GPU session 16 (1 GPU): whole suite after the widening work, smoke, headline bench with host timing.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s16; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/pytest.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5 ) > $O/smoke.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
echo finished > $O/done.txt
