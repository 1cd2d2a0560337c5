#!/bin/bash
# GPU session 13 (1 GPU): full suite, smoke, headline bench, LSSM bench with the allocation trace, GMM bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s13; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $O/pytest.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5 ) > $O/smoke.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
BPK_TRACE_SLOW=1 timeout 600 python bench.py --workload lssm --steps 10 --warmup 4 --e2e-steps 2 > $O/bench_lssm.json 2> $O/bench_lssm.err
timeout 600 python bench.py --workload gmm --steps 20 --warmup 5 --e2e-steps 3 --no-cpu-baseline > $O/bench_gmm.json 2> $O/bench_gmm.err
echo finished > $O/done.txt
