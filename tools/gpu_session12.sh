#!/bin/bash
# GPU session 12 (1 GPU): block cyclic reduction on the tensor pipe + persistent workspace: parity, LSSM iteration.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s12; mkdir -p $O
( timeout 600 python -m pytest tests/test_gmc.py tests/test_doc_examples.py -m gpu -q -x 2>&1 | tail -30 ) > $O/pytest_first.txt
timeout 400 python tools/lssm_node_timing.py > $O/lssm_nodes.txt 2>&1
timeout 600 python bench.py --workload lssm --steps 10 --warmup 4 --e2e-steps 2 > $O/bench_lssm.json 2> $O/bench_lssm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_lssm.csv python bench.py --workload lssm --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > $O/ncu_lssm.log 2>&1
echo finished > $O/done.txt
