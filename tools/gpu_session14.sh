#!/bin/bash
# GPU session 14 (1 GPU): chain rotation + user-guide example on the device, LSSM bench with the sampler started before warm-up.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s14; mkdir -p $O
( timeout 900 python -m pytest tests/test_rotation.py tests/test_doc_examples.py -m gpu -q 2>&1 | tail -30 ) > $O/pytest_first.txt
BPK_TRACE_SLOW=1 timeout 600 python bench.py --workload lssm --steps 10 --warmup 4 --e2e-steps 2 > $O/bench_lssm.json 2> $O/bench_lssm.err
echo finished > $O/done.txt
