#!/bin/bash
# GPU session 10 (1 GPU): block cyclic reduction v3 (warp per node), 16-byte elementwise path: parity, LSSM iteration timing.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s10; mkdir -p $O
( timeout 900 python -m pytest tests/test_gmc.py tests/test_kernels.py -m gpu -q -x 2>&1 | tail -30 ) > $O/pytest_first.txt
timeout 600 python bench.py --workload lssm --steps 5 --warmup 4 --e2e-steps 2 > $O/bench_lssm.json 2> $O/bench_lssm.err
BPK_GMC_BCR_V2=1 timeout 600 python bench.py --workload lssm --steps 5 --warmup 4 --e2e-steps 2 --no-cpu-baseline > $O/bench_lssm_bcr2.json 2> $O/bench_lssm_bcr2.err
BPK_EWISE_GENERIC=1 timeout 600 python bench.py --workload lssm --steps 5 --warmup 4 --e2e-steps 2 --no-cpu-baseline > $O/bench_lssm_ewgeneric.json 2> $O/bench_lssm_ewgeneric.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_lssm.csv python bench.py --workload lssm --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > $O/ncu_lssm.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $O/pytest.txt
echo finished > $O/done.txt
