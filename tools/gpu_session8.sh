#!/bin/bash
# GPU session 8 (1 GPU): masked GEMM kernels v3, resident mixture loop, gate/take; full suite; benches.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s8; mkdir -p $O
( timeout 900 python -m pytest tests/test_resident.py tests/test_sweeps.py tests/test_models.py tests/test_gate.py -m gpu -q -x 2>&1 | tail -30 ) > $O/pytest_first.txt
timeout 300 python tools/bench_masked.py > $O/masked_v3.txt 2>&1
BPK_PMASK_V2=1 timeout 300 python tools/bench_masked.py > $O/masked_v2.txt 2>&1
timeout 600 python bench.py --workload gmm --steps 20 --warmup 5 --e2e-steps 3 > $O/bench_gmm.json 2> $O/bench_gmm.err
timeout 600 python bench.py --workload pca_masked --steps 10 --warmup 3 --e2e-steps 2 > $O/bench_masked.json 2> $O/bench_masked.err
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > $O/pytest.txt
BPK_VB_DEBUG=1 TAIL_SWEEPS=20 timeout 300 python tools/vb_tail_timing.py 1250000 > $O/probe_1250k.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --e2e-steps 3 > $O/bench_default.json 2> $O/bench_default.err
echo finished > $O/done.txt
