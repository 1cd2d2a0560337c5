#!/bin/bash
# GPU session 17 (1 GPU): last changes (masked chunk size, chain distribution choice, program cache) + ncu of the smoother kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s17; mkdir -p $O
( timeout 900 python -m pytest tests/test_sweeps.py tests/test_gmc.py tests/test_resident.py tests/test_models.py tests/test_reference_node_tests.py -m gpu -q 2>&1 | tail -20 ) > $O/pytest.txt
timeout 300 python tools/bench_gmc.py > $O/bench_gmc.txt 2>&1
timeout 600 python bench.py --workload pca_masked --steps 10 --warmup 3 --e2e-steps 2 --no-cpu-baseline > $O/bench_masked.json 2> $O/bench_masked.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gmc_bcr3 -c 3 -o $O/bcr3 python tools/bench_gmc.py > $O/ncu_bcr3.log 2>&1
ncu -i $O/bcr3.ncu-rep --page raw --csv > $O/bcr3_raw.csv 2>/dev/null
echo finished > $O/done.txt
