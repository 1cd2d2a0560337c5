#!/bin/bash
# GPU session 15 (1 GPU): gradient learning / mixing chains on the device; host time around the resident launches.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s15; mkdir -p $O
( timeout 900 python -m pytest tests/test_gradients.py tests/test_gmc.py tests/test_resident.py -m gpu -q 2>&1 | tail -30 ) > $O/pytest_first.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --columns 1250000 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_1250k.json 2> $O/bench_1250k.err
echo finished > $O/done.txt
