#!/bin/bash
# GPU session 5 (1 GPU): loop kernel v3 with the timed dry run and the side-by-side tail ops: tests, probes, A/B.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s5; mkdir -p $O
( timeout 900 python -m pytest tests/test_resident.py tests/test_fullsize.py tests/test_models.py tests/test_rotation.py tests/test_doc_examples.py -m gpu -q 2>&1 | tail -30 ) > $O/pytest.txt
for v in "X=1" "BPK_VB_SERIAL=1" "BPK_VB_DRY_FIRST_ONLY=1" "BPK_PCA_STATIC=1"; do
  env $v BPK_VB_DEBUG=1 TAIL_SWEEPS=20 timeout 300 python tools/vb_tail_timing.py 1250000 > "$O/probe_1250k_$v.txt" 2>&1
  env $v timeout 300 python bench.py --columns 1250000 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 > "$O/bench_1250k_$v.json" 2> "$O/bench_1250k_$v.err"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_v3.json 2> $O/bench_v3.err
BPK_PCA_STATIC=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_static.json 2> $O/bench_static.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_v3_b.json 2> $O/bench_v3_b.err
echo finished > $O/done.txt
