#!/bin/bash
# GPU session 1 (1 GPU): tests, bench, A/B of the tail variants, tail stamps at the 8-GPU per-rank size.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s1; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
for v in "BPK_VB_NO_LL=1" "BPK_VB_GJ1=1" "BPK_VB_DRY_FIRST_ONLY=1"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_$v.json 2> $O/bench_$v.err
done
# per-rank size of the 8-GPU strong-scaling run
for v in "X=1" "BPK_VB_NO_LL=1" "BPK_VB_GJ1=1"; do
  env $v timeout 300 python bench.py --n 1250000 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_1250k_$v.json 2> $O/bench_1250k_$v.err
  env $v BPK_VB_DEBUG=1 timeout 300 python tools/vb_tail_timing.py 1250000 > $O/tail_1250k_$v.txt 2>&1
done
BPK_VB_DEBUG=1 timeout 300 python tools/vb_tail_timing.py 10000000 > $O/tail_10m.txt 2>&1
timeout 600 python bench.py --workload gmm --steps 10 --warmup 3 > $O/bench_gmm.json 2> $O/bench_gmm.err
BPK_GMM_V1=1 timeout 600 python bench.py --workload gmm --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/bench_gmm_v1.json 2> $O/bench_gmm_v1.err
timeout 300 python tools/bench_gmm.py > $O/gmm_kernel_v2.txt 2>&1
BPK_GMM_V1=1 timeout 300 python tools/bench_gmm.py > $O/gmm_kernel_v1.txt 2>&1
timeout 600 python tools/bench_gemm.py > $O/gemm.txt 2>&1
timeout 900 python bench.py --workload lssm --steps 5 --warmup 3 > $O/bench_lssm.json 2> $O/bench_lssm.err
echo finished > $O/done.txt
