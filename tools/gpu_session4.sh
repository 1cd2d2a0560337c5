#!/bin/bash
# GPU session 4 (1 GPU): persistent loop kernel v3 (service CTA + dynamic tiles), masked v2, GMM v3, split-K, full test suite, ncu.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s4; mkdir -p $O
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_v3.json 2> $O/bench_v3.err
BPK_PCA_STATIC=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_static.json 2> $O/bench_static.err
timeout 300 python bench.py --columns 1250000 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_1250k_v3.json 2> $O/bench_1250k_v3.err
BPK_PCA_STATIC=1 timeout 300 python bench.py --columns 1250000 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_1250k_static.json 2> $O/bench_1250k_static.err
BPK_VB_DEBUG=1 TAIL_SWEEPS=20 timeout 300 python tools/vb_tail_timing.py 1250000 > $O/probe_1250k_v3.txt 2>&1
BPK_VB_DEBUG=1 TAIL_SWEEPS=20 timeout 300 python tools/vb_tail_timing.py 10000000 > $O/probe_10m_v3.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > $O/pytest.txt
timeout 300 python tools/bench_masked.py 10000000 > $O/masked_v2.txt 2>&1
BPK_PMASK_INV_REGS=128 timeout 300 python tools/bench_masked.py 10000000 > $O/masked_v2_inv128.txt 2>&1
BPK_PMASK_CHUNK_TILES=8 timeout 300 python tools/bench_masked.py 10000000 > $O/masked_v2_chunk8.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pmask -c 12 --csv --log-file $O/pmask_times.csv python tools/bench_masked.py 1000000 > $O/ncu_pmask.log 2>&1
timeout 300 python tools/bench_gmm.py > $O/gmm_kernel_v2.txt 2>&1
BPK_GMM_V3=1 timeout 300 python tools/bench_gmm.py > $O/gmm_kernel_v3.txt 2>&1
timeout 600 python tools/bench_gemm.py > $O/gemm.txt 2>&1
timeout 900 python bench.py --workload lssm --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_lssm.json 2> $O/bench_lssm.err
timeout 900 python bench.py --workload pca_masked --steps 5 --warmup 3 > $O/bench_masked.json 2> $O/bench_masked.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pca_vbloop -s 1 -c 1 -o $O/pca_loop_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu_loop.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_pca.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu_launches.log 2>&1
echo finished > $O/done.txt
