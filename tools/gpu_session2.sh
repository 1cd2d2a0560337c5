#!/bin/bash
# GPU session 2 (1 GPU): masked fused sweep (tests + timing), probe stamps of the fused PCA sweep, GMM repeatability, ncu.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s2; mkdir -p $O
( timeout 900 python -m pytest tests/test_sweeps.py tests/test_models.py tests/test_fullsize.py tests/test_gmc.py tests/test_reference_seams.py -m gpu -q 2>&1 | tail -40 ) > $O/pytest.txt
for n in 1000000 10000000; do
  timeout 300 python tools/bench_masked.py $n > $O/masked_$n.txt 2>&1
done
BPK_PMASK_CHUNK_TILES=1 timeout 300 python tools/bench_masked.py 10000000 > $O/masked_1e7_chunk1.txt 2>&1
BPK_PMASK_CHUNK_TILES=4 timeout 300 python tools/bench_masked.py 10000000 > $O/masked_1e7_chunk4.txt 2>&1
timeout 900 python bench.py --workload pca_masked --n 10000000 --steps 5 --warmup 3 > $O/bench_masked.json 2> $O/bench_masked.err
# probe stamps at the 8-GPU per-rank size
for v in "X=1" "BPK_VB_DRY_FIRST_ONLY=1" "BPK_PCA_VARIANT=1" "BPK_PCA_VARIANT=3"; do
  env $v BPK_VB_DEBUG=1 TAIL_SWEEPS=20 timeout 300 python tools/vb_tail_timing.py 1250000 > "$O/probe_1250k_$v.txt" 2>&1
done
BPK_VB_DEBUG=1 TAIL_SWEEPS=20 timeout 300 python tools/vb_tail_timing.py 10000000 > $O/probe_10m.txt 2>&1
# GMM repeatability (v2 / v1 / v2)
for i in 1 2; do
  timeout 600 python bench.py --workload gmm --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/bench_gmm_v2_$i.json 2> $O/bench_gmm_v2_$i.err
  BPK_GMM_V1=1 timeout 600 python bench.py --workload gmm --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $O/bench_gmm_v1_$i.json 2> $O/bench_gmm_v1_$i.err
done
# ncu: launch list of the default bench, then --set full on the shipped fused sweep kernel (one sweep per launch)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_pca.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu_launches.log 2>&1
BPK_VB_NO_LOOP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pca_xsweep_ws_kernel -s 3 -c 1 -o $O/pca_fused_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:pmask -c 3 -o $O/pmask_full python tools/bench_masked.py 1000000 > $O/ncu_pmask.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gmm_sweep_dmma2 -s 1 -c 1 -o $O/gmm_full python tools/bench_gmm.py 10000000 > $O/ncu_gmm.log 2>&1
echo finished > $O/done.txt
