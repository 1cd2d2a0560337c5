#!/bin/bash
# GPU session 3 (2 GPUs): in-kernel LL exchange across ranks: parity, bench, stamps.  Plus the 1-GPU probe of one sweep.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s3; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 900 python -m pytest tests/test_fullsize.py -m gpu -q -k two_ranks 2>&1 | tail -30 ) > $O/pytest_2rank.txt
BPK_VB_DEBUG=1 TAIL_SWEEPS=20 timeout 300 python tools/vb_tail_timing.py 1250000 > $O/probe_1gpu_1250k.txt 2>&1
BPK_VB_DEBUG=1 TAIL_SWEEPS=20 timeout 300 $TR --master-port 29551 tools/vb_tail_timing.py 1250000 > $O/probe_2gpu_1250k.txt 2>&1
timeout 600 $TR --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
timeout 600 $TR --master-port 29553 bench.py --gpus 2 --n 2500000 --steps 200 --warmup 5 --e2e-steps 1 > $O/bench_2gpu_2500k.json 2> $O/bench_2gpu_2500k.err
BPK_NO_P2P=1 timeout 600 $TR --master-port 29554 bench.py --gpus 2 --n 2500000 --steps 200 --warmup 5 --e2e-steps 1 > $O/bench_2gpu_2500k_nccl.json 2> $O/bench_2gpu_2500k_nccl.err
timeout 600 python bench.py --n 1250000 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $O/bench_1gpu_1250k.json 2> $O/bench_1gpu_1250k.err
timeout 600 $TR --master-port 29555 bench.py --workload gmm --gpus 2 --steps 10 --warmup 3 --e2e-steps 2 > $O/bench_gmm_2gpu.json 2> $O/bench_gmm_2gpu.err
timeout 600 $TR --master-port 29556 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --ref-budget-s 15 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err
echo finished > $O/done.txt
