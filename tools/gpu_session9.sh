#!/bin/bash
# GPU session 9 (8 GPUs, short): aligned start instant, service CTA without a reduction share, resident mixture loop over 8 ranks.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s9; mkdir -p $O
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
run 8 bench.py --gpus 8 --steps 20 --warmup 5 --e2e-steps 1 > $O/bench_8gpu.json 2> $O/bench_8gpu.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-steps 1 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err
BPK_VB_DEBUG=1 TAIL_SWEEPS=20 run 8 tools/vb_tail_timing.py 1250000 > $O/probe_8gpu.txt 2>&1
run 8 tests/dist_gpu_check.py > $O/dist_check_8.txt 2>&1
run 8 bench.py --workload gmm --gpus 8 --steps 20 --warmup 5 --e2e-steps 1 > $O/bench_gmm_8gpu.json 2> $O/bench_gmm_8gpu.err
run 8 bench.py --gpus 8 --steps 20 --warmup 5 --e2e-steps 1 > $O/bench_8gpu_b.json 2> $O/bench_8gpu_b.err
echo finished > $O/done.txt
