#!/bin/bash
# GPU session 7 (8 GPUs, one box): strong scaling of the headline metric at 1/2/4/8 ranks, config 5 (N = 1e8), GMM, parity.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s7; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
run 8 bench.py --gpus 8 --steps 20 --warmup 5 --e2e-steps 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err
run 8 bench.py --gpus 8 --steps 200 --warmup 5 --e2e-steps 1 > $O/bench_8gpu_200.json 2> $O/bench_8gpu_200.err
BPK_VB_DEBUG=1 TAIL_SWEEPS=20 run 8 tools/vb_tail_timing.py 1250000 > $O/probe_8gpu.txt 2>&1
run 4 bench.py --gpus 4 --steps 20 --warmup 5 --e2e-steps 3 > $O/bench_4gpu.json 2> $O/bench_4gpu.err
run 2 bench.py --gpus 2 --steps 20 --warmup 5 --e2e-steps 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-steps 3 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err
run 8 tests/dist_gpu_check.py > $O/dist_check_8.txt 2>&1
run 8 bench.py --gpus 8 --columns 100000000 --steps 10 --warmup 3 --e2e-steps 1 > $O/bench_1e8_8gpu.json 2> $O/bench_1e8_8gpu.err
run 4 bench.py --gpus 4 --columns 100000000 --steps 10 --warmup 3 --e2e-steps 1 > $O/bench_1e8_4gpu.json 2> $O/bench_1e8_4gpu.err
run 8 bench.py --workload gmm --gpus 8 --steps 10 --warmup 3 --e2e-steps 2 > $O/bench_gmm_8gpu.json 2> $O/bench_gmm_8gpu.err
BPK_PCA_STATIC=1 run 8 bench.py --gpus 8 --steps 200 --warmup 5 --e2e-steps 1 > $O/bench_8gpu_static.json 2> $O/bench_8gpu_static.err
echo finished > $O/done.txt
