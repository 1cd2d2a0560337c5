#!/bin/bash
# GPU session 11 (1 GPU): where the LSSM iteration's wall time goes.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s11; mkdir -p $O
timeout 400 python tools/lssm_node_timing.py > $O/lssm_nodes.txt 2>&1
BPK_GMC_BCR_V2=1 BPK_EWISE_GENERIC=1 timeout 400 python tools/lssm_node_timing.py > $O/lssm_nodes_old_kernels.txt 2>&1
echo finished > $O/done.txt
