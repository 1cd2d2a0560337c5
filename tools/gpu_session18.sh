#!/bin/bash
# GPU session 18 (1 GPU, short): the nodes added in this session (GaussianGamma, chain inputs, plated Varying chains,
# multi-axis rotation, pairwise contraction fallback) on libbpk, then the neighbouring suites as a regression check.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s18; mkdir -p $O
( timeout 200 python -m pytest tests/test_gaussian_gamma.py tests/test_gmc.py tests/test_reference_node_tests.py -m gpu -q --durations=8 2>&1 | tail -60 ) > $O/pytest_new.txt
( timeout 150 python -m pytest tests/test_dot_node.py tests/test_models.py tests/test_rotation.py tests/test_known_answers.py tests/test_gate.py tests/test_take.py tests/test_slice.py -m gpu -q --durations=5 2>&1 | tail -40 ) > $O/pytest_regress.txt
echo finished > $O/done.txt
