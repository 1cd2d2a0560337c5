#!/bin/bash
# GPU session 19 (1 GPU, the last minutes of the round's budget): everything added since session 18 (general rotations,
# Multinomial, CategoricalMarkovChain / hmm.rst, the reference's demos and vmp tests) on libbpk, then the core suites.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s19; mkdir -p $O
( timeout 70 python -m pytest tests/test_reference_node_tests.py tests/test_reference_demos.py tests/test_gaussian_gamma.py tests/test_multinomial.py tests/test_rotation.py tests/test_doc_examples.py tests/test_gmc.py tests/test_checkpoint.py tests/test_gate.py tests/test_take.py tests/test_models.py -m gpu -q --durations=6 -p no:cacheprovider 2>&1 | tail -80 ) > $O/pytest_new.txt
( timeout 45 python -m pytest tests/test_resident.py tests/test_dot_node.py tests/test_kernels.py tests/test_known_answers.py tests/test_slice.py tests/test_gradients.py tests/test_reference_seams.py tests/test_spd16.py tests/test_sweeps.py -m gpu -q --durations=4 -p no:cacheprovider 2>&1 | tail -40 ) > $O/pytest_core.txt
echo finished > $O/done.txt
