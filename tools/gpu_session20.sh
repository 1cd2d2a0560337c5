#!/bin/bash
# GPU session 20 (1 GPU, the last ~2 minutes of the round's budget): smoke() and the N = 1e5 reference-trajectory test
# after the host-side changes of the last sessions; the kernel-limit skip of the reference's rotation gradient test.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/s20
( timeout 18 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/s20/smoke.txt
( timeout 24 python -m pytest tests/test_fullsize.py tests/test_reference_node_tests.py -m gpu -q -k "1e5 or cost_gradient or sharding" -p no:cacheprovider 2>&1 | tail -6 ) > gpurun_out/s20/pytest.txt
