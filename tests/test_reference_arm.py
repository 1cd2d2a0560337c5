"""The staged reference package (oracle/_ref, recipe oracle/make_ref.py) is the unmodified reference: it reproduces
its own published doctest, and bench.py's reference arm runs it (kind "reference")."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden


def _ref():
    from oracle import make_ref
    make_ref.build()
    if not make_ref.available():
        pytest.skip("oracle/_ref is not staged and /root/reference is absent")
    return make_ref


def test_staged_reference_reproduces_the_quickstart_doctest():
    """doc/source/user_guide/quickstart.rst:113-118 through oracle/_ref, in a clean interpreter."""
    _ref()
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle import make_ref; make_ref.import_reference()\n"
        "import numpy as np\n"
        "from bayespy.nodes import GaussianARD, Gamma\n"
        "from bayespy.inference import VB\n"
        "np.random.seed(1)\n"
        "data = np.random.normal(5, 10, size=(10,))\n"
        "mu = GaussianARD(0, 1e-6); tau = Gamma(1e-6, 1e-6)\n"
        "y = GaussianARD(mu, tau, plates=(10,)); y.observe(data)\n"
        "Q = VB(mu, tau, y); Q.update(repeat=20)\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith("Iteration 1: loglike=-6.020956e+01")
    assert lines[3].startswith("Iteration 4: loglike=-5.820288e+01")
    assert lines[-1] == "Converged at iteration 4."


def test_staged_reference_matches_the_committed_golden():
    """Same PCA run as tests/golden/pca_small.npz, through oracle/ref_models (the bench's model builder)."""
    _ref()
    from oracle import ref_models
    g = golden("pca_small")
    Q, n = ref_models.pca(g["y"], g["C_init"].shape[-1], g["C_init"])
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-12)
    np.testing.assert_allclose(n["tau"].u[0], g["tau_u0"], rtol=1e-12)


def test_bench_reference_arm_runs_the_reference():
    _ref()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-budget-s", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference"
    assert line["unit"] == "it/s" and line["value"] > 0 and line["config"]["value_is_extrapolated"] is True
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
