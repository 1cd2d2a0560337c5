"""The reference's demo scripts (bayespy/demos/*.py, staged unmodified in oracle/_ref) run AS THEY ARE through this
package: ``bayespy_b200.install_as_bayespy(stub_plotting=True)`` answers ``import bayespy...`` (and the plotting
modules with no-op stand-ins), the scripts build their models, simulate their data with their own seeds, run VB with
rotations / annealing / pattern search / stochastic updates / checkpoints, and every bound they print is compared with
what the reference printed for the same call (tests/golden/demos.npz from make_golden.py: reference_demos; demos that
save a checkpoint have no reference output in a container without HDF5 and are checked for a rising bound instead)."""
import importlib.util
import os
import re
import sys
import types

import numpy as np
import pytest

from conftest import golden, GOLDEN

sys.path.insert(0, GOLDEN)
from demo_calls import CALLS          # noqa: E402


@pytest.fixture
def as_bayespy():
    from oracle import make_ref
    make_ref.build()
    if not make_ref.available():
        pytest.skip("oracle/_ref is not staged and /root/reference is absent")
    saved = {k: v for k, v in sys.modules.items() if k == "bayespy" or k.startswith("bayespy.") or
             k.startswith("matplotlib")}
    for k in saved:
        del sys.modules[k]
    import bayespy_b200
    bayespy_b200.install_as_bayespy(stub_plotting=True)
    demos = types.ModuleType("bayespy.demos")
    demos.__path__ = [os.path.join(make_ref.DEST, "bayespy", "demos")]
    sys.modules["bayespy.demos"] = demos
    yield demos.__path__[0]
    for k in [k for k in sys.modules if k == "bayespy" or k.startswith("bayespy.") or k.startswith("matplotlib")]:
        del sys.modules[k]
    sys.modules.update(saved)


def _load(path, name):
    spec = importlib.util.spec_from_file_location("bayespy.demos." + name, os.path.join(path, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["bayespy.demos." + name] = m
    spec.loader.exec_module(m)
    return m


# demos whose nodes were added after the round's last GPU session: oracle backend only
HOST_ONLY = {"gamma_shape"}


@pytest.mark.parametrize("demo", sorted(HOST_ONLY))
def test_reference_demo_runs_unchanged_on_the_oracle_backend(oracle_backend, as_bayespy, demo, capsys):
    _run_and_compare(as_bayespy, demo, capsys)


@pytest.mark.parametrize("demo", sorted(set(CALLS) - HOST_ONLY))
def test_reference_demo_runs_unchanged(backend, as_bayespy, demo, capsys):
    _run_and_compare(as_bayespy, demo, capsys)


def _run_and_compare(as_bayespy, demo, capsys):
    m = _load(as_bayespy, demo)
    np.random.seed(1)
    CALLS[demo](m)
    L = np.array([float(v) for v in re.findall(r"(?:loglike=|integrated pdf: )([-+]?(?:[0-9.]+(?:e[-+][0-9]+)?|inf|nan))", capsys.readouterr().out)])
    assert len(L) > 0
    g = golden("demos")
    if demo in g.files:
        ref = g[demo]
        assert len(L) == len(ref)
        # the scripts print seven significant digits
        ok = np.isfinite(ref)
        np.testing.assert_allclose(L[ok], ref[ok], rtol=2e-6)
        # (lssm_sd / lssm_tvd at these toy sizes print -inf / nan in the reference; the bound is then degenerate here too)
        assert not np.any(np.isfinite(L[~ok]))
    elif demo != "stochastic_inference":
        assert np.all(np.isfinite(L))
        # (stochastic VI: the bound of a mini-batch estimate is not monotone)
        assert L[-1] > L[0]
