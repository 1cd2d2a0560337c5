"""The reference's documentation (doc/source/user_guide/*.rst, doc/source/examples/*.rst) as doctests, run twice: with
the unmodified reference and with this package answering ``import bayespy`` (tests/doc_runner.py, one clean interpreter
each, side by side).  Every example that passes with the reference must pass here; where a printed VB trajectory does not match the
documentation text (the rotation optimiser inside the loop amplifies round-off, so the iteration at which the
convergence test fires can move by one or two — it also does between the reference's own documentation and the
reference run in this environment), the bounds printed here are compared with the bounds the reference prints.

Needs the documentation sources (/root/reference/doc), which exist in the build container only; host-logic test."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = "/root/reference/doc/source"
FILES = sorted(f[len(DOCS) + 1:] for f in glob.glob(DOCS + "/user_guide/*.rst") + glob.glob(DOCS + "/examples/*.rst")
               if ">>>" in open(f).read())                       # the files that hold doctest examples

# Examples whose expected text cannot match here for a reason that is not arithmetic.
NOT_COMPARABLE = {
    ("user_guide/inference.rst", "Q['X']"): "prints the node's class path (bayespy.inference.vmp.nodes.gaussian...)",
}


def _start(mode, rel):
    return subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "doc_runner.py"), mode, os.path.join(DOCS, rel)],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT)


def _result(proc):
    out, err = proc.communicate(timeout=900)
    lines = [l for l in out.splitlines() if l.startswith("DOCRESULT ")]
    assert lines, err[-2000:]
    return json.loads(lines[-1][len("DOCRESULT "):])


@pytest.mark.skipif(not os.path.isdir(DOCS), reason="the reference's documentation sources are not present")
@pytest.mark.parametrize("rel", FILES or ["none"])
def test_documentation_examples_run_like_with_the_reference(rel):
    from oracle import make_ref
    make_ref.build()
    procs = _start("ours", rel), _start("reference", rel)          # the two interpreters run side by side
    ours, ref = _result(procs[0]), _result(procs[1])
    assert ours["tries"] == ref["tries"]
    ref_failed = {l for l, _, _ in ref["failed"]}
    for lineno, source, got in ours["failed"]:
        if lineno in ref_failed or (rel, source) in NOT_COMPARABLE:
            continue                     # the documentation text does not match the reference in this environment either
        mine, theirs = ours["bounds"].get(str(lineno)), ref["bounds"].get(str(lineno))
        assert mine and theirs, "%s:%d  %s\n%s" % (rel, lineno, source, got)
        # a VB run: same trajectory up to the optimiser's round-off, convergence within two iterations of each other
        n = min(len(mine), len(theirs))
        assert abs(len(mine) - len(theirs)) <= 2, (rel, lineno, len(mine), len(theirs))
        # (with a rotation after every iteration the paths separate by up to ~1e-3 in the steep part and meet again)
        np.testing.assert_allclose(mine[:n], theirs[:n], rtol=3e-3, err_msg="%s:%d %s" % (rel, lineno, source))
        np.testing.assert_allclose(mine[-1], theirs[-1], rtol=3e-4)
