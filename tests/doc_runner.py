"""Run the doctests of one of the reference's documentation files (doc/source/**/*.rst) and print, as JSON, which
examples ran and which failed.  ``ours`` mode answers ``import bayespy`` with this package on the oracle backend;
``reference`` mode imports the unmodified reference.  Plotting statements are skipped in both modes (matplotlib is not
installed), the hidden ``testsetup`` blocks of the Sphinx doctest extension (the seeds) are executed first.

    python tests/doc_runner.py ours|reference /root/reference/doc/source/examples/gmm.rst
"""
import doctest
import json
import os
import re
import sys
import textwrap
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLOTTING = re.compile(r"bpplt|pyplot|plt\.|\.plot\(|\.show\(\)")


def main(mode, path):
    warnings.filterwarnings("ignore")
    sys.path.insert(0, ROOT)
    from bayespy_b200 import _Permissive
    if mode == "ours":
        from bayespy_b200 import _bpk
        from oracle.bpk_ref import RefBackend
        _bpk._set_backend_for_testing(RefBackend())
        import bayespy_b200
        bayespy_b200.install_as_bayespy(stub_plotting=True)
    else:
        from oracle import make_ref
        for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors",
                    "matplotlib.patches", "matplotlib.gridspec"):
            sys.modules.setdefault(mod, _Permissive())
        make_ref.import_reference()
    text = open(path).read()
    globs = {}
    for m in re.finditer(r"\.\. testsetup::\n\n((?:[ ]{3,}.*\n|\n)+)", text):
        exec(textwrap.dedent(m.group(1)), globs)
    test = doctest.DocTestParser().get_doctest(text, globs, os.path.basename(path), path, 0)
    for ex in test.examples:
        if PLOTTING.search(ex.source):
            ex.options[doctest.SKIP] = True
    failed, bounds = [], {}

    def keep_bounds(example, got):
        vals = re.findall(r"loglike=([-+]?(?:[0-9.]+e[-+][0-9]+|inf|nan))", got)
        if vals:
            bounds[str(example.lineno + 1)] = [float(v) for v in vals]

    class Runner(doctest.DocTestRunner):
        def report_success(self, out, test, example, got):
            keep_bounds(example, got)

        def report_failure(self, out, test, example, got):
            keep_bounds(example, got)
            failed.append([example.lineno + 1, example.source.strip()[:80], got.strip()[-400:]])

        def report_unexpected_exception(self, out, test, example, exc_info):
            failed.append([example.lineno + 1, example.source.strip()[:80], "%s: %s" % (exc_info[0].__name__, exc_info[1])])

    runner = Runner(verbose=False, optionflags=doctest.ELLIPSIS | doctest.NORMALIZE_WHITESPACE)
    runner.run(test, out=lambda s: None, clear_globs=False)
    print("DOCRESULT " + json.dumps({"tries": runner.tries, "failed": failed, "bounds": bounds}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
