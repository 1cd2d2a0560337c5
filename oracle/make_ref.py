"""TEST INFRASTRUCTURE / CPU BASELINE — recipe that stages the UNMODIFIED reference package for the
GPU box.

    python oracle/make_ref.py            # /root/reference/bayespy -> oracle/_ref/bayespy (+ two stub modules)

The reference is pure Python (setup.py:39-45: numpy, scipy, h5py, truncnorm; no native code), so
"building" it is a byte-for-byte copy of its package directory.  The copy lands in ``oracle/_ref/``,
which is listed in ``.gitignore`` (it never enters the history of this repo) but NOT in
``.gpurunignore`` (it travels to the GPU box with the snapshot, like the built ``libbpk.so``).
``h5py`` and ``truncnorm`` are not installed in this image and there is no network; both are
imported at module import time by the reference (vmp.py:11, stochastic.py:12, gaussian.py:15) but
only used off the hot path (HDF5 save/load; truncated Gaussians), so two stub modules are written
next to the copy (same stubs as tests/golden/_stubs).

Used by: ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg (kind "reference"),
and tests/test_reference_arm.py.  Nothing under ``bayespy_b200/`` imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = "/root/reference/bayespy"

H5PY_STUB = '''"""Stub: h5py is absent from this image; only VB.save/load (vmp.py:237-356) use it."""
class _V:
    hdf5_version_tuple = (1, 14, 0)
    version = "0.0-stub"
version = _V()
def File(*a, **k):
    raise RuntimeError("h5py stub: HDF5 IO is not available in this container")
'''
TRUNCNORM_STUB = '''"""Stub: truncnorm is only used for truncated Gaussians (gaussian.py:428-438), not on the path."""
def moments(*a, **k):
    raise RuntimeError("truncnorm stub")
'''


def available():
    return os.path.isfile(os.path.join(DEST, "bayespy", "__init__.py"))


def build(force=False):
    """Stage the reference if its sources are present (build container); a no-op on the GPU box."""
    if not os.path.isdir(SRC):
        return DEST if available() else None
    marker = os.path.join(DEST, ".staged_from")
    if available() and not force and os.path.exists(marker):
        return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(SRC, os.path.join(DEST, "bayespy"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for name, text in (("h5py", H5PY_STUB), ("truncnorm", TRUNCNORM_STUB)):
        try:
            __import__(name)
            continue                      # the real module exists: no stub
        except Exception:
            pass
        os.makedirs(os.path.join(DEST, name))
        with open(os.path.join(DEST, name, "__init__.py"), "w") as f:
            f.write(text)
    with open(marker, "w") as f:
        f.write(SRC + "\n")
    return DEST


def import_reference():
    """Import the staged reference package (``bayespy``) and return the module."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` in the build container")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import warnings
    warnings.filterwarnings("ignore")
    import bayespy
    if not os.path.abspath(bayespy.__file__).startswith(DEST):
        raise RuntimeError("a different bayespy is on sys.path: %s" % bayespy.__file__)
    return bayespy


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
