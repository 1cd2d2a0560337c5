"""The C-ABI boundary: include/bpk.h, the ctypes prototypes and the built library agree,
and the library refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "bpk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bpk_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_prototypes():
    from bayespy_b200 import _bpk
    assert _header_symbols() == sorted(_bpk.PROTOTYPES)


def test_library_exports_every_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from bayespy_b200 import _bpk
    lib = ctypes.CDLL(_bpk.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(lib, name), name


def test_constants_match_header():
    from bayespy_b200 import _bpk
    src = open(os.path.join(ROOT, "include", "bpk.h")).read()
    for name, val in (("BPK_MAXD", _bpk.MAXD), ("BPK_MAXIN", _bpk.MAXIN), ("BPK_MAXDIM", _bpk.MAXDIM)):
        assert int(re.search(r"#define %s (\d+)" % name, src).group(1)) == val
    body = re.search(r"enum \{(.*?)BPK_OP_COUNT_", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    ops = [m for m in re.findall(r"BPK_OP_([A-Z0-9]+)", body)]
    assert ops == sorted(_bpk.OPS, key=_bpk.OPS.get)


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not compute on the host."""
    from bayespy_b200 import _bpk
    lib = ctypes.CDLL(_bpk.LIB_PATH)
    lib.bpk_init.restype = ctypes.c_int
    rc = lib.bpk_init(0)
    if rc == 0:
        pytest.skip("a GPU is present on this box")
    assert rc == _bpk.ENOGPU
    old = _bpk._set_backend_for_testing(None)
    try:
        with pytest.raises(_bpk.BpkError):
            _bpk.get()
        from bayespy_b200.utils import linalg
        import numpy as np
        with pytest.raises(_bpk.BpkError):
            linalg.chol(np.identity(3))
    finally:
        _bpk._set_backend_for_testing(old)
    # compute entry points refuse too
    lib.bpk_sync.restype = ctypes.c_int
    assert lib.bpk_sync() == _bpk.ENOGPU
