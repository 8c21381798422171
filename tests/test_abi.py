"""The C-ABI library loads without a GPU and exports every symbol that
include/ocbo_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'ocbo_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ocbo_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported():
    from ocbo_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_library_reports_version_and_workspace_size():
    from ocbo_b200 import _lib
    assert _lib.LIB.ocbo_version() >= 100
    assert _lib.LIB.ocbo_invdiag_elems(1024) == 8 * 128 * 128
    assert _lib.LIB.ocbo_last_error_string() is not None


def test_argument_validation_needs_no_gpu():
    """Argument errors are detected before any CUDA call: -k for argument k."""
    from ocbo_b200 import _lib
    rc = _lib.LIB.ocbo_potrf(None, 128, 128, 0, None, None, 1, None)
    assert rc == -1
    rc = _lib.LIB.ocbo_potrf(ctypes.c_void_p(16), 100, 128, 0, None, None, 1, None)
    assert rc == -2 and b'128' in _lib.LIB.ocbo_last_error_string()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure; the product package must not reach it."""
    pkg = os.path.join(ROOT, 'ocbo_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def test_engine_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from ocbo_b200.gp.gp_interface import make_gp
    with pytest.raises(RuntimeError):
        make_gp([[0.1, 0.2]], [1.0], 'se', 1.0, [0.3, 0.3], 0.01, 0.0)
