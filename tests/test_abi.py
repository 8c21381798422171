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


def test_workspace_sizes_need_no_gpu():
    """The workspace queries of the int8-tensor-core entry points are pure host arithmetic: 0 for shapes that would
    not use the path, 7 bytes per element of the factor (+ row scales) for the sweep's chol(Sigma), and the sampling
    tail's workspace holds the factor's slices first (so ocbo_potrf_ws can leave them there)."""
    from ocbo_b200 import _lib
    L = _lib.LIB
    assert L.ocbo_potrf_ws_bytes(1024, 2) == 0 and L.ocbo_post_cov_ws_bytes(1024, 256, 1) == 0
    assert L.ocbo_sample_ws_bytes(512, 64, 1) == 0
    m, B = 8192, 2
    potrf, cov, smp = L.ocbo_potrf_ws_bytes(m, B), L.ocbo_post_cov_ws_bytes(m, 1024, B), L.ocbo_sample_ws_bytes(m, 256, B)
    assert potrf >= 7 * m * m * B and potrf <= 7 * m * m * B + 2 * (8 * m * B + 512)
    assert cov >= 7 * m * 1024 * B and smp >= potrf + 7 * 256 * m * B
    assert L.ocbo_ozaki_slice_pairs() == 28


def test_int8_kernel_is_compiled_to_tcgen05():
    """SASS evidence in the shipped library (cuobjdump; skipped where the CUDA toolkit is absent): the update kernel
    issues UTCIMMA (tcgen05.mma kind::i8), reads tensor memory with LDTM and is fed by UBLKCP bulk copies; the fp64
    tile core still carries DMMA and UTMALDG."""
    import shutil
    import subprocess
    import pytest
    from ocbo_b200 import _lib
    tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(tool):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([tool, '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    body, name = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            body[name] = []
        elif name:
            body[name].append(line)
    upd = [n for n in body if 'ozaki_update_kernel' in n]
    assert len(upd) == 3                                   # the three modes
    for n in upd:
        text = '\n'.join(body[n])
        assert 'UTCIMMA' in text and 'LDTM' in text and 'UBLKCP' in text, n
    whole = '\n'.join('\n'.join(v) for v in body.values())
    assert 'DMMA.8x8x4' in whole and 'UTMALDG' in whole
