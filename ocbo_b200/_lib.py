"""ctypes binding of ``libocbo_b200.so`` (C ABI in ``include/ocbo_b200.h``).

There is no CPU fallback: if the library cannot be loaded the import fails."""
import ctypes
import os

from . import build as _build

_VP = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_U64 = ctypes.c_uint64

_PROTOS = {
    'ocbo_version': (ctypes.c_int, []),
    'ocbo_launch_count': (ctypes.c_uint64, []),
    'ocbo_invdiag_elems': (_L, [_I]),
    'ocbo_gram': (_I, [_I, _I, _VP, _I, _L, _VP, _VP, _I, _L, _VP, _VP, _L, _VP, _I, _L, _I, _I, _VP]),
    'ocbo_add_diag': (_I, [_VP, _I, _I, _L, _VP, _VP, _I, _VP]),
    'ocbo_diag_max': (_I, [_VP, _I, _I, _L, _VP, _VP, _I, _VP]),
    'ocbo_potrf': (_I, [_VP, _I, _I, _L, _VP, _VP, _I, _VP]),
    'ocbo_potrf_ws_bytes': (_L, [_I, _I]),
    'ocbo_potrf_ws': (_I, [_VP, _I, _I, _L, _VP, _VP, _I, _VP, _L, _VP]),
    'ocbo_potrf_nv': (_I, [_VP, _I, _I, _L, _VP, _VP, _VP, _I, _VP]),
    'ocbo_ladder_restore': (_I, [_VP, _VP, _I, _I, _L, _I, _L, _VP, _VP, ctypes.c_double, _I, _VP]),
    'ocbo_ladder_advance': (_I, [_VP, _VP, _I, _I, _I, _VP]),
    'ocbo_ladder_adopt': (_I, [_VP, _VP, _I, _I, _L, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP]),
    'ocbo_potrf_ws_profile': (_I, [_VP, _I, _I, _L, _VP, _VP, _I, _VP, _L, _VP, _VP]),
    'ocbo_potrf_profile': (_I, [_VP, _I, _I, _L, _VP, _VP, _I, _VP, _VP]),
    'ocbo_trsm_lower': (_I, [_VP, _I, _I, _L, _VP, _VP, _I, _I, _L, _VP, _I, _VP]),
    'ocbo_solve_alpha_lml': (_I, [_VP, _I, _I, _L, _VP, _VP, _L, _VP, _VP, _L, _VP, _L, _VP, _VP, _I, _VP]),
    'ocbo_solve_alpha_w_lml': (_I, [_VP, _I, _I, _L, _VP, _VP, _L, _VP, _VP, _L, _VP, _L, _VP, _L, _VP,
                                    _VP, _I, _VP]),
    'ocbo_post_mean': (_I, [_I, _I, _VP, _I, _L, _VP, _VP, _I, _L, _VP, _VP, _L, _VP, _L, _VP, _L, _I, _VP]),
    'ocbo_post_cov': (_I, [_I, _I, _VP, _I, _L, _VP, _VP, _I, _I, _L, _VP, _L, _VP, _I, _L, _I, _I, _VP]),
    'ocbo_post_cov_ws_bytes': (_L, [_I, _I, _I]),
    'ocbo_post_cov_ws': (_I, [_I, _I, _VP, _I, _L, _VP, _VP, _I, _I, _L, _VP, _L, _VP, _I, _L, _I, _I, _VP, _L, _VP]),
    'ocbo_post_cross_cov': (_I, [_I, _I, _VP, _I, _L, _VP, _VP, _I, _L, _VP, _VP, _I, _L, _VP, _I, _L,
                                 _I, _VP, _L, _VP, _I, _L, _I, _VP]),
    'ocbo_solve_forward_lml': (_I, [_VP, _I, _I, _L, _VP, _VP, _L, _VP, _VP, _L, _VP, _L, _VP, _VP, _I,
                                    _VP]),
    'ocbo_post_mean_var_from_v': (_I, [_VP, _I, _I, _L, _VP, _VP, _L, _VP, _L, _I, _VP, _VP, _L, _VP,
                                       _L, _I, _VP]),
    'ocbo_post_var_ei': (_I, [_VP, _I, _I, _L, _VP, _VP, _L, _VP, _L, _VP, _I, _VP, _VP, _L, _VP, _L,
                              _I, _VP]),
    'ocbo_symmetrize_lower': (_I, [_VP, _I, _I, _L, _I, _VP]),
    'ocbo_philox_normal': (_I, [_VP, _I, _I, _I, _L, _U64, _VP, _I, _VP]),
    'ocbo_sample': (_I, [_VP, _L, _VP, _I, _I, _L, _VP, _I, _I, _L, _VP, _I, _L, _VP, _I, _VP]),
    'ocbo_sample_tile_argmax': (_I, [_VP, _L, _VP, _I, _I, _L, _VP, _I, _I, _L, _VP, _VP, _VP, _I, _VP]),
    'ocbo_sample_ws_bytes': (_L, [_I, _I, _I]),
    'ocbo_sample_tile_argmax_ws': (_I, [_VP, _L, _VP, _I, _I, _L, _VP, _I, _I, _L, _VP, _VP, _VP, _I, _VP, _L, _I, _VP]),
    'ocbo_segment_argmax': (_I, [_VP, _I, _I, _L, _VP, _I, _VP, _VP, _I, _VP]),
    'ocbo_joint_pick': (_I, [_VP, _VP, _I, _I, _I, _I, _VP, _VP, _VP, _I, _VP]),
    'ocbo_kg': (_I, [_VP, _I, _VP, _L, _VP, ctypes.c_double, _I, _I, _I, _VP, _VP]),
    'ocbo_probe_dmma': (_I, [_VP, _I, _I, _VP]),
    'ocbo_probe_dfma': (_I, [_VP, _I, _I, _VP]),
    'ocbo_probe_dgemm_nt': (_I, [_VP, _VP, _VP, _I, _VP]),
    'ocbo_probe_kern_se': (_I, [_VP, ctypes.c_double, _L, _VP, _VP, _VP]),
    'ocbo_probe_igemm_nt': (_I, [_VP, _L, _VP, _L, _VP, _L, _I, _I, _I, _VP]),
    'ocbo_ozaki_ws_bytes': (_L, [_I, _I, _I]),
    'ocbo_ozaki_slice_pairs': (_I, []),
    'ocbo_probe_ozaki_update': (_I, [_VP, _I, _I, _I, _VP, _I, _VP, _L, _I, _VP]),
}
EXPORTED_SYMBOLS = sorted(list(_PROTOS) + ['ocbo_last_error_string'])

TILE = 128
MAX_DIM = 16
KERNEL_SE, KERNEL_MATERN12, KERNEL_MATERN32, KERNEL_MATERN52 = 0, 1, 2, 3
GRAM_SYM, GRAM_LOWER, GRAM_ADD_NOISE = 1, 2, 4


class OcboError(RuntimeError):
    pass


def _load():
    # OCBO_B200_LIB: an experimental build of the same library (tests/tools/arrive_placement.py)
    path = os.environ.get('OCBO_B200_LIB') or _build.LIB_PATH
    if not os.path.exists(path):
        try:
            _build.build()
        except Exception as exc:  # no silent fallback
            raise ImportError('libocbo_b200.so is missing and could not be built (%s); run '
                              '`python -c "import __graft_entry__ as g; g.build()"`' % exc)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib.ocbo_last_error_string.restype = ctypes.c_char_p
    lib.ocbo_last_error_string.argtypes = []
    return lib


LIB = _load()
LIB_PATH = _build.LIB_PATH


def call(name, *args):
    """Invoke an entry point; raise with the library's message on failure."""
    rc = getattr(LIB, name)(*args)
    if rc != 0:
        msg = LIB.ocbo_last_error_string().decode()
        if rc < 0:
            raise ValueError('%s: %s' % (name, msg))
        raise OcboError('%s failed (%d): %s' % (name, rc, msg))
