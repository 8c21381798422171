"""Build libocbo_b200.so in-tree with nvcc for sm_100a (no JIT cache: the built
library travels to the GPU box with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libocbo_b200.so')
SOURCES = ['gemm_kernels.cu', 'gram.cu', 'potrf.cu', 'potrf_diag_reg.cu', 'potrf_diag_blk.cu', 'ladder.cu', 'sampling.cu', 'ei.cu', 'kg.cu',
           'igemm_probe.cu', 'ozaki.cu', 'capi.cu']
HEADERS = ['common.cuh', 'gemm_core.cuh', 'kernels.h', 'philox.cuh', 'tma.cuh',
           os.path.join('..', '..', 'include', 'ocbo_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-extended-lambda', '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out_dir=None):
    """Compile every CUDA source of the package into ``lib/libocbo_b200.so`` (``defines`` /
    ``out_dir``: an experimental variant of the library elsewhere, tests/tools only)."""
    lib_dir = out_dir or LIB_DIR
    lib_path = os.path.join(lib_dir, 'libocbo_b200.so')
    if out_dir is None and not force and not _stale():
        return LIB_PATH
    os.makedirs(lib_dir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(lib_dir, src.replace('.cu', '.o'))
        cmd = [_nvcc()] + NVCC_FLAGS + ['-D' + d for d in defines] + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s' % src)
    cmd = [_nvcc(), '-shared', '-o', lib_path] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.check_call(cmd)
    return lib_path


HOSTPACK_SRC = os.path.join(CSRC, 'hostpack.c')


def hostpack_path():
    import sysconfig
    return os.path.join(LIB_DIR, '_hostpack' + (sysconfig.get_config_var('EXT_SUFFIX') or '.so'))


def build_hostpack(force=False):
    """gcc: the CPython staging helper (csrc/hostpack.c; host-side copies only, no CUDA)."""
    import sysconfig
    import numpy as np
    out = hostpack_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(HOSTPACK_SRC):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [os.environ.get('CC', 'gcc'), '-O2', '-shared', '-fPIC', '-I', sysconfig.get_paths()['include'],
           '-I', np.get_include(), HOSTPACK_SRC, '-o', out]
    subprocess.check_call(cmd)
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
    print(build_hostpack(force='--force' in sys.argv))
