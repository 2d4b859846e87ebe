"""Builds ``libfdsb200.so`` (the CUDA step engine behind the C ABI in ``include/fdsb200.h``) in-tree
with nvcc for sm_100a. Cross-compiles without a GPU."""

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIBRARY = os.path.join(_HERE, 'libfdsb200.so')
SOURCES = ['fds_abi.cu']
HEADERS = ['fds_common.cuh', 'fds_aux.cuh', 'fds_step1d.cuh', 'fds_line1d.cuh', 'fds_step2d.cuh',
           'fds_stream2d.cuh',
           'fds_streamv.cuh',
           os.path.join('..', '..', 'include', 'fdsb200.h')]

NVCC_FLAGS = [
    '-O3', '-std=c++17',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo',
    # the reference arithmetic is unfused IEEE double: never contract a*b+c
    '-fmad=false',
    '-Xcompiler', '-fPIC', '-shared',
]


def nvcc_path():
    for candidate in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if candidate and os.path.exists(candidate):
            return candidate
    raise RuntimeError('nvcc not found: libfdsb200.so cannot be built')


def is_stale():
    if not os.path.exists(LIBRARY):
        return True
    built = os.path.getmtime(LIBRARY)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force=False, verbose=False, extra_flags=(), output=None):
    """Compile the library if it is missing or older than its sources; returns its path.
    ``output``: build a tuning variant next to the library instead (always compiled; load it with
    FDS_LIBRARY_PATH)."""
    if output is None and not force and not is_stale():
        return LIBRARY
    target = output or LIBRARY
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + \
        ['-o', target] + [os.path.join(CSRC, f) for f in SOURCES] + ['-ldl']
    if verbose:
        print(' '.join(cmd))
    result = subprocess.run(cmd, capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + result.stdout + result.stderr)
    if verbose and (result.stdout or result.stderr):
        print(result.stdout + result.stderr)
    return target


if __name__ == '__main__':
    # python -m pyfds_b200._build [--force] [--ptxas] [--variant NAME -DFLAG=...]
    import sys
    flags = [a for a in sys.argv[1:] if a.startswith('-D')]
    if '--ptxas' in sys.argv:
        flags += ['-Xptxas', '-v']
    out = None
    if '--variant' in sys.argv:
        out = os.path.join(_HERE, 'libfdsb200_{}.so'.format(sys.argv[sys.argv.index('--variant') + 1]))
    print(build_library(force='--force' in sys.argv, verbose=True, extra_flags=flags, output=out))
