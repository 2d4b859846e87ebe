"""Compiles the plain C restatement of the oracle's inner loop (``oracle/dia_matvec.c``) into
``oracle/_ref/liboracle.so`` with gcc. TEST INFRASTRUCTURE: called by ``__graft_entry__.build()`` and by
the oracle's ``backend='c'``; nothing under ``pyfds_b200/`` uses it."""

import ctypes
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(_HERE, 'dia_matvec.c')
OUT_DIR = os.path.join(_HERE, '_ref')
LIBRARY = os.path.join(OUT_DIR, 'liboracle.so')
# -ffp-contract=off: never fuse a*b+c (the reference arithmetic is one rounded multiply, one rounded add)
FLAGS = ['-O2', '-std=c99', '-ffp-contract=off', '-fPIC', '-shared']


def build(force=False):
    """Compile if missing or older than the source; returns the library path."""
    if not force and os.path.exists(LIBRARY) and \
            os.path.getmtime(LIBRARY) >= os.path.getmtime(SOURCE):
        return LIBRARY
    gcc = shutil.which('gcc') or shutil.which('cc')
    if gcc is None:
        raise RuntimeError('gcc not found: the C restatement of the oracle cannot be built')
    os.makedirs(OUT_DIR, exist_ok=True)
    result = subprocess.run([gcc] + FLAGS + ['-o', LIBRARY, SOURCE], capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError('gcc failed:\n' + result.stdout + result.stderr)
    return LIBRARY


QUOTIENT_SOURCE = os.path.join(_HERE, 'quotient_check.c')
QUOTIENT_BINARY = os.path.join(OUT_DIR, 'quotient_check')


def build_quotient_check(force=False):
    """Compiles ``oracle/quotient_check.c`` (the CPU check of the axisymmetric kernel's quotient
    sequence against the IEEE division; run by tests/test_fast_division.py) into ``oracle/_ref/``."""
    if not force and os.path.exists(QUOTIENT_BINARY) and \
            os.path.getmtime(QUOTIENT_BINARY) >= os.path.getmtime(QUOTIENT_SOURCE):
        return QUOTIENT_BINARY
    gcc = shutil.which('gcc') or shutil.which('cc')
    if gcc is None:
        raise RuntimeError('gcc not found: oracle/quotient_check.c cannot be built')
    os.makedirs(OUT_DIR, exist_ok=True)
    flags = ['-O2', '-std=c99', '-ffp-contract=off']
    try:
        with open('/proc/cpuinfo') as handle:
            if ' fma ' in handle.read():
                flags.append('-mfma')   # hardware fma; libm's fma() computes the same, slowly
    except OSError:
        pass
    result = subprocess.run([gcc] + flags + ['-o', QUOTIENT_BINARY, QUOTIENT_SOURCE, '-lm'],
                            capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError('gcc failed:\n' + result.stdout + result.stderr)
    return QUOTIENT_BINARY


_lib = None


def library():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        p = ctypes.c_void_p
        _lib.fds_oracle_dia_matvec.argtypes = [ctypes.c_int64, ctypes.c_int64, p, p, p, p]
        _lib.fds_oracle_dia_matvec.restype = None
    return _lib


if __name__ == '__main__':
    print(build(force=True))
    print(build_quotient_check(force=True))
