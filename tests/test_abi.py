"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/fdsb200.h declares. No compute calls here."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    header = open(os.path.join(ROOT, 'include', 'fdsb200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    return sorted(set(re.findall(r'\b(fds_[a-z0-9_]+)\s*\(', header)))


def test_library_exports_every_declared_symbol(library):
    lib = ctypes.CDLL(library)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name


def test_binding_covers_the_header(library):
    from pyfds_b200 import _engine
    lib = _engine.load_library()
    for name in declared_symbols():
        assert getattr(lib, name).argtypes is not None, name


def test_create_without_gpu_reports_an_error(library):
    from pyfds_b200 import _engine
    lib = _engine.load_library()
    if lib.fds_device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(_engine.EngineError) as info:
        _engine.Engine('acoustic2d', 16, 16, 1, False)
    assert 'no CUDA device' in str(info.value)


def test_library_contains_sm100a_code(library):
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '--list-elf', library], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
